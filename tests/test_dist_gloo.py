"""Host-side logic of the data-parallel path under torch.distributed/gloo, world_size 2, on CPU.

The product's collectives are NCCL calls inside libppo_b200.so (csrc/dist.cu) and need GPUs; what can be
checked here is everything AROUND them, with gloo standing in for NCCL as the transport:
  * the row split of a global minibatch (ppo_c_b200.shard_rows == update_device in csrc/ppo.cu) and the
    loss-gradient scaling by the GLOBAL minibatch size: sum over ranks of the rank-local gradients ==
    the single-process gradient (SURVEY.md §8e "Equivalence"), integers (row indices) identical;
  * the ordered Welford merge of per-rank (mean, M2, n) triples (gae_merge_ranks_kernel) == global stats;
  * the out-of-band broadcast of rank 0's 128-byte communicator id (dist_init_from_torch's transport).
The arithmetic on each rank is the oracle's (test infrastructure), not the product's.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200
import cabi
import oracle

f32 = np.float32
SIZES, ACTS = [3, 16, 16, 1], ["relu", "relu", "none"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    cabi.srand(3)
    params = oracle.init_params(SIZES)
    rng = np.random.default_rng(0)
    n, mb = 1000, 64
    state = rng.standard_normal((n, 3)).astype(f32)
    target = rng.standard_normal(n).astype(f32)
    cabi.srand(4)
    perm = oracle.shuffle(n)
    return params, state, target, perm, n, mb


def _value_grad(params, x, y_true, m_total):
    """MSE gradient of the value net on rows x, scaled by the GLOBAL minibatch size (loss.cu:16-23)."""
    y, cache = oracle.mlp_forward(params, SIZES, ACTS, x)
    g = (2.0 * (y.ravel() - y_true) / m_total).astype(f32).reshape(-1, 1)
    return oracle.mlp_backward(params, SIZES, ACTS, cache, g)


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk = b200.package()
    params, state, target, perm, n, mb = _problem()
    res = {}
    # ---- (1) gradient equivalence over two global minibatches
    grads = []
    for k in range(2):
        if mode == 0:      # same buffer + permutation on every rank, rows [rank*mb/G, (rank+1)*mb/G)
            row0, local, total = pk.shard_rows(mb, rank, world, 0)
            rows = perm[(k * mb + row0 + np.arange(local)) % n]
        else:              # own rows per rank; the global minibatch is the union
            row0, local, total = pk.shard_rows(mb // world, rank, world, 1)
            rows = perm[(k * mb + rank * local + np.arange(local)) % n]
        g = torch.from_numpy(_value_grad(params, state[rows], target[rows], total))
        dist.all_reduce(g)                                   # stands in for ncclAllReduce(sum) of the flat gradient
        grads.append(g.numpy().copy())
        res.setdefault("rows", []).append(rows.copy())
    res["grads"] = grads
    # ---- (2) ordered Welford merge of per-rank advantage statistics
    rng = np.random.default_rng(10 + rank)
    adv = rng.standard_normal(5000 + 777 * rank) * (1 + rank) + 0.3 * rank
    mine = torch.tensor([adv.mean(), ((adv - adv.mean()) ** 2).sum(), float(adv.size)], dtype=torch.float64)
    gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, mine)                          # stands in for ncclAllGather of the triples
    res["welford"] = pk.welford_merge([tuple(t.tolist()) for t in gathered])
    res["adv"] = adv
    # ---- (3) out-of-band broadcast of a 128-byte id
    raw = bytes(range(128)) if rank == 0 else bytes(128)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    res["id"] = bytes(t.numpy().tobytes())
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_world2_gradient_welford_and_id_broadcast(mode):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    params, state, target, perm, n, mb = _problem()
    for k in range(2):
        rows = perm[(k * mb + np.arange(mb)) % n]
        want = _value_grad(params, state[rows], target[rows], mb)
        # integer work: the union of the rank-local row indices is exactly the global minibatch, in order
        assert np.array_equal(np.concatenate([out[r]["rows"][k] for r in range(world)]), rows)
        for r in range(world):
            got = out[r]["grads"][k]
            assert np.array_equal(got, out[0]["grads"][k])                  # every rank ends with the same sum
            assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))  # == single-process gradient (fp32 order)
    allv = np.concatenate([out[r]["adv"] for r in range(world)])
    mean, m2, cnt = out[0]["welford"]
    assert cnt == allv.size and abs(mean - allv.mean()) < 1e-12 and abs(m2 / cnt - allv.var()) < 1e-10
    assert out[1]["welford"] == out[0]["welford"]
    assert out[1]["id"] == bytes(range(128)) == out[0]["id"]


def test_shard_rows_contract():
    pk = b200.package()
    assert pk.shard_rows(64, 0, 1) == (0, 64, 64)
    assert [pk.shard_rows(64, r, 4, 0) for r in range(4)] == [(0, 16, 64), (16, 16, 64), (32, 16, 64), (48, 16, 64)]
    assert pk.shard_rows(64, 3, 8, 1) == (0, 64, 512)
    with pytest.raises(ValueError):
        pk.shard_rows(10, 0, 4, 0)
