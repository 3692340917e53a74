"""Run under torchrun (one process per GPU): data-parallel update == single-GPU update (SURVEY.md §8e "Equivalence").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_equivalence.py

Shard mode 0: every rank holds the same buffer and the same host rand() permutation chain and takes rows
[rank*mb/G, (rank+1)*mb/G) of each global minibatch; gradients are summed with the NCCL all-reduce inside
libppo_b200.so.  Afterwards every rank repeats the update alone (communicator torn down) from the same initial
state.  Integer work (Adam step counts, rand() stream position) must be identical, weights within 1e-5 norm-wise
(only the summation order of the cross-rank reduction differs).  Shard mode 1 (rank-local buffers, weak scaling)
is checked for cross-rank consistency: all ranks must end with bit-identical weights."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import b200
import cabi
import oracle
from test_gpu_train import fill_host_buffer, make_ppo, synthetic_buffer


def run_update(L, sizes, acts, n, mb, seed, buf_seed, ent=0.0, npol=2, nval=2):
    cabi.srand(seed)
    ppo = make_ppo(L, sizes, acts, n, ent=ent)
    T = oracle.Trainer(sizes, acts, batch_size=mb, n_epochs_policy=2, n_epochs_value=2, init=False)
    T.mu[:] = b200.nn_get_params(L, ppo.contents.policy.contents.mu)     # logprob_old from the same initial policy
    b = synthetic_buffer(T, np.random.default_rng(buf_seed), sizes, acts, n)
    fill_host_buffer(ppo, b)
    L.ppo_b200_set_permutation_mode(ppo, 0, 0)
    cabi.srand(seed + 1)
    L.ppo_b200_update(ppo, 0.99, mb, npol, nval)
    after = cabi.rand()
    out = (b200.nn_get_params(L, ppo.contents.V).copy(), b200.nn_get_params(L, ppo.contents.policy.contents.mu).copy(),
           np.ctypeslib.as_array(ppo.contents.policy.contents.log_std, shape=(sizes[-1],)).copy(),
           ppo.contents.adam_V.contents.time_step, after)
    L.free_ppo(ppo)
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    L = b200.lib()
    L.ppo_b200_set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for sizes, acts, n, mb, ent in [([3, 64, 64, 1], ["tanh", "tanh", "none"], 4096, 512, 0.0),        # persistent phase kernel + NVLink peer exchange
                                    ([3, 64, 64, 1], ["tanh", "tanh", "none"], 4096, 1024, 0.01),      # entropy bonus under DP (src/ppo.cu:436-438)
                                    ([17, 32, 32, 6], ["relu", "relu", "none"], 4096, 1024, 0.01),     # A = 6, generic loss head
                                    ([17, 256, 256, 6], ["relu", "relu", "none"], 4096, 1024, 0.01)]:  # layer-wise kernels + NCCL all-reduce
        b200.package().dist_init_from_torch(L)
        L.ppo_b200_dist_set_shard_mode(0)
        dp = run_update(L, sizes, acts, n, mb, 5, 77, ent)
        L.ppo_b200_dist_set_shard_mode(1)
        weak = run_update(L, sizes, acts, n, mb // world, 5, 100 + rank, ent)      # different buffer per rank
        L.ppo_b200_dist_finalize()
        single = run_update(L, sizes, acts, n, mb, 5, 77, ent)
        for name, a, s in zip(("V", "mu", "log_std"), dp[:3], single[:3]):
            err = float(np.max(np.abs(a - s)) / max(np.max(np.abs(s)), 1e-30))
            frac = float(np.mean(np.abs(a - s) > 1e-5 * np.max(np.abs(s))))
            print("rank %d %s %s: DP vs single-GPU norm-wise err %.2e, fraction > 1e-5: %.4f" % (rank, sizes, name, err, frac), flush=True)
            ok &= err < 1e-4 and frac < 0.01
        ok &= dp[3] == single[3] and dp[4] == single[4]          # Adam step count, rand() stream position
        # weak-scaling mode: every rank must hold the same weights bit for bit
        for a in weak[:3]:
            t = torch.from_numpy(a.copy()).cuda()
            ref = t.clone()
            dist.broadcast(ref, 0)
            ok &= bool(torch.equal(t, ref))
    # ---- soak: more than 2^16 consecutive gradient exchanges per net through the {tag, value} parity protocol of the phase
    # kernel (one 64-row tile per rank and minibatch, tiny buffer, thousands of epochs): the replicas must stay bit-identical
    # and finite.  DIST_SOAK_EXCHANGES=0 skips it.
    target = int(os.environ.get("DIST_SOAK_EXCHANGES", "66000"))
    if target > 0:
        sizes, acts, n = [3, 64, 64, 1], ["tanh", "tanh", "none"], 1024
        mb = 64 * world
        epochs = -(-target // (n // mb))
        b200.package().dist_init_from_torch(L)
        L.ppo_b200_dist_set_shard_mode(0)
        t0 = __import__("time").time()
        soak = run_update(L, sizes, acts, n, mb, 9, 55, 0.0, npol=epochs, nval=epochs)
        dt = __import__("time").time() - t0
        L.ppo_b200_dist_finalize()
        same = True
        for a in soak[:3]:
            t = torch.from_numpy(a.copy()).cuda()
            ref = t.clone()
            dist.broadcast(ref, 0)
            same &= bool(torch.equal(t, ref)) and bool(np.isfinite(a).all())
        print("rank %d soak: %d exchanges per net (%d epochs x %d minibatches, world %d) in %.1f s, Adam steps %d, replicas identical: %s"
              % (rank, epochs * (n // mb), epochs, n // mb, world, dt, soak[3], same), flush=True)
        ok &= same and soak[3] == epochs * (n // mb)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_EQUIVALENCE_OK" if int(flag.item()) == 1 else "DIST_EQUIVALENCE_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
