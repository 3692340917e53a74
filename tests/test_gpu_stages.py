"""Stage-level parity of the CUDA path against the oracle, through the C ABI (-m gpu).

Tolerance: per-tensor norm-wise error max|a-b|/max|b| <= 1e-5 for fp32 results
(TOL below); integer / index / byte work and Adam are compared bit-exactly."""
import ctypes as C

import numpy as np
import pytest

import b200
import cabi
import oracle
from conftest import nerr

pytestmark = pytest.mark.gpu
TOL = 1e-5
f32, u8, i32 = np.float32, np.uint8, np.int32
RELU3 = ["relu", "relu", "none"]


@pytest.fixture(scope="module")
def L():
    lib = b200.lib()
    assert lib.ppo_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return lib


def synth(rng, T, N, p_term=1e-3, trunc_every=1000):
    n = T * N
    r, v, vn = (rng.standard_normal(n).astype(f32) for _ in range(3))
    term = (rng.random(n) < p_term).astype(u8)
    t = np.tile(np.arange(T), N)
    trunc = (((t + 1) % trunc_every) == 0).astype(u8)
    trunc[t == T - 1] = 1
    return r, v, vn, term, trunc


def run_gae(L, r, v, vn, term, trunc, gamma=0.99, lam=0.95, normalize=True):
    n = r.shape[0]
    d = [b200.dev(x) for x in (r, v, vn, term, trunc)]
    adv, tgt, stats = b200.dev_empty(n), b200.dev_empty(n), b200.dev_empty(2)
    L.ppo_b200_gae(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, n, gamma, lam, adv.ptr, tgt.ptr, int(normalize), stats.ptr)
    out = adv.numpy(), tgt.numpy(), stats.numpy()
    for x in d + [adv, tgt, stats]:
        x.free()
    return out


# ------------------------------------------------------------------------------------------ GAE
@pytest.mark.parametrize("tag", ["pend", "ragged", "long"])
def test_gae_vs_reference_golden(L, golden, tag):
    a = {k: golden[f"gae_{tag}_{k}"] for k in ["r", "v", "vn", "term", "trunc"]}
    adv, tgt, stats = run_gae(L, a["r"], a["v"], a["vn"], a["term"], a["trunc"])
    assert nerr(tgt, golden[f"gae_{tag}_adv_target"]) < TOL
    # normalised advantages: the float64 restatement is the arbiter ...
    raw64, tgt64, norm64, mean64, std64 = oracle.gae_f64(a["r"], a["v"], a["vn"], a["term"], a["trunc"], 0.99, 0.95)
    assert np.max(np.abs(adv - norm64)) < TOL
    assert abs(stats[0] - mean64) < 1e-6 and abs(stats[1] - std64) / std64 < 1e-6
    # ... and at these sizes the reference's own float result agrees too
    assert np.max(np.abs(adv - golden[f"gae_{tag}_adv_norm"])) < 5e-5


@pytest.mark.parametrize("n,pattern", [(1, "end"), (5, "end"), (511, "end"), (512, "end"), (513, "end"),
                                        (4096, "end"), (4097, "end"), (20000, "end"), (20000, "sparse"),
                                        (33333, "dense"), (100000, "every512"), (3000, "pend")])
def test_gae_done_patterns_raw(L, n, pattern):
    """Edge cases: tiny / ragged sizes, chunk-aligned and unaligned dones, a single episode spanning
    dozens of 512-element chunks (forces the decoupled look-back to walk)."""
    rng = np.random.default_rng(n)
    r, v, vn = (rng.standard_normal(n).astype(f32) for _ in range(3))
    term, trunc = np.zeros(n, u8), np.zeros(n, u8)
    if pattern == "sparse":
        term[rng.integers(0, n, 3)] = 1
    elif pattern == "dense":
        term = (rng.random(n) < 0.3).astype(u8)
        trunc = (rng.random(n) < 0.3).astype(u8)
    elif pattern == "every512":
        trunc[511::512] = 1
    elif pattern == "pend":
        trunc[199::200] = 1
    trunc[-1] = 1
    adv, tgt, _ = run_gae(L, r, v, vn, term, trunc, normalize=False)
    raw, tgt_o, *_ = oracle.gae(r, v, vn, term, trunc, 0.99, 0.95)
    raw64, tgt64, *_ = oracle.gae_f64(r, v, vn, term, trunc, 0.99, 0.95)
    assert nerr(adv, raw) < TOL and nerr(tgt, tgt_o) < TOL
    assert nerr(adv, raw64) < TOL and nerr(tgt, tgt64) < TOL


def test_gae_done_masks_are_exact(L):
    """Integer work: wherever a step is done the advantage must equal delta exactly (no leakage)."""
    rng = np.random.default_rng(7)
    r, v, vn, term, trunc = synth(rng, 300, 40, p_term=0.02, trunc_every=100)
    adv, tgt, _ = run_gae(L, r, v, vn, term, trunc, normalize=False)
    done = (term | trunc).astype(bool)
    delta = r + f32(0.99) * vn * (1 - term).astype(f32) - v
    assert np.array_equal(adv[done], delta[done])


def test_gae_large_sampled_envs(L):
    """BASELINE config 5 shape family (T=2048, N scaled to 8192 here; bench.py runs N=65536):
    whole-buffer run, then the oracle re-derives 48 randomly chosen env streams (streams are
    independent because every env ends with a done) + linearity in the rewards."""
    T, N = 2048, 8192
    rng = np.random.default_rng(5)
    r, v, vn, term, trunc = synth(rng, T, N)
    adv, tgt, stats = run_gae(L, r, v, vn, term, trunc, normalize=False)
    for e in rng.integers(0, N, 48):
        s = slice(e * T, (e + 1) * T)
        raw, tg, *_ = oracle.gae(r[s], v[s], vn[s], term[s], trunc[s], 0.99, 0.95)
        assert nerr(adv[s], raw) < TOL and nerr(tgt[s], tg) < TOL
    # float64-combined statistics vs numpy float64
    assert abs(stats[0] - adv.astype(np.float64).mean()) < 1e-6
    assert abs(stats[1] - adv.astype(np.float64).std()) / adv.astype(np.float64).std() < 1e-6
    # linearity: GAE(r1 + r2, v=0) = GAE(r1, 0) + GAE(r2, 0)
    z = np.zeros_like(r)
    r2 = rng.standard_normal(T * N).astype(f32)
    a1, _, _ = run_gae(L, r, z, z, term, trunc, normalize=False)
    a2, _, _ = run_gae(L, r2, z, z, term, trunc, normalize=False)
    a12, _, _ = run_gae(L, r + r2, z, z, term, trunc, normalize=False)
    assert nerr(a12, a1 + a2) < TOL


def test_compute_gae_cuda_reference_entry_point(L, golden):
    """compute_gae_cuda(V, buffer, ...) through the reference structs, identity V-net recipe."""
    a = {k: golden[f"gae_pend_{k}"] for k in ["r", "v", "vn", "term", "trunc"]}
    n = a["r"].shape[0]
    V = L.create_neural_network(cabi.int_array([1, 1]), cabi.cstr_array(["none"]), 2)
    b200.nn_set_params(L, V, np.array([1.0, 0.0], f32))
    buf = L.create_trajectory_buffer(n, 1, 1)
    b = buf.contents
    np.ctypeslib.as_array(b.state_p, shape=(n,))[:] = a["v"]
    np.ctypeslib.as_array(b.next_state_p, shape=(n,))[:] = a["vn"]
    np.ctypeslib.as_array(b.reward_p, shape=(n,))[:] = a["r"]
    np.ctypeslib.as_array(b.terminated_p, shape=(n,))[:] = a["term"].astype(bool)
    np.ctypeslib.as_array(b.truncated_p, shape=(n,))[:] = a["trunc"].astype(bool)
    b.idx, b.full = 0, True
    L.buffer_to_device(buf)
    L.compute_gae_cuda(V, buf, 0.99, 0.95, 200)
    L.buffer_to_host(buf)
    adv = np.ctypeslib.as_array(b.advantage_p, shape=(n,)).copy()
    tgt = np.ctypeslib.as_array(b.adv_target_p, shape=(n,)).copy()
    assert nerr(tgt, golden["gae_pend_adv_target"]) < TOL
    assert np.max(np.abs(adv - golden["gae_pend_adv_norm"])) < 5e-5
    # host-pointer twin gives the same
    L.compute_gae(V, buf, 0.99, 0.95)
    assert np.array_equal(np.ctypeslib.as_array(b.advantage_p, shape=(n,)), adv)
    L.free_trajectory_buffer(buf, True)
    L.free_neural_network(V)


# ------------------------------------------------------------------------------------------ Adam
def _ulp_close(a, b, max_ulp=1, max_frac=1e-3):
    ai, bi = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    d = np.abs(ai - bi)
    return d.max() <= max_ulp and (d > 0).mean() <= max_frac


def test_adam_flat_bitexact_golden(L, golden):
    w, m, v = b200.dev(golden["adam_w0"]), b200.dev(np.zeros(257, f32)), b200.dev(np.zeros(257, f32))
    for t, g in enumerate(golden["adam_grads"], 1):
        dg = b200.dev(g)
        L.ppo_b200_adam_flat(w.ptr, dg.ptr, m.ptr, v.ptr, 257, 3e-4, 0.9, 0.999, t)
        dg.free()
    assert np.array_equal(m.numpy(), golden["adam_m"])
    assert _ulp_close(v.numpy(), golden["adam_v"]) and _ulp_close(w.numpy(), golden["adam_w"])


@pytest.mark.parametrize("n", [1, 3, 1000, 4481, 1 << 20])
def test_adam_flat_vs_oracle(L, n):
    rng = np.random.default_rng(n)
    w0 = rng.standard_normal(n).astype(f32)
    wo, mo, vo, t = w0.copy(), np.zeros(n, f32), np.zeros(n, f32), 0
    w, m, v = b200.dev(w0), b200.dev(mo), b200.dev(vo)
    for step in range(3):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 2)).astype(f32)
        t = oracle.adam(wo, g, mo, vo, 3e-4, t)
        dg = b200.dev(g)
        L.ppo_b200_adam_flat(w.ptr, dg.ptr, m.ptr, v.ptr, n, 3e-4, 0.9, 0.999, t)
        dg.free()
    assert np.array_equal(m.numpy(), mo)
    assert _ulp_close(v.numpy(), vo) and _ulp_close(w.numpy(), wo, max_ulp=2)


def test_adam_update_cuda_tensor_list(L):
    """create_adam_cuda over a NON-contiguous tensor list (the reference's multi-tensor form)."""
    rng = np.random.default_rng(3)
    lens = [5, 2049, 64]
    ws = [rng.standard_normal(k).astype(f32) for k in lens]
    gs = [rng.standard_normal(k).astype(f32) for k in lens]
    dw, dg = [b200.dev(x) for x in ws], [b200.dev(x) for x in gs]
    wp = (cabi.c_float_p * 3)(*[x.fp() for x in dw])
    gp = (cabi.c_float_p * 3)(*[x.fp() for x in dg])
    ad = L.create_adam_cuda(wp, gp, cabi.int_array(lens), 3, sum(lens), 0.9, 0.999)
    for _ in range(2):
        L.adam_update_cuda(ad, 1e-3)
    wo, go = np.concatenate(ws), np.concatenate(gs)
    mo, vo, t = np.zeros_like(wo), np.zeros_like(wo), 0
    for _ in range(2):
        t = oracle.adam(wo, go, mo, vo, 1e-3, t)
    got = np.concatenate([x.numpy() for x in dw])
    assert ad.contents.time_step == 2 and _ulp_close(got, wo, max_ulp=2)
    L.free_adam_cuda(ad)


# ------------------------------------------------------------------------------- permutation / gather
def _fill_buffer(L, state, action, logprob, adv, advt):
    n, S = state.shape
    A = action.shape[1]
    buf = L.create_trajectory_buffer(n, S, A)
    b = buf.contents
    np.ctypeslib.as_array(b.state_p, shape=(n, S))[:] = state
    np.ctypeslib.as_array(b.action_p, shape=(n, A))[:] = action
    np.ctypeslib.as_array(b.logprob_p, shape=(n,))[:] = logprob
    np.ctypeslib.as_array(b.advantage_p, shape=(n,))[:] = adv
    np.ctypeslib.as_array(b.adv_target_p, shape=(n,))[:] = advt
    np.ctypeslib.as_array(b.next_state_p, shape=(n, S))[:] = 0
    np.ctypeslib.as_array(b.reward_p, shape=(n,))[:] = 0
    np.ctypeslib.as_array(b.terminated_p, shape=(n,))[:] = False
    np.ctypeslib.as_array(b.truncated_p, shape=(n,))[:] = False
    b.idx, b.full = 0, True
    return buf


def test_shuffle_and_gather_bitexact_golden(L, golden):
    g = golden
    buf = _fill_buffer(L, g["perm_state"], g["perm_action"], g["perm_logprob"], g["perm_adv"], g["perm_advt"])
    n, mb = g["perm_state"].shape[0], int(g["perm_mb"][0])
    L.buffer_to_device(buf)
    cabi.srand(int(g["perm_seed"][0]))
    outs = {}
    for s in range(2):
        L.shuffle_buffer_cuda(buf)
        perm = b200.d2h(L, buf.contents.random_idx, (n,), i32)
        assert np.array_equal(perm, g["perm_perms"][s])
        for k in range(n // mb):
            o = [b200.dev_empty((mb, 3)), b200.dev_empty((mb, 2)), b200.dev_empty(mb), b200.dev_empty(mb), b200.dev_empty(mb)]
            L.get_batch_cuda(buf, k, mb, *[x.fp() for x in o])
            outs[s * (n // mb) + k] = [x.numpy() for x in o]
            for x in o:
                x.free()
    assert np.array_equal(outs[0][0], g["perm_b0_states"]) and np.array_equal(outs[0][1], g["perm_b0_actions"])
    assert np.array_equal(outs[0][2], g["perm_b0_logprob"])
    assert np.array_equal(outs[7][0], g["perm_b7_states"]) and np.array_equal(outs[7][4], g["perm_b7_advt"])
    L.free_trajectory_buffer(buf, True)


@pytest.mark.parametrize("S,A,n,mb", [(3, 1, 3000, 64), (17, 6, 5000, 4096), (1, 1, 10, 10), (24, 4, 777, 100)])
def test_gather_vs_oracle(L, S, A, n, mb):
    rng = np.random.default_rng(S * 100 + A)
    st, ac = rng.standard_normal((n, S)).astype(f32), rng.standard_normal((n, A)).astype(f32)
    lp, ad, at = (rng.standard_normal(n).astype(f32) for _ in range(3))
    idx = rng.permutation(n).astype(i32)
    d = [b200.dev(x) for x in (idx, st, ac, lp, ad, at)]
    for k in [0, n // mb - 1]:
        o = [b200.dev_empty((mb, S)), b200.dev_empty((mb, A)), b200.dev_empty(mb), b200.dev_empty(mb), b200.dev_empty(mb)]
        L.ppo_b200_gather(d[0].ptr, k * mb, n, mb, S, A, *[x.ptr for x in d[1:]], *[x.ptr for x in o])
        ref = oracle.get_batch(idx, k, mb, st, ac, lp, ad, at)
        for got, want in zip(o, ref):
            assert np.array_equal(got.numpy(), want)


@pytest.mark.parametrize("S,A,n,mb", [(40, 8, 300, 37), (3, 1, 9, 1), (64, 2, 1000, 1000)])
def test_gather_wide_rows_and_odd_batches(L, S, A, n, mb):
    """Packed rows wider than one warp (S + A + 3 > 32), minibatches that are not a multiple of the 8 rows a warp
    takes per iteration, and the wrap-around of (offset + i) % limit (src/trajectory_buffer.cu:171-173)."""
    rng = np.random.default_rng(n + mb)
    st, ac = rng.standard_normal((n, S)).astype(f32), rng.standard_normal((n, A)).astype(f32)
    lp, ad, at = (rng.standard_normal(n).astype(f32) for _ in range(3))
    idx = rng.permutation(n).astype(i32)
    d = [b200.dev(x) for x in (idx, st, ac, lp, ad, at)]
    for offset in [0, n - mb // 2 - 1]:                 # the second one wraps
        o = [b200.dev_empty((mb, S)), b200.dev_empty((mb, A)), b200.dev_empty(mb), b200.dev_empty(mb), b200.dev_empty(mb)]
        L.ppo_b200_gather(d[0].ptr, offset, n, mb, S, A, *[x.ptr for x in d[1:]], *[x.ptr for x in o])
        rows = idx[(offset + np.arange(mb)) % n]
        for got, want in zip(o, (st[rows], ac[rows], lp[rows], ad[rows], at[rows])):
            assert np.array_equal(got.numpy(), want)


@pytest.mark.parametrize("S,A,n,mb", [(17, 6, 5000, 1237), (3, 1, 4099, 4096), (40, 9, 3001, 1000), (1, 1, 700, 333), (100, 25, 600, 599)])
def test_packed_gather_is_bit_exact(L, S, A, n, mb):
    """Row-packed mirror + gather from it (csrc/buffer.cu) against numpy fancy indexing: same five output arrays as the SoA
    gather, bit for bit, for rows of 1 .. 16 sectors, odd minibatches and the (offset + i) % limit wrap-around."""
    rng = np.random.default_rng(n + mb)
    st, ac = rng.standard_normal((n, S)).astype(f32), rng.standard_normal((n, A)).astype(f32)
    lp, ad, at = (rng.standard_normal(n).astype(f32) for _ in range(3))
    idx = rng.permutation(n).astype(i32)
    d = [b200.dev(x) for x in (idx, st, ac, lp, ad, at)]
    pw = L.ppo_b200_packed_row_floats(S, A)
    assert pw % 8 == 0 and pw >= S + A + 3
    packed = b200.dev_empty((n, pw))
    L.ppo_b200_pack_rows(packed.ptr, n, S, A, *[x.ptr for x in d[1:]])
    pk = packed.numpy()
    assert np.array_equal(pk[:, :S], st) and np.array_equal(pk[:, S:S + A], ac) and np.array_equal(pk[:, S + A + 2], at)
    for offset in [0, n - mb // 2 - 1]:
        o = [b200.dev_empty((mb, S)), b200.dev_empty((mb, A)), b200.dev_empty(mb), b200.dev_empty(mb), b200.dev_empty(mb)]
        L.ppo_b200_gather_packed(d[0].ptr, offset, n, mb, S, A, packed.ptr, *[x.ptr for x in o])
        rows = idx[(offset + np.arange(mb)) % n]
        for got, want in zip(o, (st[rows], ac[rows], lp[rows], ad[rows], at[rows])):
            assert np.array_equal(got.numpy(), want)


def test_empty_inputs_are_noops(L):
    """n = 0 / batch 0 calls must return without launching on garbage (the reference would index out of bounds)."""
    z = b200.dev_empty(4)
    L.ppo_b200_gae(z.ptr, z.ptr, z.ptr, z.ptr, z.ptr, 0, 0.99, 0.95, z.ptr, z.ptr, 1, z.ptr)
    L.ppo_b200_gather(z.ptr, 0, 1, 0, 3, 1, *[z.ptr] * 10)
    L.ppo_b200_adam_flat(z.ptr, z.ptr, z.ptr, z.ptr, 0, 3e-4, 0.9, 0.999, 1)
    L.ppo_b200_sync()


@pytest.mark.parametrize("n", [1, 2, 1000, 819200])
def test_device_permutation_is_a_permutation(L, n):
    d = b200.dev_empty(n, i32)
    L.ppo_b200_permutation(d.ptr, n, 1234, 0)
    p0 = d.numpy()
    L.ppo_b200_permutation(d.ptr, n, 1234, 1)
    p1 = d.numpy()
    assert np.array_equal(np.sort(p0), np.arange(n)) and np.array_equal(np.sort(p1), np.arange(n))
    if n >= 1000:
        assert (p0 != p1).mean() > 0.9 and (p0 != np.arange(n)).mean() > 0.9


# ------------------------------------------------------------------------------------------- MLP
def _mlp_case(L, sizes, acts, m, seed, params=None):
    cabi.srand(seed)
    nn = L.create_neural_network(cabi.int_array(sizes), cabi.cstr_array(acts), len(sizes))
    cabi.srand(seed)
    p = oracle.init_params(sizes)
    assert np.array_equal(b200.nn_get_params(L, nn), p)      # rand()-driven init is bit-exact
    if params is not None:
        p = params
        b200.nn_set_params(L, nn, p)
    return nn, p


@pytest.mark.parametrize("sizes,acts,m", [([3, 64, 64, 1], RELU3, 64), ([3, 64, 64, 1], ["tanh", "tanh", "none"], 4096),
                                          ([17, 256, 256, 6], RELU3, 1000), ([17, 256, 256, 1], RELU3, 257),
                                          ([1, 1], ["none"], 5), ([5, 8, 3], ["relu", "relu"], 7),
                                          ([17, 1024, 1024, 1024, 6], ["relu", "relu", "relu", "none"], 300),
                                          ([3, 128, 128, 1], RELU3, 1)])
def test_forward_backward_vs_oracle(L, sizes, acts, m):
    nn, p = _mlp_case(L, sizes, acts, m, seed=sum(sizes))
    rng = np.random.default_rng(m)
    x, g = rng.standard_normal((m, sizes[0])).astype(f32), rng.standard_normal((m, sizes[-1])).astype(f32)
    dx, dg = b200.dev(x), b200.dev(g)
    L.forward_propagation_cuda(nn, dx.fp(), m)
    y = b200.d2h(L, nn.contents.d_output, (m, sizes[-1]))
    L.backward_propagation_cuda(nn, dg.fp(), m)
    grads = b200.nn_get_device_grads(L, nn)
    y_o, cache = oracle.mlp_forward(p, sizes, acts, x)
    g_o = oracle.mlp_backward(p, sizes, acts, cache, g)
    assert nerr(y, y_o) < TOL
    # per-tensor comparison of the gradients
    o = 0
    for i in range(len(sizes) - 1):
        for cnt in (sizes[i] * sizes[i + 1], sizes[i + 1]):
            assert nerr(grads[o:o + cnt], g_o[o:o + cnt]) < TOL, (i, cnt)
            o += cnt
    # host-pointer twins agree with the device twins
    L.forward_propagation(nn, cabi.fptr(x), m)
    assert np.array_equal(np.ctypeslib.as_array(nn.contents.output, shape=(m, sizes[-1])), y)
    L.free_neural_network(nn)


def test_forward_backward_vs_reference_golden(L, golden):
    for tag in ["pend64", "cheetah32", "relu_out"]:
        sizes = [int(s) for s in golden[f"mlp_{tag}_sizes"]]
        acts = ["relu"] * (len(sizes) - 2) + ["relu" if golden[f"mlp_{tag}_relu_last"][0] else "none"]
        cabi.srand(int(golden[f"mlp_{tag}_seed"][0]))
        nn = L.create_neural_network(cabi.int_array(sizes), cabi.cstr_array(acts), len(sizes))
        assert np.array_equal(b200.nn_get_params(L, nn), golden[f"mlp_{tag}_params"])
        x, g = golden[f"mlp_{tag}_x"], golden[f"mlp_{tag}_g"]
        dx, dg = b200.dev(x), b200.dev(g)
        L.forward_propagation_cuda(nn, dx.fp(), x.shape[0])
        assert nerr(b200.d2h(L, nn.contents.d_output, golden[f"mlp_{tag}_y"].shape), golden[f"mlp_{tag}_y"]) < TOL
        L.backward_propagation_cuda(nn, dg.fp(), x.shape[0])
        assert nerr(b200.nn_get_device_grads(L, nn), golden[f"mlp_{tag}_grads"]) < TOL
        L.free_neural_network(nn)


def test_mat_mul_cuda_entry_points(L):
    rng = np.random.default_rng(0)
    m, n, l = 130, 70, 33
    x, w, b, g = (rng.standard_normal(s).astype(f32) for s in [(m, n), (l, n), (l,), (m, l)])
    dx, dw, db, dg, dout, dgx, dgw = b200.dev(x), b200.dev(w), b200.dev(b), b200.dev(g), b200.dev_empty((m, l)), b200.dev_empty((m, n)), b200.dev_empty((l, n))
    L.mat_mul_cuda(None, dout.fp(), dx.fp(), dw.fp(), db.fp(), m, n, l)
    L.mat_mul_backwards_cuda(None, dgx.fp(), dgw.fp(), dg.fp(), dx.fp(), dw.fp(), m, n, l)
    x64, w64, g64 = x.astype(np.float64), w.astype(np.float64), g.astype(np.float64)
    assert nerr(dout.numpy(), x64 @ w64.T + b) < TOL
    assert nerr(dgx.numpy(), g64 @ w64) < TOL and nerr(dgw.numpy(), g64.T @ x64) < TOL
    # activations
    y = rng.standard_normal((m, l)).astype(f32)
    dy = b200.dev(y)
    L.ReLU_cuda(dy.fp(), m, l)
    assert np.array_equal(dy.numpy(), np.maximum(y, 0))
    dgrad = b200.dev(g)
    L.ReLU_derivative_cuda(dy.fp(), dgrad.fp(), m, l)
    assert np.array_equal(dgrad.numpy(), np.where(np.maximum(y, 0) > 0, g, 0))


@pytest.mark.parametrize("shapes", [[(32768, 3, 64), (32768, 64, 1)], [(32768, 64, 1), (32768, 17, 64), (32768, 3, 64)],
                                    [(20000, 64, 6), (20000, 64, 6)]])
def test_mat_mul_backwards_cuda_narrow_layers_large_m(L, shapes):
    """Narrow layers at large m take the skinny row-split dW kernels (R > 1 staging + fold).  Their staging must not
    alias or free the caller's split-K slabs (round-1 advisor finding: both lived in one scratch slot); called back
    to back with growing and shrinking shapes, every gradient must match float64."""
    rng = np.random.default_rng(7)
    for (m, n, l) in shapes:
        x, w, g = (rng.standard_normal(s).astype(f32) for s in [(m, n), (l, n), (m, l)])
        dx, dw, dg, dgx, dgw = b200.dev(x), b200.dev(w), b200.dev(g), b200.dev_empty((m, n)), b200.dev_empty((l, n))
        L.mat_mul_backwards_cuda(None, dgx.fp(), dgw.fp(), dg.fp(), dx.fp(), dw.fp(), m, n, l)
        x64, w64, g64 = x.astype(np.float64), w.astype(np.float64), g.astype(np.float64)
        assert nerr(dgw.numpy(), g64.T @ x64) < TOL, (m, n, l)
        assert nerr(dgx.numpy(), g64 @ w64) < TOL, (m, n, l)
        for d in (dx, dw, dg, dgx, dgw):
            d.free()


@pytest.mark.parametrize("m,n,l", [(4096, 17, 256), (20000, 3, 64), (5000, 32, 1024), (4096, 256, 6), (20000, 1024, 1), (3000, 64, 8), (2050, 24, 264)])
def test_mat_mul_cuda_narrow_layers_forward(L, m, n, l):
    """Forward of the narrow layers of a wide net at large m (csrc/narrow.cu: first layer with <= 32 inputs, heads with <= 8
    outputs) through the reference entry point, against float64."""
    rng = np.random.default_rng(m + n + l)
    x, w, b = (rng.standard_normal(s).astype(f32) for s in [(m, n), (l, n), (l,)])
    dx, dw, db, dout = b200.dev(x), b200.dev(w), b200.dev(b), b200.dev_empty((m, l))
    L.mat_mul_cuda(None, dout.fp(), dx.fp(), dw.fp(), db.fp(), m, n, l)
    assert nerr(dout.numpy(), x.astype(np.float64) @ w.astype(np.float64).T + b) < TOL
    for d in (dx, dw, db, dout):
        d.free()


@pytest.mark.parametrize("sizes,act,m", [([17, 256, 256, 6], "relu", 4096), ([17, 256, 256, 1], "tanh", 3000), ([3, 128, 1024, 8], "tanh", 2048)])
def test_wide_net_with_narrow_ends_vs_oracle(L, sizes, act, m):
    """Nets with a low-dimensional input and a <= 8 wide head above wide layers at m >= 1024: the first layer and the head run
    in the streaming kernels of csrc/narrow.cu (the head's dW and dX in ONE pass over its input); outputs and every gradient
    tensor against the oracle at the fp32 tolerance."""
    acts = [act] * (len(sizes) - 2) + ["none"]
    cabi.srand(9)
    nn = L.create_neural_network(cabi.int_array(sizes), cabi.cstr_array(acts), len(sizes))
    p = b200.nn_get_params(L, nn)
    rng = np.random.default_rng(m)
    x, g = rng.standard_normal((m, sizes[0])).astype(f32), rng.standard_normal((m, sizes[-1])).astype(f32)
    dx, dg = b200.dev(x), b200.dev(g)
    L.forward_propagation_cuda(nn, dx.fp(), m)
    y = b200.d2h(L, nn.contents.d_output, (m, sizes[-1]))
    L.backward_propagation_cuda(nn, dg.fp(), m)
    grads = b200.nn_get_device_grads(L, nn)
    y_o, cache = oracle.mlp_forward(p, sizes, acts, x)
    g_o = oracle.mlp_backward(p, sizes, acts, cache, g)
    assert nerr(y, y_o) < TOL
    o = 0
    for i in range(len(sizes) - 1):
        for cnt in (sizes[i] * sizes[i + 1], sizes[i + 1]):
            assert nerr(grads[o:o + cnt], g_o[o:o + cnt]) < TOL, (i, cnt)
            o += cnt
    L.free_neural_network(nn)
    dx.free(); dg.free()


# ------------------------------------------------------------------------------------------ policy
def _policy(L, sizes, params, log_std):
    pol = L.create_gaussian_policy(cabi.int_array(sizes), cabi.cstr_array(RELU3), len(sizes), 1.0)
    b200.nn_set_params(L, pol.contents.mu, params)
    A = sizes[-1]
    np.ctypeslib.as_array(pol.contents.log_std, shape=(A,))[:] = log_std
    L.ppo_b200_h2d(C.cast(pol.contents.d_log_std, C.c_void_p), log_std.ctypes.data, 4 * A)
    return pol


@pytest.mark.parametrize("pre", ["pol_", "pol6_"])
def test_policy_stage_vs_reference_golden(L, golden, pre):
    g = golden
    sizes = [int(s) for s in g[pre + "sizes"]]
    A = sizes[-1]
    m = g[pre + "state"].shape[0]
    pol = _policy(L, sizes, g[pre + "params"], g[pre + "log_std"])
    ds, da, dlp = b200.dev(g[pre + "state"]), b200.dev(g[pre + "action"]), b200.dev_empty(m)
    L.compute_log_prob_cuda(pol, dlp.fp(), ds.fp(), da.fp(), m)
    lp = dlp.numpy()
    assert nerr(b200.d2h(L, pol.contents.mu.contents.d_output, (m, A)), g[pre + "mu"]) < TOL
    assert nerr(lp, g[pre + "logprob"]) < TOL
    ent = L.compute_entropy_cuda(pol)
    assert abs(ent - float(g[pre + "entropy"])) < 1e-6
    ec = float(g["pol_ent_coeff"][0]) if A == 1 else 0.0
    dgl, dadv, dold = b200.dev_empty(m), b200.dev(g[pre + "adv"]), b200.dev(g[pre + "lp_old"])
    ge = C.c_float()
    loss = L.policy_loss_and_grad_cuda(dgl.fp(), C.byref(ge), dadv.fp(), dlp.fp(), dold.fp(), ent, ec, 0.2, m)
    assert abs(loss - float(g[pre + "loss"])) < 1e-5 * max(1.0, abs(float(g[pre + "loss"])))
    gl = dgl.numpy()
    # the clip decisions (integer work) must match exactly: same zero pattern
    assert np.array_equal(gl == 0, g[pre + "grad_logprob"] == 0)
    assert nerr(gl, g[pre + "grad_logprob"]) < TOL and ge.value == -ec
    dgmu, dgls = b200.dev_empty((m, A)), b200.dev_empty(A)
    L.log_prob_backwards_cuda(pol, dgl.fp(), dgmu.fp(), dgls.fp(), m)
    # stage-level: the oracle gets exactly the inputs the kernel got (the device mu and grad_logprob)
    mu_dev = b200.d2h(L, pol.contents.mu.contents.d_output, (m, A))
    gmu_o, gls_o = oracle.log_prob_backwards(mu_dev, g[pre + "log_std"], g[pre + "action"], gl, ref_index=False)
    assert nerr(dgmu.numpy(), gmu_o) < TOL and nerr(dgls.numpy(), gls_o) < TOL
    if A == 1:
        assert nerr(dgmu.numpy(), g["pol_grad_mu"]) < TOL and nerr(dgls.numpy(), g["pol_grad_log_std"]) < TOL
        L.backward_propagation_cuda(pol.contents.mu, dgmu.fp(), m)
        assert nerr(b200.nn_get_device_grads(L, pol.contents.mu), g["pol_grads"]) < TOL
    L.free_gaussian_policy(pol)


def test_mse_vs_reference_golden(L, golden):
    y, yt = b200.dev(golden["mse_y"]), b200.dev(golden["mse_yt"])
    n = golden["mse_y"].size
    loss = L.mean_squared_error_cuda(y.fp(), yt.fp(), n, 1)
    assert abs(loss - golden["mse_loss"][0]) < 1e-6 * golden["mse_loss"][0] + 1e-7
    g = b200.dev_empty(n)
    L.mean_squared_error_derivative_cuda(g.fp(), y.fp(), yt.fp(), n, 1)
    assert np.array_equal(g.numpy(), golden["mse_grad"])      # mul/sub/div only: bit-exact


def test_sample_action_consumes_reference_rand_stream(L, golden):
    """sample_action with mu == 0, std == 1: action == Box-Muller noise of the reference's rand() stream."""
    pol = _policy(L, [1, 1], np.zeros(2, f32), np.zeros(1, f32)) if False else None
    pol = L.create_gaussian_policy(cabi.int_array([1, 1]), cabi.cstr_array(["none"]), 2, 1.0)
    b200.nn_set_params(L, pol.contents.mu, np.zeros(2, f32))
    cabi.srand(int(golden["noise_seed"][0]))
    s, a, lp = np.zeros(1, f32), np.zeros(1, f32), np.zeros(1, f32)
    acts, lps = [], []
    for _ in range(len(golden["noise_actions"])):
        L.sample_action(pol, cabi.fptr(s), cabi.fptr(a), cabi.fptr(lp), 1)
        acts.append(a[0]); lps.append(lp[0])
    assert np.max(np.abs(np.array(acts) - golden["noise_actions"])) < 2e-6
    assert np.max(np.abs(np.array(lps) - golden["noise_logprob"])) < 1e-5
    cabi.srand(int(golden["noise_seed"][0]))
    [cabi.rand() for _ in range(2 * len(acts))]
    nxt = cabi.rand()
    cabi.srand(int(golden["noise_seed"][0]))
    for _ in range(len(acts)):
        L.sample_action(pol, cabi.fptr(s), cabi.fptr(a), cabi.fptr(lp), 1)
    assert cabi.rand() == nxt      # exactly two draws per step
    L.free_gaussian_policy(pol)


def test_pendulum_step_kernel_vs_oracle(L):
    rng = np.random.default_rng(2)
    n = 1000
    th, thd = rng.uniform(-10, 10, n), rng.uniform(-8, 8, n)
    act = rng.uniform(-3, 3, n).astype(f32)
    dth, dthd, dact = b200.dev(th), b200.dev(thd), b200.dev(act)
    dobs, drew = b200.dev_empty((n, 3)), b200.dev_empty(n)
    L.ppo_b200_pendulum_step(dth.ptr, dthd.ptr, dact.ptr, dobs.ptr, drew.ptr, n)
    th1, thd1, obs, rew = dth.numpy(), dthd.numpy(), dobs.numpy(), drew.numpy()
    for i in range(0, n, 7):
        t, td, o, r = oracle.pendulum_step(th[i], thd[i], act[i])
        assert abs(th1[i] - t) < 1e-12 and abs(thd1[i] - td) < 1e-12
        assert np.max(np.abs(obs[i] - o)) < 1e-6 and abs(rew[i] - r) < 1e-5 * max(1, abs(r))
