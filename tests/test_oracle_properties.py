"""Size-independent properties of the path, checked on the oracle (CPU).  The same properties are what the full-size GPU
tests lean on (tests/test_gpu_stages.py::test_gae_large_sampled_envs etc.): the scan is linear in (reward, v, v') for a
fixed done pattern, segments never leak across done flags, normalisation is idempotent up to rounding, the shuffle is a
permutation and the gather is plain indexing, Adam with a zero gradient only decays its moments."""
import numpy as np
import pytest

import cabi
import oracle

f32, u8 = np.float32, np.uint8


def _inputs(n, seed, p_done=0.01):
    rng = np.random.default_rng(seed)
    r, v, vn = (rng.standard_normal(n).astype(f32) for _ in range(3))
    term = (rng.random(n) < p_done).astype(u8)
    trunc = (rng.random(n) < p_done).astype(u8)
    trunc[-1] = 1                                    # the reference's buffers always end with a done flag (ppo.cu:70-74)
    return r, v, vn, term, trunc


@pytest.mark.parametrize("n", [1, 7, 513, 4096])
def test_gae_is_linear_for_a_fixed_done_pattern(n):
    r1, v1, vn1, term, trunc = _inputs(n, 1)
    r2, v2, vn2, _, _ = _inputs(n, 2)
    a1 = oracle.gae_f64(r1, v1, vn1, term, trunc, 0.99, 0.95)[0]
    a2 = oracle.gae_f64(r2, v2, vn2, term, trunc, 0.99, 0.95)[0]
    a12 = oracle.gae_f64((r1 + r2).astype(f32), (v1 + v2).astype(f32), (vn1 + vn2).astype(f32), term, trunc, 0.99, 0.95)[0]
    # float64 recursion on float32 inputs: (r1 + r2) is rounded to float32 once, nothing else differs
    assert np.max(np.abs(a12 - (a1 + a2))) < 1e-4


def test_gae_segments_do_not_leak_across_done_flags():
    n = 2000
    r, v, vn, term, trunc = _inputs(n, 3, p_done=0.02)
    base = oracle.gae(r, v, vn, term, trunc, 0.99, 0.95)[0]
    cut = int(np.flatnonzero(term | trunc)[3])       # everything after the 4th done flag is another episode
    r2, v2, vn2 = r.copy(), v.copy(), vn.copy()
    r2[cut + 1:] += 100.0
    v2[cut + 1:] -= 7.0
    vn2[cut + 1:] *= 3.0
    pert = oracle.gae(r2, v2, vn2, term, trunc, 0.99, 0.95)[0]
    assert np.array_equal(base[:cut + 1], pert[:cut + 1])


def test_all_done_reduces_gae_to_delta_and_termination_drops_the_bootstrap():
    n = 300
    r, v, vn, _, _ = _inputs(n, 4)
    term = np.zeros(n, u8); trunc = np.ones(n, u8)
    raw, tgt, _, _, _ = oracle.gae(r, v, vn, term, trunc, 0.99, 0.95)
    assert np.allclose(raw, r + f32(0.99) * vn - v, atol=1e-6)          # truncation keeps the bootstrap (ppo.cu:340-342)
    assert np.allclose(tgt, v + raw, atol=1e-6)
    term = np.ones(n, u8)
    raw_t = oracle.gae(r, v, vn, term, trunc, 0.99, 0.95)[0]
    assert np.allclose(raw_t, r - v, atol=1e-6)                            # termination zeroes it


def test_normalised_advantages_have_zero_mean_unit_std_and_normalising_twice_is_a_fixed_point():
    n = 3000
    r, v, vn, term, trunc = _inputs(n, 5)
    _, _, norm, mean, std = oracle.gae(r, v, vn, term, trunc, 0.99, 0.95)
    assert abs(float(norm.astype(np.float64).mean())) < 1e-5 and abs(float(norm.astype(np.float64).std()) - 1) < 1e-4
    again = (norm - norm.mean()) / (norm.std() + 1e-8)
    assert np.max(np.abs(again - norm)) < 1e-4


@pytest.mark.parametrize("n", [1, 2, 64, 3000])
def test_shuffle_is_a_permutation_and_gather_is_indexing(n):
    cabi.srand(n)
    idx = oracle.shuffle(n)
    assert np.array_equal(np.sort(idx), np.arange(n))
    rng = np.random.default_rng(n)
    S, A, mb = 5, 2, max(1, n // 3)
    st, ac = rng.standard_normal((n, S)).astype(f32), rng.standard_normal((n, A)).astype(f32)
    lp, ad, at = (rng.standard_normal(n).astype(f32) for _ in range(3))
    for k in range(n // mb):
        rows = idx[(k * mb + np.arange(mb)) % n]
        got = oracle.get_batch(idx, k, mb, st, ac, lp, ad, at)
        for g, want in zip(got, (st[rows], ac[rows], lp[rows], ad[rows], at[rows])):
            assert np.array_equal(g, want)


def test_adam_with_zero_gradient_keeps_weights_and_first_step_moves_by_lr():
    w = np.linspace(-1, 1, 101).astype(f32)
    w0 = w.copy()
    m, v = np.zeros_like(w), np.zeros_like(w)
    t = oracle.adam(w, np.zeros_like(w), m, v, 3e-4, 0)
    assert t == 1 and np.array_equal(w, w0) and not m.any() and not v.any()
    g = np.where(np.arange(101) % 2 == 0, 0.5, -2.0).astype(f32)
    oracle.adam(w, g, m, v, 3e-4, t)
    # bias-corrected first real step: |dw| = lr * |m_hat| / (sqrt(v_hat) + eps) = lr up to the second-step bias terms
    step = np.abs(w - w0)
    assert np.all(step > 0.5 * 3e-4) and np.all(step < 2.5 * 3e-4) and np.array_equal(np.sign(w0 - w), np.sign(g))
