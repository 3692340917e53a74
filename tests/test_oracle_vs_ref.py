"""Live re-check of the oracle against the compiled reference (oracle/_ref/libppo_ref.so), on fresh
random inputs each parametrisation.  Skipped when the .so is absent (it is git-ignored; it exists in
the build container and travels to the GPU box with the gpurun snapshot).  CPU only."""
import os

import numpy as np
import pytest

import cabi
import oracle

pytestmark = pytest.mark.skipif(not os.path.exists(cabi.REF_SO), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def R():
    import refdrive
    return refdrive.Ref()


@pytest.mark.parametrize("seed,sizes,acts,m", [(1, [3, 64, 64, 1], ["relu", "relu", "none"], 33),
                                               (2, [17, 24, 24, 6], ["relu", "none", "none"], 20),
                                               (3, [4, 9, 2], ["none", "relu"], 1)])
def test_mlp_live(R, seed, sizes, acts, m):
    cabi.srand(seed)
    nn = R.create_nn(sizes, acts)
    cabi.srand(seed)
    p = oracle.init_params(sizes)
    assert np.array_equal(p, R.nn_get_params(nn))
    rng = np.random.default_rng(seed)
    x, g = rng.standard_normal((m, sizes[0])).astype(np.float32), rng.standard_normal((m, sizes[-1])).astype(np.float32)
    y, cache = oracle.mlp_forward(p, sizes, acts, x)
    assert np.array_equal(y, R.forward(nn, x))
    assert np.array_equal(oracle.mlp_backward(p, sizes, acts, cache, g), R.backward(nn, g))


@pytest.mark.parametrize("T,N", [(1, 1), (200, 15), (513, 3), (1000, 40)])
def test_gae_live(R, T, N):
    rng = np.random.default_rng(T * 1000 + N)
    n = T * N
    r, v, vn = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    term = (rng.random(n) < 0.01).astype(np.uint8)
    trunc = np.zeros(n, np.uint8)
    trunc[T - 1::T] = 1
    adv, tgt = R.gae(r, v, vn, term, trunc, 0.99, 0.95)
    raw, tgt_o, norm, mean, std = oracle.gae(r, v, vn, term, trunc, 0.99, 0.95)
    assert np.array_equal(norm, adv) and np.array_equal(tgt_o, tgt)


def test_whole_path_live(R):
    res = R.train_toy(5, 16, 202, 202, 64, 1, 2)
    cabi.srand(5)
    T = oracle.Trainer([1, 16, 16, 1], ["relu", "relu", "none"], batch_size=64, n_epochs_policy=1, n_epochs_value=2)
    b = T.make_buffer(202)
    T.collect(b, 202, 0)
    T.update(b)
    for k in ["mu", "v", "log_std"]:
        assert np.array_equal(getattr(T, k), res[k]), k
    assert np.array_equal(b["advantage"], res["advantage"])


@pytest.mark.parametrize("blas", [False, True], ids=["naive-cblas", "openblas"])
def test_pendulum_training_live(R, blas):
    """The unmodified reference's train_ppo_epoch on the C Pendulum (oracle hooks behind include/env.h) against the oracle
    port, same srand: bit-exact with the sequential-k cblas shim; with the real OpenBLAS (its own summation order) the
    weights agree to fp32 rounding - this is the reference build bench.py's C1 line times."""
    import ctypes as C
    lib = cabi.load_ref_blas() if blas else R.lib
    if lib is None:
        pytest.skip("bundled OpenBLAS not loadable")
    sizes, acts, cap, mb = [3, 64, 64, 1], ["relu", "relu", "none"], 600, 64
    env = cabi.oracle_pendulum_env(oracle.lib())
    cabi.srand(9)
    ppo = lib.create_ppo(cabi.cstr_array(acts), cabi.int_array(sizes), 4, cap, C.c_float(3e-4), C.c_float(3e-4),
                         C.c_float(0.95), C.c_float(0.2), C.c_float(0.0), C.c_float(1.0), False)
    lib.train_ppo_epoch(ppo, C.byref(env), cap, mb, 1, 2)
    after = cabi.rand()
    import refdrive
    mu, v = refdrive.Ref.nn_get_params(ppo.contents.policy.contents.mu), refdrive.Ref.nn_get_params(ppo.contents.V)
    cabi.srand(9)
    T = oracle.Trainer(sizes, acts, batch_size=mb, n_epochs_policy=1, n_epochs_value=2)
    b = T.make_buffer(cap)
    T.collect(b, cap, 1)
    T.update(b)
    assert cabi.rand() == after
    rw = np.ctypeslib.as_array(ppo.contents.buffer.contents.reward_p, shape=(cap,))
    if blas:
        assert np.max(np.abs(rw - b["reward"])) < 1e-3 and np.max(np.abs(mu - T.mu)) < 1e-4 and np.max(np.abs(v - T.v)) < 1e-4
    else:
        assert np.array_equal(rw, b["reward"]) and np.array_equal(mu, T.mu) and np.array_equal(v, T.v)
