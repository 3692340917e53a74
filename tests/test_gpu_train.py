"""Whole-path parity and behaviour of the training path through the reference API (-m gpu)."""
import ctypes as C
import os

import numpy as np
import pytest

import b200
import cabi
import oracle
from conftest import nerr

pytestmark = pytest.mark.gpu
f32, u8, i32 = np.float32, np.uint8, np.int32
RELU3 = ["relu", "relu", "none"]


@pytest.fixture(scope="module")
def L():
    lib = b200.lib()
    assert lib.ppo_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return lib


def make_ppo(L, sizes, acts, capacity, lr=3e-4, lam=0.95, eps=0.2, ent=0.0, init_std=1.0):
    return L.create_ppo(cabi.cstr_array(acts), cabi.int_array(sizes), len(sizes), capacity, lr, lr, lam, eps, ent, init_std, True)


def fill_host_buffer(ppo, b):
    buf = ppo.contents.buffer.contents
    n = b["reward"].shape[0]
    S, A = b["state"].shape[1], b["action"].shape[1]
    np.ctypeslib.as_array(buf.h_state_p, shape=(n, S))[:] = b["state"]
    np.ctypeslib.as_array(buf.h_next_state_p, shape=(n, S))[:] = b["next_state"]
    np.ctypeslib.as_array(buf.h_action_p, shape=(n, A))[:] = b["action"]
    np.ctypeslib.as_array(buf.h_reward_p, shape=(n,))[:] = b["reward"]
    np.ctypeslib.as_array(buf.h_logprob_p, shape=(n,))[:] = b["logprob"]
    np.ctypeslib.as_array(buf.h_terminated_p, shape=(n,))[:] = b["terminated"].astype(bool)
    np.ctypeslib.as_array(buf.h_truncated_p, shape=(n,))[:] = b["truncated"].astype(bool)


def host_field(ppo, name, shape, dtype=f32):
    return np.ctypeslib.as_array(getattr(ppo.contents.buffer.contents, "h_" + name + "_p"), shape=shape).astype(dtype)


def synthetic_buffer(T, rng, sizes, acts, n):
    """SURVEY.md §8d C3 generator (scaled): N(0,1) states/actions/rewards, logprob_old near the
    initial policy's log-prob, Bernoulli terminations, forced done at the end."""
    S, A = sizes[0], sizes[-1]
    b = T.make_buffer(n)
    b["state"][:] = rng.standard_normal((n, S))
    b["next_state"][:] = rng.standard_normal((n, S))
    b["action"][:] = rng.standard_normal((n, A))
    b["reward"][:] = rng.standard_normal(n)
    mu, _ = oracle.mlp_forward(T.mu, T.sizes_mu, acts, b["state"])
    b["logprob"][:] = oracle.log_prob(mu, T.log_std, b["action"]) + 0.1 * rng.standard_normal(n)
    b["terminated"][:] = rng.random(n) < 0.01
    b["truncated"][199::200] = 1
    b["truncated"][-1] = 1
    return b


@pytest.mark.parametrize("path", [1, 0], ids=["fused", "layerwise"])
@pytest.mark.parametrize("sizes,acts,n,mb,npol,nval", [
    ([1, 8, 8, 1], RELU3, 302, 32, 2, 3),
    ([3, 64, 64, 1], RELU3, 3000, 64, 1, 2),                   # the reference's Pendulum shape
    ([3, 64, 64, 1], ["tanh", "tanh", "none"], 2048, 256, 2, 2),  # [EXT] tanh
    ([3, 128, 128, 1], RELU3, 1500, 100, 1, 1),                # the reference's default width (main.c:20)
    ([17, 32, 32, 6], RELU3, 4096, 512, 2, 2),                 # HalfCheetah-shaped, A = 6
    ([17, 256, 256, 6], RELU3, 2048, 1024, 1, 1),              # C3 widths (layer-wise kernels on both paths)
])
def test_update_phase_matches_oracle(L, sizes, acts, n, mb, npol, nval, path):
    L.ppo_b200_set_kernel_path(path)
    """GAE + value epochs + policy epochs on an identical buffer, identical rand() stream:
    permutation-driven minibatches are the same, post-Adam weights agree to 1e-5 (norm-wise)."""
    seed = 100 + n
    cabi.srand(seed)
    ppo = make_ppo(L, sizes, acts, n)
    cabi.srand(seed)
    T = oracle.Trainer(sizes, acts, batch_size=mb, n_epochs_policy=npol, n_epochs_value=nval, ref_index=False)
    assert np.array_equal(b200.nn_get_params(L, ppo.contents.policy.contents.mu), T.mu)
    assert np.array_equal(b200.nn_get_params(L, ppo.contents.V), T.v)
    b = synthetic_buffer(T, np.random.default_rng(seed), sizes, acts, n)
    fill_host_buffer(ppo, b)
    cabi.srand(seed + 1)
    L.ppo_b200_update(ppo, 0.99, mb, npol, nval)
    after_gpu = cabi.rand()
    cabi.srand(seed + 1)
    perms, losses = T.update(b, log_perms=True, log_losses=True)
    assert cabi.rand() == after_gpu                    # same number of rand() draws -> same permutations
    nb = n // mb
    adv = host_field(ppo, "advantage", (n,))
    assert np.max(np.abs(adv - b["advantage"])) < 2e-5          # normalised advantages (float ref at small B)
    assert nerr(host_field(ppo, "adv_target", (n,)), b["adv_target"]) < 1e-5
    # Post-Adam weights.  Adam's early steps move every weight by ~lr*sign(g): an element whose
    # gradient sits inside fp32 summation noise of zero moves by a different fraction of lr in ANY two
    # implementations (the oracle vs float64 included).  So: essentially all weights within 1e-5
    # (norm-wise), and the worst element bounded by a few % of one lr step (1e-4 norm-wise).
    for got, want in ((b200.nn_get_params(L, ppo.contents.V, sync=False), T.v),
                      (b200.nn_get_params(L, ppo.contents.policy.contents.mu, sync=False), T.mu)):
        scale = np.max(np.abs(want))
        assert nerr(got, want) < 1e-4
        assert np.mean(np.abs(got - want) > 1e-5 * scale) < 0.01
    ls = np.ctypeslib.as_array(ppo.contents.policy.contents.log_std, shape=(sizes[-1],))
    assert np.max(np.abs(ls - T.log_std)) < 1e-6
    assert ppo.contents.adam_V.contents.time_step == nval * nb == T.model.t_v
    assert ppo.contents.adam_policy.contents.time_step == npol * nb == T.model.t_mu
    assert abs(L.ppo_b200_last_value_loss(ppo) - losses[:nval * nb].mean()) < 1e-4 * abs(losses[:nval * nb].mean()) + 1e-6
    assert abs(L.ppo_b200_last_policy_loss(ppo) - losses[nval * nb:].mean()) < 1e-4 + 1e-4 * abs(losses[nval * nb:].mean())
    L.ppo_b200_set_kernel_path(-1)
    L.free_ppo(ppo)


@pytest.mark.parametrize("sizes,acts,n,mb", [([3, 64, 64, 1], ["tanh", "tanh", "none"], 4096, 1024),
                                             ([17, 32, 32, 6], RELU3, 2048, 512),
                                             ([3, 128, 128, 1], RELU3, 2048, 512),
                                             ([17, 256, 256, 6], RELU3, 4096, 2048)])
def test_update_is_bitwise_reproducible(L, sizes, acts, n, mb):
    """Every reduction in the update path runs in a fixed order (slab sums, split-K, loss heads; the reference uses float
    atomics, src/policy.cu:157), so two runs from the same state must agree BIT FOR BIT.  This doubles as the race
    detector for the hand-synchronised shared-memory kernels (compute-sanitizer is not available on the GPU pool)."""
    outs = []
    for _ in range(3):
        cabi.srand(31)
        ppo = make_ppo(L, sizes, acts, n)
        T = oracle.Trainer(sizes, acts, batch_size=mb, n_epochs_policy=2, n_epochs_value=2, init=False)
        T.mu[:] = b200.nn_get_params(L, ppo.contents.policy.contents.mu)
        b = synthetic_buffer(T, np.random.default_rng(5), sizes, acts, n)
        fill_host_buffer(ppo, b)
        cabi.srand(32)
        L.ppo_b200_update(ppo, 0.99, mb, 2, 2)
        outs.append((b200.nn_get_params(L, ppo.contents.V, sync=False).copy(),
                     b200.nn_get_params(L, ppo.contents.policy.contents.mu, sync=False).copy(),
                     host_field(ppo, "advantage", (n,)).copy(), L.ppo_b200_last_policy_loss(ppo), L.ppo_b200_last_value_loss(ppo)))
        L.free_ppo(ppo)
    for o in outs[1:]:
        for a, ref in zip(o[:3], outs[0][:3]):
            assert np.array_equal(a, ref)
        assert o[3] == outs[0][3] and o[4] == outs[0][4]


def test_train_ppo_epoch_toy_env_matches_oracle(L):
    """Reference entry point, opaque host env (toy env of src/env.c): rollout on the GPU one step at a
    time from the reference's rand() stream, then the update.  One iteration, so GPU-vs-libm ulps in
    the sampled actions stay ulps."""
    seed, cap, mb = 21, 302, 32
    cabi.srand(seed)
    env = L.create_simple_env(0, seed)
    ppo = make_ppo(L, [1, 8, 8, 1], RELU3, cap)
    L.train_ppo_epoch(ppo, env, cap, mb, 2, 3)
    after = cabi.rand()
    cabi.srand(seed)
    T = oracle.Trainer([1, 8, 8, 1], RELU3, batch_size=mb, n_epochs_policy=2, n_epochs_value=3)
    b = T.make_buffer(cap)
    T.collect(b, cap, 0)
    T.update(b)
    assert cabi.rand() == after
    assert np.array_equal(host_field(ppo, "terminated", (cap,), u8), b["terminated"])     # done masks: bit-exact
    assert np.array_equal(host_field(ppo, "truncated", (cap,), u8), b["truncated"])
    assert np.array_equal(host_field(ppo, "reward", (cap,)), b["reward"])
    assert np.max(np.abs(host_field(ppo, "action", (cap, 1)) - b["action"])) < 1e-5
    assert np.max(np.abs(host_field(ppo, "logprob", (cap,)) - b["logprob"])) < 1e-5
    assert nerr(b200.nn_get_params(L, ppo.contents.V, sync=False), T.v) < 1e-4
    assert nerr(b200.nn_get_params(L, ppo.contents.policy.contents.mu, sync=False), T.mu) < 1e-4
    L.eval_ppo(ppo, env, cap)       # prints J / R / Episodes like the reference
    L.free_ppo(ppo)
    env.contents.free_env()


def test_host_pendulum_env_semantics(L):
    env = L.create_pendulum_env(0, 3)
    e = env.contents
    assert (e.state_size, e.action_size, e.horizon) == (3, 1, 200) and abs(e.gamma - 0.99) < 1e-7
    obs, nobs, r = np.zeros(3, f32), np.zeros(3, f32), C.c_float()
    term, trunc = C.c_bool(), C.c_bool()
    e.reset_env(cabi.fptr(obs))
    assert abs(obs[0] ** 2 + obs[1] ** 2 - 1) < 1e-6 and abs(obs[2]) <= 1
    th, thd = np.arctan2(float(obs[1]), float(obs[0])), float(obs[2])
    a = np.array([1.5], f32)
    for t in range(200):
        e.step_env(cabi.fptr(a), cabi.fptr(nobs), C.byref(r), C.byref(term), C.byref(trunc), 1)
        th, thd, o, rr = oracle.pendulum_step(th, thd, 1.5)
        assert np.max(np.abs(nobs - o)) < 1e-5 and abs(r.value - rr) < 1e-4
        assert not term.value and trunc.value == (t == 199)


@pytest.mark.parametrize("H,acts", [(64, ["tanh", "tanh", "none"]), (64, RELU3), (128, RELU3), (128, ["tanh", "tanh", "none"]), (32, RELU3)],
                         ids=["64-tanh(bench config)", "64-relu", "128-relu", "128-tanh", "32-relu"])
def test_device_rollout_is_self_consistent(L, H, acts):
    """Fused device rollout (4096 envs x 16 steps; hidden widths 64 / 128 (two 64-unit blocks) / 32): every stored row must be reproducible by the
    oracle from the stored state/action: log-prob under the policy, reward and next state from the
    Pendulum definition, reference bookkeeping of next_state -> state and the forced last-step flag."""
    n_envs, T = 4096, 16
    cabi.srand(5)
    env = L.create_pendulum_env_cuda(n_envs, 7)
    assert L.ppo_b200_env_is_device(env) == 1 and L.ppo_b200_env_num_envs(env) == n_envs
    sizes = [3, H, H, 1]
    ppo = make_ppo(L, sizes, acts, n_envs * T)
    L.collect_trajectories(ppo.contents.buffer, env, ppo.contents.policy, n_envs * T)
    L.ppo_b200_sync_host(ppo)
    n = n_envs * T
    st, ns = host_field(ppo, "state", (n_envs, T, 3)), host_field(ppo, "next_state", (n_envs, T, 3))
    ac, rw, lp = host_field(ppo, "action", (n_envs, T)), host_field(ppo, "reward", (n_envs, T)), host_field(ppo, "logprob", (n_envs, T))
    term, trunc = host_field(ppo, "terminated", (n_envs, T), u8), host_field(ppo, "truncated", (n_envs, T), u8)
    assert not term.any() and trunc[:, -1].all() and not trunc[:, :-1].any()
    assert np.array_equal(st[:, 1:], ns[:, :-1])                      # src/ppo.cu:68
    assert np.max(np.abs(st[..., 0] ** 2 + st[..., 1] ** 2 - 1)) < 1e-5
    p = b200.nn_get_params(L, ppo.contents.policy.contents.mu)
    mu, _ = oracle.mlp_forward(p, sizes, acts, st.reshape(n, 3))
    lp_o = oracle.log_prob(mu, np.zeros(1, f32), ac.reshape(n, 1))
    assert np.max(np.abs(lp.ravel() - lp_o)) < 2e-5
    z = (ac.ravel() - mu.ravel())                                      # std = 1 -> noise
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02           # Box-Muller sanity
    th, thd = np.arctan2(st[..., 1].astype(np.float64), st[..., 0].astype(np.float64)), st[..., 2].astype(np.float64)
    u = np.clip(ac.astype(np.float64), -2, 2)
    cost = th ** 2 + 0.1 * thd ** 2 + 0.001 * u ** 2
    assert np.max(np.abs(rw + cost)) < 2e-4
    nthd = np.clip(thd + (15 * np.sin(th) + 3 * u) * 0.05, -8, 8)
    assert np.max(np.abs(ns[..., 2] - nthd)) < 1e-4
    # initial states: theta ~ U(-pi, pi), theta_dot ~ U(-1, 1), different across envs and rollouts
    assert abs(th[:, 0].mean()) < 0.15 and abs(th[:, 0].std() - np.pi / np.sqrt(3)) < 0.1
    assert np.max(np.abs(thd[:, 0])) <= 1.0 + 1e-6
    first = st[:, 0].copy()
    L.collect_trajectories(ppo.contents.buffer, env, ppo.contents.policy, n)
    L.ppo_b200_sync_host(ppo)
    assert not np.array_equal(host_field(ppo, "state", (n_envs, T, 3))[:, 0], first)
    L.free_ppo(ppo)
    env.contents.free_env()


def test_obs_normalisation_running_statistics(L):
    """[EXT] Welford running observation normalisation (no reference counterpart; merge formula of
    include/welford_var.h:33-40).  Rollout 1 runs with identity statistics, so its stored states ARE the raw
    observations: the running mean/std must equal numpy's over them.  Rollout 2 stores
    (raw - mean1) / (std1 + 1e-8); un-normalising it and pooling both rollouts must give the merged statistics."""
    n_envs, T = 512, 50
    cabi.srand(3)
    env = L.create_pendulum_env_cuda(n_envs, 11)
    ppo = make_ppo(L, [3, 64, 64, 1], RELU3, n_envs * T)
    L.ppo_b200_set_obs_norm(ppo, 1)
    mean, std, cnt = np.zeros(3, f32), np.zeros(3, f32), C.c_double()
    n = n_envs * T
    L.collect_trajectories(ppo.contents.buffer, env, ppo.contents.policy, n)
    L.ppo_b200_sync_host(ppo)
    raw1 = host_field(ppo, "state", (n, 3)).astype(np.float64)
    L.ppo_b200_get_obs_norm(env, mean.ctypes.data, std.ctypes.data, C.byref(cnt))
    assert cnt.value == n
    assert np.max(np.abs(mean - raw1.mean(0))) < 1e-5 and np.max(np.abs(std - raw1.std(0))) < 1e-5
    m1, s1 = mean.copy(), std.copy()
    L.collect_trajectories(ppo.contents.buffer, env, ppo.contents.policy, n)
    L.ppo_b200_sync_host(ppo)
    norm2 = host_field(ppo, "state", (n, 3)).astype(np.float64)
    assert abs(norm2[:, 2].std() - 1) < 0.5          # roughly unit scale now (theta_dot spreads out as episodes go on)
    raw2 = norm2 * (s1.astype(np.float64) + 1e-8) + m1
    assert np.max(np.abs(raw2[:, 0] ** 2 + raw2[:, 1] ** 2 - 1)) < 1e-4   # un-normalised cos/sin are on the unit circle
    L.ppo_b200_get_obs_norm(env, mean.ctypes.data, std.ctypes.data, C.byref(cnt))
    both = np.concatenate([raw1, raw2])
    assert cnt.value == 2 * n
    assert np.max(np.abs(mean - both.mean(0))) < 1e-4 and np.max(np.abs(std - both.std(0))) < 1e-4
    # next_state rows use the same statistics as the state rows of the same rollout (src/ppo.cu:68 carry-over)
    st, ns = host_field(ppo, "state", (n_envs, T, 3)), host_field(ppo, "next_state", (n_envs, T, 3))
    assert np.array_equal(st[:, 1:], ns[:, :-1])
    L.free_ppo(ppo)
    env.contents.free_env()


@pytest.mark.parametrize("acts", [["tanh", "tanh", "none"], RELU3], ids=["tanh(bench config)", "relu"])
def test_pendulum_learns_on_device(L, acts):
    """Learning-curve check: 1024 vectorised envs, full 200-step episodes, reference
    hyper-parameters except minibatch 4096: mean episode return must rise well above the random
    policy's (about -1200 .. -1500) within 40 iterations."""
    n_envs, T = 1024, 200
    cabi.srand(1)
    env = L.create_pendulum_env_cuda(n_envs, 1)
    ppo = make_ppo(L, [3, 64, 64, 1], acts, n_envs * T)
    L.ppo_b200_set_permutation_mode(ppo, 1, 99)
    L.ppo_b200_train_iterations(ppo, env, 1, 4096, 4, 10)
    r0 = L.ppo_b200_last_mean_return(ppo)
    best = r0
    for _ in range(8):
        L.ppo_b200_train_iterations(ppo, env, 5, 4096, 4, 10)
        best = max(best, L.ppo_b200_last_mean_return(ppo))
    print("pendulum mean return: first %.1f best %.1f" % (r0, best))
    assert -1800 < r0 < -900
    assert best > r0 + 300
    L.free_ppo(ppo)
    env.contents.free_env()


def test_checkpoint_roundtrip_and_reference_format(L, tmp_path):
    cabi.srand(9)
    sizes = [3, 16, 16, 1]
    ppo = make_ppo(L, sizes, RELU3, 640)
    T = oracle.Trainer(sizes, RELU3, batch_size=64, n_epochs_policy=1, n_epochs_value=1, init=False)
    b = synthetic_buffer(T, np.random.default_rng(0), sizes, RELU3, 640)
    fill_host_buffer(ppo, b)
    L.ppo_b200_update(ppo, 0.99, 64, 1, 1)
    path = str(tmp_path / "ppo_model.bin").encode()
    L.save_ppo(ppo, path)
    ppo2 = L.load_ppo(path, True)
    for get in (lambda p: p.contents.V, lambda p: p.contents.policy.contents.mu):
        assert np.array_equal(b200.nn_get_params(L, get(ppo)), b200.nn_get_params(L, get(ppo2)))
    assert ppo2.contents.adam_V.contents.time_step == ppo.contents.adam_V.contents.time_step == 10
    m1 = b200.d2h(L, ppo.contents.adam_V.contents.m, (ppo.contents.adam_V.contents.size,))
    m2 = b200.d2h(L, ppo2.contents.adam_V.contents.m, (ppo2.contents.adam_V.contents.size,))
    assert np.array_equal(m1, m2) and np.abs(m1).max() > 0
    # continuing from the checkpoint gives the same weights as continuing in place
    for p in (ppo, ppo2):
        fill_host_buffer(p, b)
        cabi.srand(77)
        L.ppo_b200_update(p, 0.99, 64, 1, 1)
    assert np.array_equal(b200.nn_get_params(L, ppo.contents.V), b200.nn_get_params(L, ppo2.contents.V))
    # byte format == the reference's (src/ppo.cu:585-607): let the unmodified reference read it back
    if os.path.exists(cabi.REF_SO):
        L.save_ppo(ppo, path)
        ref = cabi.load_ref()
        rp = ref.load_ppo(path, False)
        import refdrive
        assert np.array_equal(refdrive.Ref.nn_get_params(rp.contents.V), b200.nn_get_params(L, ppo.contents.V))
        assert np.array_equal(refdrive.Ref.nn_get_params(rp.contents.policy.contents.mu),
                              b200.nn_get_params(L, ppo.contents.policy.contents.mu))
        assert rp.contents.adam_policy.contents.time_step == ppo.contents.adam_policy.contents.time_step
        rm = np.ctypeslib.as_array(rp.contents.adam_V.contents.m, shape=(rp.contents.adam_V.contents.size,))
        assert np.array_equal(rm, b200.d2h(L, ppo.contents.adam_V.contents.m, rm.shape))
    L.free_ppo(ppo)
    L.free_ppo(ppo2)
