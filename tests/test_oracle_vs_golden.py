"""Pin the oracle (oracle/ppo_oracle.c) to the reference: every restated function must reproduce,
BIT FOR BIT, the golden vectors minted from the unmodified reference (tests/golden/make_golden.py).
CPU only."""
import numpy as np

import cabi
import oracle

RELU3 = ["relu", "relu", "none"]


def _acts(g, tag):
    n = len(g[f"mlp_{tag}_sizes"]) - 1
    return ["relu"] * (n - 1) + ["relu" if g[f"mlp_{tag}_relu_last"][0] else "none"]


def test_init_forward_backward_bitexact(golden):
    for tag in ["pend64", "cheetah32", "relu_out"]:
        sizes = [int(s) for s in golden[f"mlp_{tag}_sizes"]]
        cabi.srand(int(golden[f"mlp_{tag}_seed"][0]))
        p = oracle.init_params(sizes)
        assert np.array_equal(p, golden[f"mlp_{tag}_params"]), tag
        y, cache = oracle.mlp_forward(p, sizes, _acts(golden, tag), golden[f"mlp_{tag}_x"])
        assert np.array_equal(y, golden[f"mlp_{tag}_y"]), tag
        g = oracle.mlp_backward(p, sizes, _acts(golden, tag), cache, golden[f"mlp_{tag}_g"])
        assert np.array_equal(g, golden[f"mlp_{tag}_grads"]), tag


def test_gae_bitexact(golden):
    for tag in ["pend", "ragged", "long"]:
        a = {k: golden[f"gae_{tag}_{k}"] for k in ["r", "v", "vn", "term", "trunc"]}
        raw, tgt, norm, mean, std = oracle.gae(a["r"], a["v"], a["vn"], a["term"], a["trunc"], 0.99, 0.95)
        assert np.array_equal(norm, golden[f"gae_{tag}_adv_norm"]), tag
        assert np.array_equal(tgt, golden[f"gae_{tag}_adv_target"]), tag
        assert np.array_equal(tgt, a["v"] + raw)


def test_gae_float64_arbiter_close(golden):
    a = {k: golden[f"gae_long_{k}"] for k in ["r", "v", "vn", "term", "trunc"]}
    raw, tgt, norm, mean, std = oracle.gae(a["r"], a["v"], a["vn"], a["term"], a["trunc"], 0.99, 0.95)
    raw64, tgt64, norm64, mean64, std64 = oracle.gae_f64(a["r"], a["v"], a["vn"], a["term"], a["trunc"], 0.99, 0.95)
    assert np.max(np.abs(raw - raw64)) / np.max(np.abs(raw64)) < 2e-6
    assert np.max(np.abs(norm - norm64)) < 1e-4


def test_permutation_and_gather_bitexact(golden):
    cabi.srand(int(golden["perm_seed"][0]))
    n = golden["perm_state"].shape[0]
    mb = int(golden["perm_mb"][0])
    p0, p1 = oracle.shuffle(n), oracle.shuffle(n)
    assert np.array_equal(np.stack([p0, p1]), golden["perm_perms"])
    assert sorted(p0.tolist()) == list(range(n))
    args = [golden["perm_state"], golden["perm_action"], golden["perm_logprob"], golden["perm_adv"], golden["perm_advt"]]
    b0 = oracle.get_batch(p0, 0, mb, *args)
    assert np.array_equal(b0[0], golden["perm_b0_states"])
    assert np.array_equal(b0[1], golden["perm_b0_actions"])
    assert np.array_equal(b0[2], golden["perm_b0_logprob"])
    b7 = oracle.get_batch(p1, 7 - n // mb, mb, *args)  # 5 batches per shuffle -> batch 7 = 2nd shuffle, k=2
    assert np.array_equal(b7[0], golden["perm_b7_states"])
    assert np.array_equal(b7[4], golden["perm_b7_advt"])


def _policy_stage(g, pre, A_is_one):
    sizes = [int(s) for s in g[pre + "sizes"]]
    mu, cache = oracle.mlp_forward(g[pre + "params"], sizes, RELU3, g[pre + "state"])
    assert np.array_equal(mu, g[pre + "mu"])
    lp = oracle.log_prob(mu, g[pre + "log_std"], g[pre + "action"])
    assert np.array_equal(lp, g[pre + "logprob"])
    ent = oracle.entropy(g[pre + "log_std"])
    assert np.float32(ent) == g[pre + "entropy"]
    ec = float(g["pol_ent_coeff"][0]) if A_is_one else 0.0
    loss, gl, ge = oracle.policy_loss_and_grad(g[pre + "adv"], lp, g[pre + "lp_old"], ent, ec, 0.2)
    assert np.float32(loss) == g[pre + "loss"]
    assert np.array_equal(gl, g[pre + "grad_logprob"])
    assert np.float32(ge) == g[pre + "grad_entropy"]
    # both clip sides and the unclipped branch are exercised by the golden inputs
    assert (gl == 0).any() and (gl != 0).any()
    return sizes, mu, cache, gl


def test_policy_stage_bitexact_A1(golden):
    sizes, mu, cache, gl = _policy_stage(golden, "pol_", True)
    gmu, gls = oracle.log_prob_backwards(mu, golden["pol_log_std"], golden["pol_action"], gl, ref_index=True)
    assert np.array_equal(gmu, golden["pol_grad_mu"])
    assert np.array_equal(gls, golden["pol_grad_log_std"])
    # for A == 1 the corrected index is the same arithmetic
    gmu2, gls2 = oracle.log_prob_backwards(mu, golden["pol_log_std"], golden["pol_action"], gl, ref_index=False)
    assert np.array_equal(gmu, gmu2) and np.array_equal(gls, gls2)
    grads = oracle.mlp_backward(golden["pol_params"], sizes, RELU3, cache, gmu)
    assert np.array_equal(grads, golden["pol_grads"])


def test_policy_forward_side_bitexact_A6(golden):
    _policy_stage(golden, "pol6_", False)


def test_mse_bitexact(golden):
    assert np.float32(oracle.mse(golden["mse_y"], golden["mse_yt"])) == golden["mse_loss"][0]
    assert np.array_equal(oracle.mse_derivative(golden["mse_y"], golden["mse_yt"]), golden["mse_grad"])


def test_adam_bitexact(golden):
    w = golden["adam_w0"].copy()
    m, v, t = np.zeros_like(w), np.zeros_like(w), 0
    for g in golden["adam_grads"]:
        t = oracle.adam(w, g, m, v, 3e-4, t)
    assert t == int(golden["adam_t"][0])
    assert np.array_equal(w, golden["adam_w"]) and np.array_equal(m, golden["adam_m"]) and np.array_equal(v, golden["adam_v"])


def test_box_muller_bitexact(golden):
    cabi.srand(int(golden["noise_seed"][0]))
    z = np.array([oracle.gaussian_noise(1)[0] for _ in range(len(golden["noise_actions"]))], np.float32)
    assert np.array_equal(z, golden["noise_actions"])  # mu = 0, std = 1 -> action == noise
    lp = oracle.log_prob(np.zeros((len(z), 1), np.float32), np.zeros(1, np.float32), z[:, None])
    assert np.array_equal(lp, golden["noise_logprob"])


def test_whole_training_path_toy_env_bitexact(golden):
    seed, hidden, cap, steps, mb, n_pol, n_val = (int(x) for x in golden["toy_cfg"])
    cabi.srand(seed)
    T = oracle.Trainer([1, hidden, hidden, 1], RELU3, batch_size=mb, n_epochs_policy=n_pol, n_epochs_value=n_val)
    assert np.array_equal(T.mu, golden["toy_init_mu"]) and np.array_equal(T.v, golden["toy_init_v"])
    b = T.make_buffer(cap)
    for _ in range(steps // cap):
        assert T.collect(b, cap, 0) == 0
        T.update(b)
    for k in ["mu", "v", "log_std", "m_mu", "v_v"]:
        assert np.array_equal(getattr(T, k), golden["toy_" + k]), k
    for k in ["state", "action", "reward", "logprob", "advantage", "adv_target", "terminated", "truncated"]:
        assert np.array_equal(b[k].ravel(), golden["toy_" + k].ravel()), k
    assert cabi.rand() == int(golden["toy_rand_after"][0])  # same number of rand() draws consumed


def test_pendulum_known_answers():
    """[EXT] gymnasium Pendulum-v1 definition: hand-derived known answers."""
    th, thd, obs, r = oracle.pendulum_step(0.0, 0.0, 0.0)
    assert th == 0.0 and thd == 0.0 and r == 0.0 and obs.tolist() == [1.0, 0.0, 0.0]
    th, thd, obs, r = oracle.pendulum_step(np.pi, 0.0, 5.0)  # torque clipped to 2, cost = pi^2 + 0.004
    assert abs(r + (np.pi ** 2 + 0.004)) < 1e-6
    assert abs(thd - (15.0 * np.sin(np.pi) + 6.0) * 0.05) < 1e-12 and abs(th - (np.pi + thd * 0.05)) < 1e-12
    th, thd, obs, r = oracle.pendulum_step(1.0, 7.99, 2.0)  # speed clipped to 8
    assert thd == 8.0
    th, thd, obs, r = oracle.pendulum_step(-3.0 * np.pi / 2, 0.0, 0.0)  # angle_normalize(-3pi/2) = pi/2
    assert abs(r + (np.pi / 2) ** 2) < 1e-6


def test_welford_combine_matches_numpy():
    rng = np.random.default_rng(1)
    chunks = [rng.standard_normal(k).astype(np.float32) * 3 + 1 for k in (5, 1, 64, 17)]
    means = [c.mean(dtype=np.float64) for c in chunks]
    m2s = [((c - c.mean(dtype=np.float64)) ** 2).sum(dtype=np.float64) for c in chunks]
    mean, m2, n = oracle.welford_combine(means, m2s, [len(c) for c in chunks])
    allx = np.concatenate(chunks).astype(np.float64)
    assert n == allx.size and abs(mean - allx.mean()) < 1e-5 and abs(m2 / n - allx.var()) < 1e-4


def test_eval_ppo_matches_reference_output():
    """eval_ppo (src/ppo.cu:560-583): the oracle's J / R / Episodes, formatted like the reference's printf, must equal the
    line the unmodified reference printed (tests/golden/eval_golden.json, minted by tests/golden/make_eval_golden.py),
    and the rand() stream must sit at the same position afterwards."""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_golden.json")))
    assert len(cases) >= 4
    for c in cases:
        cabi.srand(c["seed"])
        sizes = [1, c["hidden"], c["hidden"], 1]
        T = oracle.Trainer(sizes, ["relu", "relu", "none"], batch_size=64, n_epochs_policy=1, n_epochs_value=2)
        if c["mu_bias"] is not None:
            T.mu[-1] = c["mu_bias"]
        b = T.make_buffer(c["capacity"])
        for _ in range(c["train_epochs"]):
            T.collect(b, c["capacity"], 0)
            T.update(b)
        J, R, n = T.eval(b, c["steps"], 0)
        assert "J: %f R: %f Episodes: %d" % (J, R, n) == c["line"], c
        assert cabi.rand() == c["rand_after"]
