"""Mint tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libppo_ref.so).

Run in the build container (where /root/reference is mounted and `make -C oracle ref` works):

    python tests/golden/make_golden.py

Every array below is an INPUT we made up (seeded numpy / glibc rand) or an OUTPUT of a reference
plain-C function called through its own ABI (tests/refdrive.py).  The files are committed so the
GPU box, where /root/reference does not exist, can still check the oracle and the CUDA path
against the reference's own arithmetic.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cabi  # noqa: E402
import refdrive  # noqa: E402

f32, u8 = np.float32, np.uint8


def synth_gae(rng, T, N, p_term=1e-3, trunc_every=1000):
    """SURVEY.md §8d C5 generator, env-major flatten, forced done at each env's last step."""
    n = T * N
    r, v, vn = (rng.standard_normal(n).astype(f32) for _ in range(3))
    term = (rng.random(n) < p_term).astype(u8)
    t = np.tile(np.arange(T), N)
    trunc = (((t + 1) % trunc_every) == 0).astype(u8)
    trunc[t == T - 1] = 1
    return r, v, vn, term, trunc


def main():
    R = refdrive.Ref()
    out = {}
    rng = np.random.default_rng(20261018)

    # ---- init + forward/backward (a8-a10) --------------------------------------------------
    for tag, sizes, acts, m, seed in [("pend64", [3, 64, 64, 1], ["relu", "relu", "none"], 64, 3),
                                      ("cheetah32", [17, 32, 32, 6], ["relu", "relu", "none"], 48, 4),
                                      ("relu_out", [5, 8, 3], ["relu", "relu"], 7, 5)]:
        cabi.srand(seed)
        nn = R.create_nn(sizes, acts)
        x = rng.standard_normal((m, sizes[0])).astype(f32)
        g = rng.standard_normal((m, sizes[-1])).astype(f32)
        out[f"mlp_{tag}_seed"] = np.array([seed])
        out[f"mlp_{tag}_sizes"] = np.array(sizes)
        out[f"mlp_{tag}_relu_last"] = np.array([acts[-1] == "relu"])
        out[f"mlp_{tag}_params"] = R.nn_get_params(nn)
        out[f"mlp_{tag}_x"] = x
        out[f"mlp_{tag}_g"] = g
        out[f"mlp_{tag}_y"] = R.forward(nn, x)
        out[f"mlp_{tag}_grads"] = R.backward(nn, g)

    # ---- GAE (a5) ----------------------------------------------------------------------------
    for tag, T, N in [("pend", 200, 15), ("ragged", 37, 5), ("long", 2048, 8)]:
        r, v, vn, term, trunc = synth_gae(rng, T, N, p_term=5e-3 if tag != "long" else 1e-3)
        adv, tgt = R.gae(r, v, vn, term, trunc, 0.99, 0.95)
        assert np.isfinite(adv).all()
        for k, a in dict(r=r, v=v, vn=vn, term=term, trunc=trunc, adv_norm=adv, adv_target=tgt).items():
            out[f"gae_{tag}_{k}"] = a

    # ---- permutation + gather (a6, a7) ---------------------------------------------------------
    n, S, A, mb = 333, 3, 2, 64
    st, ac = rng.standard_normal((n, S)).astype(f32), rng.standard_normal((n, A)).astype(f32)
    lp, ad, at = (rng.standard_normal(n).astype(f32) for _ in range(3))
    perms, batches = R.shuffle_and_batches(99, st, ac, lp, ad, at, mb, n_shuffles=2)
    out.update(perm_seed=np.array([99]), perm_state=st, perm_action=ac, perm_logprob=lp, perm_adv=ad,
               perm_advt=at, perm_mb=np.array([mb]), perm_perms=perms,
               perm_b0_states=batches[0][0], perm_b0_actions=batches[0][1], perm_b0_logprob=batches[0][2],
               perm_b7_states=batches[7][0], perm_b7_advt=batches[7][4])

    # ---- policy stage (a11, a12) A == 1 (defined in the reference) ------------------------------
    sizes, acts, m = [3, 16, 16, 1], ["relu", "relu", "none"], 96
    cabi.srand(21)
    nn = R.create_nn(sizes, acts)
    params = R.nn_get_params(nn)
    state, action = rng.standard_normal((m, 3)).astype(f32), rng.standard_normal((m, 1)).astype(f32)
    adv = rng.standard_normal(m).astype(f32)
    log_std = np.array([-0.3], f32)
    probe = R.policy_stage(sizes, acts, params, log_std, state, action, adv, np.zeros(m, f32), 0.01, 0.2)
    lp_old = (probe["logprob"] + 0.3 * rng.standard_normal(m)).astype(f32)  # exercises both clip sides
    ps = R.policy_stage(sizes, acts, params, log_std, state, action, adv, lp_old, 0.01, 0.2)
    out.update(pol_sizes=np.array(sizes), pol_params=params, pol_log_std=log_std, pol_state=state,
               pol_action=action, pol_adv=adv, pol_lp_old=lp_old, pol_ent_coeff=np.array([0.01], f32),
               pol_eps=np.array([0.2], f32), **{"pol_" + k: np.asarray(v) for k, v in ps.items()})
    # A == 6: forward-side only (log-prob, entropy, loss, grad_logprob are valid for any A)
    sizes6 = [17, 16, 16, 6]
    cabi.srand(22)
    params6 = R.nn_get_params(R.create_nn(sizes6, acts))
    state6, action6 = rng.standard_normal((m, 17)).astype(f32), rng.standard_normal((m, 6)).astype(f32)
    ls6 = (0.2 * rng.standard_normal(6)).astype(f32)
    p6 = R.policy_stage(sizes6, acts, params6, ls6, state6, action6, adv, np.zeros(m, f32), 0.0, 0.2)
    lp_old6 = (p6["logprob"] + 0.3 * rng.standard_normal(m)).astype(f32)
    p6 = R.policy_stage(sizes6, acts, params6, ls6, state6, action6, adv, lp_old6, 0.0, 0.2)
    out.update(pol6_sizes=np.array(sizes6), pol6_params=params6, pol6_log_std=ls6, pol6_state=state6,
               pol6_action=action6, pol6_adv=adv, pol6_lp_old=lp_old6,
               **{"pol6_" + k: np.asarray(v) for k, v in p6.items()})

    # ---- MSE (a13) -------------------------------------------------------------------------------
    y, yt = rng.standard_normal(77).astype(f32), rng.standard_normal(77).astype(f32)
    loss, g = R.mse(y, yt)
    out.update(mse_y=y, mse_yt=yt, mse_loss=np.array([loss], f32), mse_grad=g)

    # ---- Adam (a14) ------------------------------------------------------------------------------
    w0 = rng.standard_normal(257).astype(f32)
    gs = (rng.standard_normal((5, 257)) * np.array([1, 1e-3, 10, 1e-6, 1])[:, None]).astype(f32)
    w, mm, vv, t = R.adam_steps(w0, gs, 3e-4)
    out.update(adam_w0=w0, adam_grads=gs, adam_w=w, adam_m=mm, adam_v=vv, adam_t=np.array([t]))

    # ---- Box-Muller via sample_action (a3) ---------------------------------------------------------
    noise, lps = R.gaussian_noise_via_sample(7, 16)
    out.update(noise_seed=np.array([7]), noise_actions=noise, noise_logprob=lps)

    # ---- whole path on the toy env (a2, a15): two iterations, capacity 4k+2 (see refdrive) --------
    res = R.train_toy(11, 8, 302, 604, 32, 2, 3)
    assert all(np.isfinite(v).all() for v in res.values())
    out.update({"toy_" + k: v for k, v in res.items()})
    out.update(toy_cfg=np.array([11, 8, 302, 604, 32, 2, 3]))

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
