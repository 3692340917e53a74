"""ctypes binding of oracle/libppo_oracle.so (the plain-C restatement; TEST INFRASTRUCTURE ONLY).

numpy-in / numpy-out wrappers around oracle/ppo_oracle.h.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libppo_oracle.so")

ACT = {"none": 0, "relu": 1, "tanh": 2}
f32, u8, i32 = np.float32, np.uint8, np.int32
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint8)


class OrcConfig(C.Structure):
    _fields_ = [("S", C.c_int), ("A", C.c_int), ("num_layers", C.c_int),
                ("sizes_mu", _ip), ("sizes_v", _ip), ("acts", _ip),
                ("lr_policy", C.c_float), ("lr_v", C.c_float), ("lambda_", C.c_float),
                ("epsilon", C.c_float), ("ent_coeff", C.c_float), ("gamma", C.c_float),
                ("batch_size", C.c_int), ("n_epochs_policy", C.c_int), ("n_epochs_value", C.c_int),
                ("ref_index", C.c_int)]


class OrcModel(C.Structure):
    _fields_ = [(n, _fp) for n in ["mu", "v", "log_std", "m_mu", "v_mu", "m_v", "v_v", "m_ls", "v_ls"]] + \
               [("t_mu", C.c_int), ("t_v", C.c_int), ("t_ls", C.c_int)]


class OrcBuffer(C.Structure):
    _fields_ = [("n", C.c_int)] + \
               [(n, _fp) for n in ["state", "next_state", "action", "reward", "logprob", "advantage", "adv_target"]] + \
               [("terminated", _up), ("truncated", _up)]


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


_lib = None


def use_fast_build():
    """bench.py's CPU arms only: switch to the timing-only build (`make -C oracle fast`: -O3 -march=native -ffast-math, the
    GEMM loops in vectorisable axpy order).  Rebuilt on the spot because -march=native must match the host that runs it.
    The parity tests never call this (they need the -O2 -ffp-contract=off build, bit-exact against the reference)."""
    global SO, _lib
    subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "fast"])
    SO = os.path.join(ROOT, "oracle", "libppo_oracle_fast.so")
    _lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "ppo_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
            _build()
        _lib = C.CDLL(SO, mode=C.RTLD_LOCAL)
        _lib.orc_entropy.restype = C.c_float
        _lib.orc_policy_loss_and_grad.restype = C.c_float
        _lib.orc_mse.restype = C.c_float
        _lib.orc_log_prob_one.restype = C.c_float
        _lib.orc_update.restype = C.c_float
        _lib.orc_mlp_output.restype = _fp
    return _lib


def fp(a):
    return a.ctypes.data_as(_fp)


def ip(a):
    return a.ctypes.data_as(_ip)


def up(a):
    return a.ctypes.data_as(_up)


def c32(a):
    return np.ascontiguousarray(a, dtype=f32)


def sizes_arr(sizes):
    return np.ascontiguousarray(sizes, dtype=i32)


def acts_arr(acts):
    return np.ascontiguousarray([ACT.get(a, 0) for a in acts], dtype=i32)


def param_count(sizes):
    return sum(sizes[i] * sizes[i + 1] + sizes[i + 1] for i in range(len(sizes) - 1))


def init_params(sizes):
    """neural_network.cu:40-51 with glibc rand() (call cabi.srand first)."""
    p = np.empty(param_count(sizes), f32)
    s = sizes_arr(sizes)
    lib().orc_init_params(fp(p), ip(s), len(sizes))
    return p


def mlp_forward(params, sizes, acts, x):
    m = x.shape[0]
    s, a = sizes_arr(sizes), acts_arr(acts)
    cache = np.empty(sum(sizes) * m, f32)
    x = c32(x)
    lib().orc_mlp_forward(fp(c32(params)), ip(s), ip(a), len(sizes), fp(x), m, fp(cache))
    out = cache[sum(sizes[:-1]) * m:].reshape(m, sizes[-1]).copy()
    return out, cache


def mlp_backward(params, sizes, acts, cache, grad_out, want_gx=False):
    m = grad_out.shape[0]
    s, a = sizes_arr(sizes), acts_arr(acts)
    grads = np.empty(param_count(sizes), f32)
    gx = np.empty((m, sizes[0]), f32) if want_gx else None
    lib().orc_mlp_backward(fp(c32(params)), ip(s), ip(a), len(sizes), fp(cache), fp(c32(grad_out)), m,
                           fp(grads), fp(gx) if want_gx else None)
    return (grads, gx) if want_gx else grads


def gae(reward, v, v_next, term, trunc, gamma, lam):
    n = reward.shape[0]
    raw, tgt, norm = np.empty(n, f32), np.empty(n, f32), np.empty(n, f32)
    mean, std = C.c_float(), C.c_float()
    lib().orc_gae(fp(c32(reward)), fp(c32(v)), fp(c32(v_next)), up(np.ascontiguousarray(term, u8)),
                  up(np.ascontiguousarray(trunc, u8)), n, C.c_float(gamma), C.c_float(lam),
                  fp(raw), fp(tgt), fp(norm), C.byref(mean), C.byref(std))
    return raw, tgt, norm, mean.value, std.value


def gae_f64(reward, v, v_next, term, trunc, gamma, lam):
    n = reward.shape[0]
    raw, tgt, norm = np.empty(n, np.float64), np.empty(n, np.float64), np.empty(n, np.float64)
    mean, std = C.c_double(), C.c_double()
    lib().orc_gae_f64(fp(c32(reward)), fp(c32(v)), fp(c32(v_next)), up(np.ascontiguousarray(term, u8)),
                      up(np.ascontiguousarray(trunc, u8)), n, C.c_float(gamma), C.c_float(lam),
                      raw.ctypes.data_as(_dp), tgt.ctypes.data_as(_dp), norm.ctypes.data_as(_dp),
                      C.byref(mean), C.byref(std))
    return raw, tgt, norm, mean.value, std.value


def welford_combine(means, m2s, ns):
    mean, m2, n = C.c_float(), C.c_float(), C.c_int()
    lib().orc_welford_combine(fp(c32(means)), fp(c32(m2s)), ip(np.ascontiguousarray(ns, i32)), len(ns),
                              C.byref(mean), C.byref(m2), C.byref(n))
    return mean.value, m2.value, n.value


def shuffle(limit):
    idx = np.empty(limit, i32)
    lib().orc_shuffle(ip(idx), limit)
    return idx


def get_batch(idx, batch_idx, mb, state, action, logprob, adv, advt):
    limit = idx.shape[0]
    S, A = state.shape[1], action.shape[1]
    o = [np.empty((mb, S), f32), np.empty((mb, A), f32), np.empty(mb, f32), np.empty(mb, f32), np.empty(mb, f32)]
    lib().orc_get_batch(ip(idx), limit, batch_idx, mb, S, A, fp(c32(state)), fp(c32(action)), fp(c32(logprob)),
                        fp(c32(adv)), fp(c32(advt)), *[fp(x) for x in o])
    return o


def gaussian_noise(n):
    out = np.zeros(max(n, 2) + 2, f32)
    lib().orc_gaussian_noise(fp(out), n)
    return out[:n]


def log_prob(mu, log_std, action):
    m, A = mu.shape
    out = np.empty(m, f32)
    lib().orc_log_prob(fp(c32(mu)), fp(c32(log_std)), fp(c32(action)), m, A, fp(out))
    return out


def log_prob_backwards(mu, log_std, action, grad_in, ref_index=False):
    m, A = mu.shape
    gmu, gls = np.empty((m, A), f32), np.empty(A, f32)
    lib().orc_log_prob_backwards(fp(c32(mu)), fp(c32(log_std)), fp(c32(action)), fp(c32(grad_in)), m, A,
                                 int(ref_index), fp(gmu), fp(gls))
    return gmu, gls


def entropy(log_std):
    return lib().orc_entropy(fp(c32(log_std)), len(log_std))


def policy_loss_and_grad(adv, lp, lp_old, entropy_, ent_coeff, eps):
    m = adv.shape[0]
    g = np.empty(m, f32)
    ge = C.c_float()
    loss = lib().orc_policy_loss_and_grad(fp(g), C.byref(ge), fp(c32(adv)), fp(c32(lp)), fp(c32(lp_old)),
                                          C.c_float(entropy_), C.c_float(ent_coeff), C.c_float(eps), m)
    return loss, g, ge.value


def mse(y, y_true):
    return lib().orc_mse(fp(c32(y)), fp(c32(y_true)), y.size, 1)


def mse_derivative(y, y_true):
    g = np.empty(y.size, f32)
    lib().orc_mse_derivative(fp(g), fp(c32(y)), fp(c32(y_true)), y.size, 1)
    return g


def adam(w, g, m, v, lr, t, beta1=0.9, beta2=0.999):
    """In-place on w, m, v (float32 contiguous); returns new time step."""
    tt = C.c_int(t)
    lib().orc_adam(fp(w), fp(c32(g)), fp(m), fp(v), w.size, C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
                   C.byref(tt))
    return tt.value


def pendulum_step(theta, theta_dot, action):
    th, thd = C.c_double(theta), C.c_double(theta_dot)
    obs = np.empty(3, f32)
    r = C.c_float()
    lib().orc_pendulum_step(C.byref(th), C.byref(thd), C.c_float(action), fp(obs), C.byref(r))
    return th.value, thd.value, obs, r.value


class Trainer:
    """Plain-array PPO state driven through orc_update / orc_collect (whole-path oracle)."""

    def __init__(self, sizes, acts, lr_policy=3e-4, lr_v=3e-4, lam=0.95, eps=0.2, ent_coeff=0.0,
                 init_std=1.0, gamma=0.99, batch_size=64, n_epochs_policy=4, n_epochs_value=10,
                 ref_index=None, init=True):
        self.sizes_mu = list(sizes)
        self.sizes_v = list(sizes[:-1]) + [1]
        self.acts = list(acts)
        self.S, self.A = sizes[0], sizes[-1]
        self._smu, self._sv, self._acts = sizes_arr(self.sizes_mu), sizes_arr(self.sizes_v), acts_arr(acts)
        if ref_index is None:
            ref_index = self.A == 1
        self.cfg = OrcConfig(self.S, self.A, len(sizes), ip(self._smu), ip(self._sv), ip(self._acts),
                             lr_policy, lr_v, lam, eps, ent_coeff, gamma, batch_size, n_epochs_policy,
                             n_epochs_value, int(ref_index))
        if init:  # create_ppo order: mu-net, then V-net (ppo.cu:10-16)
            self.mu = init_params(self.sizes_mu)
            self.v = init_params(self.sizes_v)
        else:
            self.mu = np.zeros(param_count(self.sizes_mu), f32)
            self.v = np.zeros(param_count(self.sizes_v), f32)
        self.log_std = np.full(self.A, np.log(f32(init_std)), f32)
        self.m_mu, self.v_mu = np.zeros_like(self.mu), np.zeros_like(self.mu)
        self.m_v, self.v_v = np.zeros_like(self.v), np.zeros_like(self.v)
        self.m_ls, self.v_ls = np.zeros_like(self.log_std), np.zeros_like(self.log_std)
        self.model = OrcModel(fp(self.mu), fp(self.v), fp(self.log_std), fp(self.m_mu), fp(self.v_mu),
                              fp(self.m_v), fp(self.v_v), fp(self.m_ls), fp(self.v_ls), 0, 0, 0)

    def make_buffer(self, n):
        b = dict(state=np.zeros((n, self.S), f32), next_state=np.zeros((n, self.S), f32),
                 action=np.zeros((n, self.A), f32), reward=np.zeros(n, f32), logprob=np.zeros(n, f32),
                 advantage=np.zeros(n, f32), adv_target=np.zeros(n, f32),
                 terminated=np.zeros(n, u8), truncated=np.zeros(n, u8))
        return b

    @staticmethod
    def _cbuf(b):
        n = b["reward"].shape[0]
        return OrcBuffer(n, fp(b["state"]), fp(b["next_state"]), fp(b["action"]), fp(b["reward"]),
                         fp(b["logprob"]), fp(b["advantage"]), fp(b["adv_target"]),
                         up(b["terminated"]), up(b["truncated"]))

    def update(self, b, log_perms=False, log_losses=False):
        n = b["reward"].shape[0]
        ne = self.cfg.n_epochs_policy + self.cfg.n_epochs_value
        perms = np.empty((ne, n), i32) if log_perms else None
        losses = np.empty(ne * (n // self.cfg.batch_size), f32) if log_losses else None
        cb = self._cbuf(b)
        lib().orc_update(C.byref(self.cfg), C.byref(self.model), C.byref(cb),
                         ip(perms) if log_perms else None, fp(losses) if log_losses else None)
        return perms, losses

    def collect(self, b, steps, env_id, start_idx=0):
        cb = self._cbuf(b)
        return lib().orc_collect(C.byref(self.cfg), C.byref(self.model), C.byref(cb), b["reward"].shape[0],
                                 start_idx, steps, env_id)

    def eval(self, b, steps, env_id, gamma=0.99):
        """eval_ppo (src/ppo.cu:560-583): returns (J, R, episodes) as float32 / int."""
        cb = self._cbuf(b)
        J, R, n = C.c_float(), C.c_float(), C.c_int()
        lib().orc_eval(C.byref(self.cfg), C.byref(self.model), C.byref(cb), b["reward"].shape[0], steps, env_id,
                       C.c_float(gamma), C.byref(J), C.byref(R), C.byref(n))
        return f32(J.value), f32(R.value), n.value
