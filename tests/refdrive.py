"""Drive the UNMODIFIED reference (oracle/_ref/libppo_ref.so) stage by stage through its own C ABI.

Used by tests/golden/make_golden.py (to mint the committed golden vectors) and by
tests/test_oracle_vs_ref.py (live re-check when the .so is present).  Only the plain-C twins are
called, so no GPU is needed.  Recipes follow SURVEY.md §8c "Verified recipe".
"""
import ctypes as C

import numpy as np

import cabi

f32, u8, i32 = np.float32, np.uint8, np.int32


class Ref:
    def __init__(self):
        cabi.unlimit_stack()
        self.lib = cabi.load_ref()

    @staticmethod
    def _release_buffer(buf):
        """Without a GPU the reference's cudaMalloc/cublasCreate calls fail silently and leave the
        d_* pointers and the cuBLAS handle uninitialised (trajectory_buffer.cu:59-67,
        neural_network.cu:29-32,68), so its free_* functions would hand garbage to cudaFree /
        cublasDestroy.  The driver therefore frees only the host arrays of buffers and leaks the
        (small) network / policy / PPO structs."""
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        b = buf.contents
        for name in ["h_state_p", "h_action_p", "h_next_state_p", "h_reward_p", "h_logprob_p",
                     "h_advantage_p", "h_adv_target_p", "h_terminated_p", "h_truncated_p"]:
            libc.free(C.cast(getattr(b, name), C.c_void_p))

    # ---- networks -------------------------------------------------------------------------
    def create_nn(self, sizes, acts):
        return self.lib.create_neural_network(cabi.int_array(sizes), cabi.cstr_array(acts), len(sizes))

    @staticmethod
    def nn_get_params(nn):
        out = []
        n = nn.contents
        for i in range(n.num_layers - 1):
            L = n.layers[i]
            out.append(np.ctypeslib.as_array(L.weights, shape=(L.input_size * L.output_size,)).copy())
            out.append(np.ctypeslib.as_array(L.biases, shape=(L.output_size,)).copy())
        return np.concatenate(out)

    @staticmethod
    def nn_get_grads(nn):
        out = []
        n = nn.contents
        for i in range(n.num_layers - 1):
            L = n.layers[i]
            out.append(np.ctypeslib.as_array(L.grad_weights, shape=(L.input_size * L.output_size,)).copy())
            out.append(np.ctypeslib.as_array(L.grad_biases, shape=(L.output_size,)).copy())
        return np.concatenate(out)

    @staticmethod
    def nn_set_params(nn, flat):
        n = nn.contents
        o = 0
        for i in range(n.num_layers - 1):
            L = n.layers[i]
            k = L.input_size * L.output_size
            np.ctypeslib.as_array(L.weights, shape=(k,))[:] = flat[o:o + k]
            o += k
            np.ctypeslib.as_array(L.biases, shape=(L.output_size,))[:] = flat[o:o + L.output_size]
            o += L.output_size

    def forward(self, nn, x):
        x = np.ascontiguousarray(x, f32)
        m = x.shape[0]
        self.lib.forward_propagation(nn, cabi.fptr(x), m)
        return np.ctypeslib.as_array(nn.contents.output, shape=(m, nn.contents.output_size)).copy()

    def backward(self, nn, grad_out):
        g = np.ascontiguousarray(grad_out, f32)
        self.lib.backward_propagation(nn, cabi.fptr(g), g.shape[0])
        return self.nn_get_grads(nn)

    # ---- GAE through an identity V-net --------------------------------------
    def gae(self, reward, v, v_next, term, trunc, gamma, lam):
        n = reward.shape[0]
        V = self.create_nn([1, 1], ["none"])
        self.nn_set_params(V, np.array([1.0, 0.0], f32))
        buf = self.lib.create_trajectory_buffer(n, 1, 1)
        b = buf.contents
        np.ctypeslib.as_array(b.state_p, shape=(n,))[:] = v
        np.ctypeslib.as_array(b.next_state_p, shape=(n,))[:] = v_next
        np.ctypeslib.as_array(b.reward_p, shape=(n,))[:] = reward
        np.ctypeslib.as_array(b.terminated_p, shape=(n,))[:] = term.astype(bool)
        np.ctypeslib.as_array(b.truncated_p, shape=(n,))[:] = trunc.astype(bool)
        np.ctypeslib.as_array(b.advantage_p, shape=(n,))[:] = 0
        b.idx = 0
        b.full = True
        self.lib.compute_gae(V, buf, C.c_float(gamma), C.c_float(lam))
        adv = np.ctypeslib.as_array(b.advantage_p, shape=(n,)).copy()
        tgt = np.ctypeslib.as_array(b.adv_target_p, shape=(n,)).copy()
        self._release_buffer(buf)
        return adv, tgt

    # ---- permutation / gather ---------------------------------------------------------------
    def shuffle_and_batches(self, seed, state, action, logprob, adv, advt, mb, n_shuffles=1):
        n, S = state.shape
        A = action.shape[1]
        buf = self.lib.create_trajectory_buffer(n, S, A)
        b = buf.contents
        np.ctypeslib.as_array(b.state_p, shape=(n, S))[:] = state
        np.ctypeslib.as_array(b.action_p, shape=(n, A))[:] = action
        np.ctypeslib.as_array(b.logprob_p, shape=(n,))[:] = logprob
        np.ctypeslib.as_array(b.advantage_p, shape=(n,))[:] = adv
        np.ctypeslib.as_array(b.adv_target_p, shape=(n,))[:] = advt
        b.idx = 0
        b.full = True
        cabi.srand(seed)
        perms, batches = [], []
        for _ in range(n_shuffles):
            self.lib.shuffle_buffer(buf)
            perms.append(np.ctypeslib.as_array(b.random_idx, shape=(n,)).copy())
            for k in range(n // mb):
                o = [np.empty((mb, S), f32), np.empty((mb, A), f32), np.empty(mb, f32), np.empty(mb, f32),
                     np.empty(mb, f32)]
                self.lib.get_batch(buf, k, mb, *[cabi.fptr(x) for x in o])
                batches.append(o)
        self._release_buffer(buf)
        return np.stack(perms), batches

    # ---- policy / losses --------------------------------------------------------------------
    def policy_stage(self, sizes, acts, params, log_std, state, action, adv, lp_old, ent_coeff, eps):
        """compute_log_prob -> entropy -> policy_loss_and_grad -> log_prob_backwards -> backward."""
        pol = self.lib.create_gaussian_policy(cabi.int_array(sizes), cabi.cstr_array(acts), len(sizes), C.c_float(1.0))
        p = pol.contents
        self.nn_set_params(p.mu, params)
        A = sizes[-1]
        np.ctypeslib.as_array(p.log_std, shape=(A,))[:] = log_std
        state = np.ascontiguousarray(state, f32)
        action = np.ascontiguousarray(action, f32)
        m = state.shape[0]
        lp = np.empty(m, f32)
        self.lib.compute_log_prob(pol, cabi.fptr(lp), cabi.fptr(state), cabi.fptr(action), m)
        mu = np.ctypeslib.as_array(p.mu.contents.output, shape=(m, A)).copy()
        ent = self.lib.compute_entropy(pol)
        g = np.empty(m, f32)
        ge = C.c_float()
        adv = np.ascontiguousarray(adv, f32)
        lp_old = np.ascontiguousarray(lp_old, f32)
        loss = self.lib.policy_loss_and_grad(cabi.fptr(g), C.byref(ge), cabi.fptr(adv), cabi.fptr(lp),
                                             cabi.fptr(lp_old), C.c_float(ent), C.c_float(ent_coeff),
                                             C.c_float(eps), m)
        out = dict(mu=mu, logprob=lp, entropy=f32(ent), loss=f32(loss), grad_logprob=g, grad_entropy=f32(ge.value))
        if A == 1:  # the reference's backward is only defined for A == 1
            gmu = np.empty((m, A), f32)
            gls = np.ctypeslib.as_array(p.log_std_grad, shape=(A,))
            self.lib.log_prob_backwards(pol, cabi.fptr(g), cabi.fptr(gmu), p.log_std_grad, m)
            out["grad_mu"] = gmu
            out["grad_log_std"] = gls.copy()
            self.lib.backward_propagation(p.mu, cabi.fptr(gmu), m)
            out["grads"] = self.nn_get_grads(p.mu)
        return out  # policy leaked on purpose, see _release_buffer

    def mse(self, y, y_true):
        y, y_true = np.ascontiguousarray(y, f32), np.ascontiguousarray(y_true, f32)
        loss = self.lib.mean_squared_error(cabi.fptr(y), cabi.fptr(y_true), y.size, 1)
        g = np.empty(y.size, f32)
        self.lib.mean_squared_error_derivative(cabi.fptr(g), cabi.fptr(y), cabi.fptr(y_true), y.size, 1)
        return f32(loss), g

    def adam_steps(self, w, grads_per_step, lr):
        w = np.ascontiguousarray(w, f32).copy()
        g = np.zeros_like(w)
        wp = (cabi.c_float_p * 1)(cabi.fptr(w))
        gp = (cabi.c_float_p * 1)(cabi.fptr(g))
        ln = cabi.int_array([w.size])
        ad = self.lib.create_adam(wp, gp, ln, 1, w.size, C.c_float(0.9), C.c_float(0.999))
        for gs in grads_per_step:
            g[:] = gs
            self.lib.adam_update(ad, C.c_float(lr))
        m = np.ctypeslib.as_array(ad.contents.m, shape=(w.size,)).copy()
        v = np.ctypeslib.as_array(ad.contents.v, shape=(w.size,)).copy()
        t = ad.contents.time_step
        self.lib.free_adam(ad)
        return w, m, v, t

    def gaussian_noise_via_sample(self, seed, n_draws):
        """sample_action with a zero mu-net and log_std=0 returns the raw Box-Muller noise (A=1)."""
        pol = self.lib.create_gaussian_policy(cabi.int_array([1, 1]), cabi.cstr_array(["none"]), 2, C.c_float(1.0))
        self.nn_set_params(pol.contents.mu, np.zeros(2, f32))
        cabi.srand(seed)
        s, a, lp = np.zeros(1, f32), np.zeros(1, f32), np.zeros(1, f32)
        out, lps = [], []
        for _ in range(n_draws):
            self.lib.sample_action(pol, cabi.fptr(s), cabi.fptr(a), cabi.fptr(lp), 1)
            out.append(a[0])
            lps.append(lp[0])
        return np.array(out, f32), np.array(lps, f32)

    # ---- whole training path on the toy env --------------------------------------------------
    def train_toy(self, seed, hidden, capacity, steps_per_epoch, mb, n_pol, n_val, n_calls=1):
        cabi.srand(seed)
        env = self.lib.create_simple_env(0, seed)
        sizes = [1, hidden, hidden, 1]
        acts = ["relu", "relu", "none"]
        ppo = self.lib.create_ppo(cabi.cstr_array(acts), cabi.int_array(sizes), 4, capacity, C.c_float(3e-4),
                                  C.c_float(3e-4), C.c_float(0.95), C.c_float(0.2), C.c_float(0.0),
                                  C.c_float(1.0), False)
        init_mu = self.nn_get_params(ppo.contents.policy.contents.mu)
        init_v = self.nn_get_params(ppo.contents.V)
        for _ in range(n_calls):
            self.lib.train_ppo_epoch(ppo, env, steps_per_epoch, mb, n_pol, n_val)
        p = ppo.contents
        b = p.buffer.contents
        n = capacity
        res = dict(
            init_mu=init_mu, init_v=init_v,
            mu=self.nn_get_params(p.policy.contents.mu), v=self.nn_get_params(p.V),
            log_std=np.ctypeslib.as_array(p.policy.contents.log_std, shape=(1,)).copy(),
            state=np.ctypeslib.as_array(b.state_p, shape=(n, 1)).copy(),
            action=np.ctypeslib.as_array(b.action_p, shape=(n, 1)).copy(),
            reward=np.ctypeslib.as_array(b.reward_p, shape=(n,)).copy(),
            logprob=np.ctypeslib.as_array(b.logprob_p, shape=(n,)).copy(),
            advantage=np.ctypeslib.as_array(b.advantage_p, shape=(n,)).copy(),
            adv_target=np.ctypeslib.as_array(b.adv_target_p, shape=(n,)).copy(),
            terminated=np.ctypeslib.as_array(b.terminated_p, shape=(n,)).astype(u8),
            truncated=np.ctypeslib.as_array(b.truncated_p, shape=(n,)).astype(u8),
            m_mu=np.ctypeslib.as_array(p.adam_policy.contents.m, shape=(p.adam_policy.contents.size,)).copy(),
            v_v=np.ctypeslib.as_array(p.adam_V.contents.v, shape=(p.adam_V.contents.size,)).copy(),
            rand_after=np.array([cabi.rand()], np.int64),
        )
        return res  # ppo leaked on purpose, see _release_buffer
