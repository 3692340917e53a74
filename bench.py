#!/usr/bin/env python
"""bench.py — benchmarks of the B200-native PPO training path (contract: DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

Workloads = BASELINE.json configs:
  c1 (configs[0])  the reference's own default run: Pendulum-v1 behind the Env hooks (host env, one step at a time), 2x64 ReLU,
       capacity 3000, minibatch 64, 10 value + 4 policy epochs, 30 000 env-steps per train_ppo_epoch (src/main.c:33-43).
       GPU arm = the drop-in library behind the unmodified call; CPU arm = the UNMODIFIED reference (oracle/_ref) linked
       against the image's OpenBLAS (one thread, src/main.c:18) or, if that cannot load, its sequential-k cblas shim.
  c2 (default, configs[1])  Pendulum-v1 PPO, 4096 device envs/GPU x T=200, 2x64 tanh, fp32.  One step =
       one PPO iteration: fused device rollout (819 200 env-steps/GPU) + GAE + 10 value + 4 policy
       epochs of minibatch 18 944 = 148 SMs x 2 CTAs x 64 rows (the reference schedule, src/main.c:33-43).  Unit: env-steps/s.
  c3 (configs[2])  HalfCheetah-shaped synthetic buffer (S=17, A=6), T=2048 x N=512 per GPU, 2x256 ReLU,
       update only (GAE + 10 value + 4 policy epochs, minibatch 65536/GPU; --mb 4096 for the small-minibatch line).
       Unit: update samples/s.
  c4 (configs[3])  3x1024 actor-critic, minibatch 65 536/GPU, TF32 tcgen05 GEMMs, 262 144-sample
       synthetic buffer, update only.  Unit: update samples/s; roofline against tensor peak.
  c5 (configs[4])  GAE/returns sweep, T=2048 x N=65 536 per GPU, synthetic r/v/v'/flags.  Unit:
       env-steps/s (buffer elements/s); roofline against measured HBM bandwidth.

Per line:  value = whole-job throughput, inputs resident in HBM (CUDA events, max over ranks);
           e2e   = the same through the reference-facing C-ABI call with HOST buffers (pinned), every
                   host<->device copy inside the timed region;
           roofline / kernels = per-kernel CUDA-event timing of one extra profiled step
                   (ppo_b200_profile_*: an event pair around every launch, on the launching stream);
           cpu_baseline = plain-C restatement of the reference path (oracle/, "port") on one host core,
                   on a bounded sample of the same workload.
N > 1: torchrun, one process per GPU, weak scaling (per-GPU work fixed; the global minibatch is the union
of the rank-local ones), NCCL all-reduce of the flat gradient / all-gather of the Welford triple; c5
shards envs with no collective in the scan.  `--impl reference` times the CPU path only (rank 0), one
replica per host core.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

f32, u8 = np.float32, np.uint8
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 148 SMs x 128 FMA lanes x 2 x 1.965 GHz (nominal)
FP32_MEASURED = {}     # filled live by run_gpu_arm: ppo_b200_measure_fp32_peak (csrc/ubench.cu)


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": p["hbm_gbs"], "bf16": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "bf16_burst": p["bf16_tflops"], "src": "measured (MEASURED_PEAKS.json)"}
    except (OSError, KeyError, ValueError):
        return {"hbm": 6650.0, "bf16": 1400.0, "bf16_burst": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def mlp_weights(sizes):
    return sum(sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1))


def mlp_params(sizes):
    return mlp_weights(sizes) + sum(sizes[1:])


def train_flops(sizes, rows):
    """fwd 2*rows*W + dW 2*rows*W + dX 2*rows*(W - S*H0): layer-0 dX is never computed."""
    w = mlp_weights(sizes)
    return 2 * rows * w + 2 * rows * w + 2 * rows * (w - sizes[0] * sizes[1])


# ======================================================================================= workloads
class Workload:
    name = "?"
    metric = "env_steps_per_s"
    unit = "env-steps/s"
    dtype = "f32"

    def __init__(self, L, rank, world):
        self.L, self.rank, self.world = L, rank, world
        self.init_sizes()

    def init_sizes(self): ...

    # -- GPU arm
    def setup(self): ...
    def step_device(self, k): ...          # k steps, inputs resident in HBM
    def step_e2e(self, k): ...             # k steps through the host-buffer call
    def units_per_step(self): ...          # per GPU
    def e2e_bytes(self): return 0, 0
    def config(self): return {}
    def extra(self, ms): return {}
    def roofline_work(self, kernels): return {}
    def teardown(self): ...
    # -- CPU arm: seconds per sample step + units in that sample
    def cpu_sample(self, n_iters, seed=1): ...
    def cpu_sample_desc(self): return ""


def _fill_ppo_host_buffer(ppo, arrays):
    buf = ppo.contents.buffer.contents
    n = arrays["reward"].shape[0]
    S, A = arrays["state"].shape[1], arrays["action"].shape[1]
    np.ctypeslib.as_array(buf.h_state_p, shape=(n, S))[:] = arrays["state"]
    np.ctypeslib.as_array(buf.h_next_state_p, shape=(n, S))[:] = arrays["next_state"]
    np.ctypeslib.as_array(buf.h_action_p, shape=(n, A))[:] = arrays["action"]
    np.ctypeslib.as_array(buf.h_reward_p, shape=(n,))[:] = arrays["reward"]
    np.ctypeslib.as_array(buf.h_logprob_p, shape=(n,))[:] = arrays["logprob"]
    np.ctypeslib.as_array(buf.h_terminated_p, shape=(n,))[:] = arrays["terminated"].astype(bool)
    np.ctypeslib.as_array(buf.h_truncated_p, shape=(n,))[:] = arrays["truncated"].astype(bool)


def synthetic_rollout(rng, T, N, S, A):
    """SURVEY.md §8d C3 generator: N(0,1) states/actions/rewards, Bernoulli(1e-3) terminations,
    truncation every 1000 steps and forced on each env's last step, env-major flatten."""
    n = T * N
    t = np.tile(np.arange(T), N)
    trunc = (((t + 1) % 1000) == 0)
    trunc[t == T - 1] = True
    return dict(state=rng.standard_normal((n, S), dtype=f32), next_state=rng.standard_normal((n, S), dtype=f32),
                action=rng.standard_normal((n, A), dtype=f32), reward=rng.standard_normal(n, dtype=f32),
                logprob=np.zeros(n, f32), terminated=(rng.random(n) < 1e-3).astype(u8), truncated=trunc.astype(u8))


class C1(Workload):
    """BASELINE.json configs[0]: the reference's default run (src/main.c:20-43) — one Pendulum env behind the Env hooks."""
    name = "c1"
    SIZES, ACTS = [3, 64, 64, 1], ["relu", "relu", "none"]
    CAP, STEPS, MB, N_POL, N_VAL = 3000, 30000, 64, 4, 10

    def setup(self):
        import cabi
        L = self.L
        cabi.srand(1234)
        self.env = L.create_pendulum_env(0, 100 + self.rank)          # host env: same hooks as the reference's gym bridge
        self.ppo = L.create_ppo(cabi.cstr_array(self.ACTS), cabi.int_array(self.SIZES), len(self.SIZES), self.CAP,
                                3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)

    def step_device(self, k):
        for _ in range(k):
            self.L.train_ppo_epoch(self.ppo, self.env, self.STEPS, self.MB, self.N_POL, self.N_VAL)

    step_e2e = step_device      # the env lives on the host: every transition crosses the boundary in both legs

    def units_per_step(self):
        return self.STEPS

    def e2e_bytes(self):
        S, A = self.SIZES[0], self.SIZES[-1]
        iters = self.STEPS // self.CAP
        buf = self.CAP * (4 * (2 * S + A + 4) + 2)
        p = mlp_params(self.SIZES) + mlp_params(self.SIZES[:-1] + [1]) + A
        # per env step: the state goes up, action + log-prob come down; per iteration: buffer up, buffer + weights down
        return self.STEPS * 4 * S + iters * buf, self.STEPS * 4 * (A + 1) + iters * (buf + 4 * p)

    def config(self):
        return {"workload": "c1: the reference's default run (src/main.c:20-43): Pendulum-v1 behind the Env hooks (one host env), "
                            "2x64 ReLU, capacity %d, minibatch %d, %d value + %d policy epochs, %d env-steps per train_ppo_epoch "
                            "(BASELINE.json configs[0])" % (self.CAP, self.MB, self.N_VAL, self.N_POL, self.STEPS),
                "parallelism": "replicas x%d" % self.world,
                "l2": "latency-bound config (18 KB nets, 3000-row buffer): everything is cache resident by construction, as in the reference",
                "permutation": "reference rand() swap chain (bit-exact mode)",
                "e2e_call": "train_ppo_epoch(ppo, env, 30000, 64, 4, 10) with a HOST env: per step the state crosses to the device and "
                            "the sampled action / log-prob come back; per iteration buffer_to_device + update + host mirrors"}

    def extra(self, ms):
        nb = self.CAP // self.MB
        return {"update_samples_per_s": self.world * (self.STEPS // self.CAP) * (self.N_POL + self.N_VAL) * nb * self.MB / (ms * 1e-3)}

    def roofline_work(self, kernels):
        it, nb = self.STEPS // self.CAP, self.CAP // self.MB
        sv = self.SIZES[:-1] + [1]
        flops = it * nb * (self.N_VAL * train_flops(sv, self.MB) + self.N_POL * train_flops(self.SIZES, self.MB))
        return {"fused_phase": ("fp32", flops), "sample_action": ("fp32", 2.0 * self.STEPS * mlp_weights(self.SIZES))}

    def teardown(self):
        self.L.free_ppo(self.ppo)
        self.env.contents.free_env()

    # CPU arm: the UNMODIFIED reference on the oracle's C Pendulum behind its own Env hooks
    ref_kind = "reference"

    def cpu_sample(self, n_iters, seed=1, blas=True):
        import cabi
        import oracle
        cabi.unlimit_stack()
        lib = cabi.load_ref_blas() if blas else None
        self.cpu_blas = "OpenBLAS 0.3.15 (bundled with the image), 1 thread" if lib is not None else "sequential-k cblas shim (oracle/shim)"
        if lib is None:
            lib = cabi.load_ref()
        env = cabi.oracle_pendulum_env(oracle.lib())
        cabi.srand(seed)
        ppo = lib.create_ppo(cabi.cstr_array(self.ACTS), cabi.int_array(self.SIZES), len(self.SIZES), self.CAP, C.c_float(3e-4),
                             C.c_float(3e-4), C.c_float(0.95), C.c_float(0.2), C.c_float(0.0), C.c_float(1.0), False)
        times = []
        for _ in range(n_iters):
            t0 = time.perf_counter()
            lib.train_ppo_epoch(ppo, C.byref(env), self.STEPS, self.MB, self.N_POL, self.N_VAL)
            times.append(time.perf_counter() - t0)
        return times, self.STEPS

    def cpu_sample_desc(self):
        return ("the UNMODIFIED reference (oracle/_ref, use_cuda=false) through its own train_ppo_epoch on the C Pendulum behind "
                "its Env hooks; BLAS = %s; %d env-steps per step" % (getattr(self, "cpu_blas", "OpenBLAS if loadable"), self.STEPS))


class C2(Workload):
    """Pendulum-v1 PPO, vectorised device envs (BASELINE.json configs[1])."""
    name = "c2"
    SIZES, ACTS = [3, 64, 64, 1], ["tanh", "tanh", "none"]
    # minibatch = 148 SMs x 2 resident CTAs x 64 rows: every SM holds exactly two tiles of each minibatch (at 16384 rows
    # 40 of the 148 SMs hold one).  43 minibatches per epoch; like the reference (src/ppo.cu:387: limit / batch_size) the
    # remainder of the shuffled buffer (0.56 %) is not visited in that epoch.
    N_ENVS, T, MB, N_POL, N_VAL = 4096, 200, 18944, 4, 10
    CPU_STEPS, CPU_MB = 8200, 2050           # 41 episodes of 200 steps, 4 minibatches per epoch

    def init_sizes(self):
        self.cap = self.N_ENVS * self.T

    def setup(self):
        import cabi
        L = self.L
        cabi.srand(1234)                      # same initial weights on every rank
        self.env = L.create_pendulum_env_cuda(self.N_ENVS, 100 + self.rank)
        self.ppo = L.create_ppo(cabi.cstr_array(self.ACTS), cabi.int_array(self.SIZES), len(self.SIZES), self.cap,
                                3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
        L.ppo_b200_set_permutation_mode(self.ppo, 1, 7 + self.rank)

    def step_device(self, k):
        self.L.ppo_b200_train_iterations(self.ppo, self.env, k, self.MB, self.N_POL, self.N_VAL)

    def step_e2e(self, k):
        """The reference's data flow with HOST buffers (src/ppo.cu:482-538): the rollout lands in the host buffer, every
        iteration uploads it (buffer_to_device), updates, and mirrors buffer + weights back to the host."""
        L, ppo = self.L, self.ppo
        for _ in range(k):
            L.collect_trajectories(ppo.contents.buffer, self.env, ppo.contents.policy, self.cap)   # fused device rollout
            L.buffer_to_host(ppo.contents.buffer)                                                   # D2H: rollout -> pinned host arrays
            L.ppo_b200_update(ppo, 0.99, self.MB, self.N_POL, self.N_VAL)                           # H2D + GAE + epochs + D2H mirrors

    def step_api(self, k):
        """train_ppo_epoch, the one call a ppo.c user makes: device rollout + update + host mirrors (no host input exists)."""
        self.L.train_ppo_epoch(self.ppo, self.env, self.cap * k, self.MB, self.N_POL, self.N_VAL)

    def units_per_step(self):
        return self.cap

    def e2e_bytes(self):
        S, A = self.SIZES[0], self.SIZES[-1]
        p_mu = mlp_params(self.SIZES)
        p_v = mlp_params(self.SIZES[:-1] + [1])
        buf = self.cap * (4 * (2 * S + A + 4) + 2)            # all nine arrays (buffer_to_host after the rollout)
        # ppo_b200_update moves the seven input arrays up and only advantage / adv_target (+ weights) down
        return buf - 8 * self.cap, buf + 8 * self.cap + 4 * (p_mu + p_v + A)

    def config(self):
        return {"workload": "c2: Pendulum-v1 PPO, %d device envs/GPU x T=%d, 2x64 tanh MLP, fp32, minibatch %d/GPU, "
                            "%d value + %d policy epochs (BASELINE.json configs[1])" % (self.N_ENVS, self.T, self.MB, self.N_VAL, self.N_POL),
                "env_steps_per_step_per_gpu": self.cap, "parallelism": "dp%d" % self.world,
                "minibatch": "%d = 148 SMs x 2 resident CTAs x 64 rows; %d minibatches per epoch, the last %d shuffled rows of an epoch are "
                             "not visited (reference semantics, src/ppo.cu:387)" % (self.MB, self.cap // self.MB, self.cap % self.MB),
                "l2": "inputs larger than L2: every step streams the 819200-row buffer (38 MB of rows + 14 permutations) "
                      "through 602 minibatch launches and rewrites it in the rollout; no explicit flush",
                "permutation": "device generator (mode 1); the reference's host rand() chain is the bit-exact mode of the parity tests",
                "e2e_call": "reference data flow with HOST buffers (src/ppo.cu:482-538): collect_trajectories (device rollout) -> "
                            "buffer_to_host (D2H of all nine arrays, pinned) -> ppo_b200_update = H2D of the seven input arrays + GAE + epochs + "
                            "D2H of advantage / adv_target + policy_to_host / nn_write_weights_to_host; `api` = train_ppo_epoch, the single call a "
                            "user makes (device rollout, no host input, host mirrors refreshed every iteration)"}

    def extra(self, ms):
        nb = self.cap // self.MB
        return {"update_samples_per_s": self.world * (self.N_POL + self.N_VAL) * nb * self.MB / (ms * 1e-3),
                "mean_episode_return": self.L.ppo_b200_last_mean_return(self.ppo)}

    def roofline_work(self, kernels):
        B, nb = self.cap, self.cap // self.MB
        S, A = self.SIZES[0], self.SIZES[-1]
        sv = self.SIZES[:-1] + [1]
        upd_flops = self.N_VAL * nb * train_flops(sv, self.MB) + self.N_POL * nb * train_flops(self.SIZES, self.MB) \
            + 2 * 2 * B * mlp_weights(sv)                                  # + the two V forwards of the GAE
        slab_v, slab_p = mlp_params(sv) + 2, mlp_params(self.SIZES) + A + 1
        ctas = -(-self.MB // 64)
        red_bytes = (self.N_VAL * nb * (ctas * slab_v * 4 + 32 * mlp_params(sv))
                     + self.N_POL * nb * (ctas * slab_p * 4 + 32 * mlp_params(self.SIZES)))
        roll_flops = 2 * B * mlp_weights(self.SIZES)
        if any("fused_phase" in k for k in kernels):          # persistent path: the whole update is one kernel per phase
            gae_fwd = 2 * 2 * B * mlp_weights(sv)
            return {"fused_phase": ("fp32", upd_flops - gae_fwd), "fused_tile64_kernel": ("fp32", gae_fwd),
                    "rollout64_kernel": ("fp32", roll_flops), "rollout_kernel": ("fp32", roll_flops),
                    "gae_scan": ("hbm", 22.0 * B), "gae_normalize_kernel": ("hbm", 8.0 * B)}
        return {"fused_tile64_kernel": ("fp32", upd_flops),
                "fused_reduce_adam_kernel": ("hbm", red_bytes),      # slab reads + 28 B/param Adam + 4 B/param image
                "rollout64_kernel": ("fp32", roll_flops), "rollout_kernel": ("fp32", roll_flops),
                "gae_scan": ("hbm", 22.0 * B), "gae_normalize_kernel": ("hbm", 8.0 * B)}

    def teardown(self):
        self.L.free_ppo(self.ppo)
        self.env.contents.free_env()

    def cpu_sample(self, n_iters, seed=1):
        import cabi
        import oracle
        cabi.srand(seed)
        tr = oracle.Trainer(self.SIZES, self.ACTS, batch_size=self.CPU_MB, n_epochs_policy=self.N_POL, n_epochs_value=self.N_VAL)
        buf = tr.make_buffer(self.CPU_STEPS)
        times = []
        for _ in range(n_iters):
            t0 = time.perf_counter()
            tr.collect(buf, self.CPU_STEPS, 1)
            tr.update(buf)
            times.append(time.perf_counter() - t0)
        return times, self.CPU_STEPS

    def cpu_sample_desc(self):
        return ("oracle port of the reference plain-C path (src/ppo.cu:373-448 + C Pendulum): 1 env, %d env-steps per step "
                "(rollout + GAE + %d value / %d policy epochs, minibatch %d), 2x64 tanh"
                % (self.CPU_STEPS, self.N_VAL, self.N_POL, self.CPU_MB))


class UpdateOnly(Workload):
    """Synthetic rollout buffer, PPO update only (GAE + value epochs + policy epochs)."""
    metric, unit = "update_samples_per_s", "samples/s"
    SIZES, ACTS = None, None
    T, N, MB, N_POL, N_VAL = 0, 0, 0, 4, 10
    TF32 = 0
    CPU_B, CPU_MB, CPU_POL, CPU_VAL = 4096, 1024, 1, 1

    def init_sizes(self):
        self.cap = self.T * self.N

    def setup(self):
        import cabi
        L = self.L
        S, A = self.SIZES[0], self.SIZES[-1]
        L.ppo_b200_set_matmul_precision(self.TF32)
        cabi.srand(1234)
        self.ppo = L.create_ppo(cabi.cstr_array(self.ACTS), cabi.int_array(self.SIZES), len(self.SIZES), self.cap,
                                3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
        rng = np.random.default_rng(1000 + self.rank)
        arrays = synthetic_rollout(rng, self.T, self.N, S, A)
        _fill_ppo_host_buffer(self.ppo, arrays)
        L.ppo_b200_buffer_upload(self.ppo)
        # logprob_old = log-prob under the initial policy + N(0, 0.1): ratios near 1, both clip branches live
        buf = self.ppo.contents.buffer.contents
        L.compute_log_prob_cuda(self.ppo.contents.policy, buf.d_logprob_p, buf.d_state_p, buf.d_action_p, self.cap)
        lp = np.empty(self.cap, f32)
        L.ppo_b200_d2h(lp.ctypes.data, C.cast(buf.d_logprob_p, C.c_void_p), lp.nbytes)
        lp += 0.1 * rng.standard_normal(self.cap, dtype=f32)
        np.ctypeslib.as_array(buf.h_logprob_p, shape=(self.cap,))[:] = lp
        L.ppo_b200_buffer_upload(self.ppo)
        L.ppo_b200_set_permutation_mode(self.ppo, 1, 7 + self.rank)

    def step_device(self, k):
        for _ in range(k):
            self.L.ppo_b200_update_device(self.ppo, 0.99, self.MB, self.N_POL, self.N_VAL)

    def step_e2e(self, k):
        for _ in range(k):
            self.L.ppo_b200_update(self.ppo, 0.99, self.MB, self.N_POL, self.N_VAL)
        self.L.ppo_b200_set_permutation_mode(self.ppo, 1, 7 + self.rank)

    def units_per_step(self):
        return (self.N_POL + self.N_VAL) * (self.cap // self.MB) * self.MB

    def e2e_bytes(self):
        S, A = self.SIZES[0], self.SIZES[-1]
        per_row = 4 * (2 * S + A + 4) + 2
        p = mlp_params(self.SIZES) + mlp_params(self.SIZES[:-1] + [1]) + A
        # ppo_b200_update: seven input arrays up, advantage / adv_target + weights down
        return self.cap * (per_row - 8), self.cap * 8 + 4 * p

    def base_config(self, label):
        return {"workload": label, "buffer_rows_per_gpu": self.cap, "minibatch_per_gpu": self.MB,
                "epochs": "%d value + %d policy" % (self.N_VAL, self.N_POL), "parallelism": "dp%d" % self.world,
                "l2": "inputs larger than L2: the %d-row buffer (%.0f MB of gathered fields) is re-streamed by every epoch"
                      % (self.cap, self.cap * 4 * (self.SIZES[0] + self.SIZES[-1] + 3) / 1e6),
                "permutation": "device generator (mode 1); the reference's host rand() chain is the bit-exact mode "
                               "used by the parity tests",
                "e2e_call": "ppo_b200_update on a HOST-filled buffer: upload of the seven input arrays (pinned host) + GAE + epochs + "
                            "download of advantage / adv_target + policy_to_host / nn_write_weights_to_host"}

    def teardown(self):
        self.L.free_ppo(self.ppo)
        self.L.ppo_b200_set_matmul_precision(0)

    def cpu_sample(self, n_iters, seed=1):
        import cabi
        import oracle
        cabi.srand(seed)
        tr = oracle.Trainer(self.SIZES, self.ACTS, batch_size=self.CPU_MB, n_epochs_policy=self.CPU_POL,
                            n_epochs_value=self.CPU_VAL, ref_index=False)
        rng = np.random.default_rng(seed)
        arrays = synthetic_rollout(rng, self.CPU_B, 1, self.SIZES[0], self.SIZES[-1])
        b = tr.make_buffer(self.CPU_B)
        for k, v in arrays.items():
            b[k][:] = v
        times = []
        for _ in range(n_iters):
            t0 = time.perf_counter()
            tr.update(b)
            times.append(time.perf_counter() - t0)
        return times, (self.CPU_POL + self.CPU_VAL) * (self.CPU_B // self.CPU_MB) * self.CPU_MB

    def cpu_sample_desc(self):
        return ("oracle port of the reference plain-C update (src/ppo.cu:373-448): %d-row synthetic buffer, GAE + %d value + %d "
                "policy epochs, minibatch %d, nets %s" % (self.CPU_B, self.CPU_VAL, self.CPU_POL, self.CPU_MB, self.SIZES))


class C3(UpdateOnly):
    name = "c3"
    SIZES, ACTS = [17, 256, 256, 6], ["relu", "relu", "none"]
    T, N, MB = 2048, 512, 65536          # --mb 4096 gives the small-minibatch (launch-bound) line, profiles/r01_bench_c3_mb4096.json

    def config(self):
        return self.base_config("c3: HalfCheetah-shaped synthetic rollout buffer (S=17, A=6), T=2048 x N=512 per GPU, 2x256 ReLU "
                                "MLP, fp32 FFMA layer kernels, PPO update only (BASELINE.json configs[2])")

    def roofline_work(self, kernels):
        nb = self.cap // self.MB
        sv = self.SIZES[:-1] + [1]
        w_mu, w_v = mlp_weights(self.SIZES), mlp_weights(sv)
        steps_v, steps_p = self.N_VAL * nb, self.N_POL * nb
        S, H = self.SIZES[0], self.SIZES[1]
        A = self.SIZES[-1]
        # keys are substrings of kernel names: "kernel<kFwd>" covers sgemm_kernel<kFwd> and sgemm128_kernel<kFwd>; the
        # <= 8-wide heads run in linear_forward_skinny_kernel / skinny_dw_kernel and are left out of these FLOP counts
        return {"kernel<kFwd>": ("fp32", 2 * self.MB * (steps_v * (w_v - H) + steps_p * (w_mu - H * A)) + 2 * 2 * self.cap * (w_v - H)),
                "kernel<kBwdInput>": ("fp32", 2 * self.MB * (steps_v * (w_v - S * H) + steps_p * (w_mu - S * H))),
                "kernel<kBwdParam>": ("fp32", 2 * self.MB * (steps_v * (w_v - H - S * H) + steps_p * (w_mu - H * A - S * H))),
                "skinny_dw_kernel": ("fp32", 2 * self.MB * (steps_v * (H + S * H) + steps_p * (H * A + S * H))),
                "gather_kernel": ("hbm", (steps_v + steps_p) * self.MB * (4 + 2 * 4 * (self.SIZES[0] + self.SIZES[-1] + 3))),
                "adam_flat_kernel": ("hbm", 28.0 * (steps_v * mlp_params(sv) + steps_p * (mlp_params(self.SIZES) + self.SIZES[-1]))),
                "gae_scan": ("hbm", 22.0 * self.cap), "gae_normalize_kernel": ("hbm", 8.0 * self.cap)}


class C3X3(C3):
    """c3 with the 256x256 layers as 3xTF32 split contractions on the tensor cores (precision 3): fp32-accurate, same parity tests."""
    name = "c3x3"
    TF32 = 3

    def config(self):
        c = self.base_config("c3x3: HalfCheetah-shaped synthetic rollout buffer (S=17, A=6), T=2048 x N=512 per GPU, 2x256 ReLU MLP, the "
                             "256x256 layers as 3xTF32 split tcgen05 contractions (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM), "
                             "PPO update only (BASELINE.json configs[2])")
        c["tolerance"] = "same as the fp32 FFMA path: 1e-5 (tests/test_gpu_tc.py::test_tc_x3_*, test_update_c3_minibatch_65536_matches_oracle[3xtf32])"
        return c

    def roofline_work(self, kernels):
        w = dict(C3.roofline_work(self, kernels))
        nb = self.cap // self.MB
        steps = (self.N_VAL + self.N_POL) * nb
        H = self.SIZES[1]
        w["tc_gemm_kernel"] = ("tensor_3xtf32", 6 * self.MB * H * H * steps + 2 * 2 * self.cap * H * H)
        return w


class C4(UpdateOnly):
    name = "c4"
    dtype = "tf32"
    SIZES, ACTS = [17, 1024, 1024, 1024, 6], ["relu", "relu", "relu", "none"]
    T, N, MB = 2048, 128, 65536
    TF32 = 1
    CPU_B, CPU_MB = 512, 256

    def config(self):
        c = self.base_config("c4: wide actor-critic 3x1024 ReLU MLP (S=17, A=6), minibatch 65536/GPU, TF32 tcgen05 GEMMs with fp32 "
                             "accumulation in TMEM, 262144-row synthetic buffer per GPU, PPO update only (BASELINE.json configs[3])")
        c["tolerance"] = "TF32 operands (RNA-rounded), stated separately from fp32: ~1e-3 norm-wise (tests/test_gpu_tc.py)"
        return c

    def roofline_work(self, kernels):
        nb = self.cap // self.MB
        sv = self.SIZES[:-1] + [1]
        steps_v, steps_p = self.N_VAL * nb, self.N_POL * nb
        # tensor-core layers = those with in/out >= 64: the three... two 1024x1024 layers fwd/dX/dW
        wide = sum(self.SIZES[i] * self.SIZES[i + 1] for i in range(len(self.SIZES) - 1)
                   if self.SIZES[i] >= 64 and self.SIZES[i + 1] >= 64)
        tc = 6 * self.MB * wide * (steps_v + steps_p) + 2 * 2 * self.cap * wide
        return {"tc_gemm_kernel": ("tensor_tf32", tc),
                "adam_flat_kernel": ("hbm", 28.0 * (steps_v * mlp_params(sv) + steps_p * (mlp_params(self.SIZES) + self.SIZES[-1]))),
                "gather_kernel": ("hbm", (steps_v + steps_p) * self.MB * (4 + 2 * 4 * (self.SIZES[0] + self.SIZES[-1] + 3))),
                "gae_scan": ("hbm", 22.0 * self.cap), "gae_normalize_kernel": ("hbm", 8.0 * self.cap)}


class C4BF16(C4):
    """c4 with BF16 operands (tcgen05 kind::f16, fp32 accumulation): the same shapes and data flow, peak = sustained cuBLAS bf16."""
    name = "c4bf16"
    dtype = "bf16"
    TF32 = 2

    def config(self):
        c = self.base_config("c4bf16: wide actor-critic 3x1024 ReLU MLP (S=17, A=6), minibatch 65536/GPU, BF16-operand tcgen05 GEMMs with "
                             "fp32 accumulation in TMEM (fp32 parameters, gradients and Adam), 262144-row synthetic buffer per GPU, PPO "
                             "update only (BASELINE.json configs[3], bf16 variant)")
        c["tolerance"] = "BF16 operands (round-to-nearest-even), stated separately from fp32: ~3e-3 norm-wise per GEMM (tests/test_gpu_tc.py)"
        return c

    def roofline_work(self, kernels):
        w = dict(C4.roofline_work(self, kernels))
        w["tc_gemm_kernel"] = ("tensor_bf16", w["tc_gemm_kernel"][1])
        return w


class C5(Workload):
    """GAE/returns-only sweep on synthetic rewards/values/dones (BASELINE.json configs[4])."""
    name = "c5"
    T, N, BLOCK = 2048, 65536, 4096
    CPU_N = 4096

    def _block(self, rng, nb):
        n = self.T * nb
        t = np.tile(np.arange(self.T), nb)
        trunc = (((t + 1) % 1000) == 0)
        trunc[t == self.T - 1] = True
        return (rng.standard_normal(n, dtype=f32), rng.standard_normal(n, dtype=f32), rng.standard_normal(n, dtype=f32),
                (rng.random(n) < 1e-3).astype(u8), trunc.astype(u8))

    def init_sizes(self):
        self.n = self.T * self.N

    def setup(self):
        import b200
        L = self.L
        rng = np.random.default_rng(500 + self.rank)
        blk = self._block(rng, self.BLOCK)
        reps = self.N // self.BLOCK
        self.dev = []
        for a in blk:                       # the 4096-env block replicated 16x on the device (3.9 GB of inputs)
            d = b200.dev_empty(self.n, a.dtype)
            for i in range(reps):
                L.ppo_b200_h2d(d.ptr + i * a.nbytes, a.ctypes.data, a.nbytes)
            self.dev.append(d)
        self.adv, self.tgt, self.st = b200.dev_empty(self.n), b200.dev_empty(self.n), b200.dev_empty(2)
        self.blk, self.host = blk, None

    def _gae(self):
        self.L.ppo_b200_gae(*[d.ptr for d in self.dev], self.n, 0.99, 0.95, self.adv.ptr, self.tgt.ptr, 1, self.st.ptr)

    def step_device(self, k):
        for _ in range(k):
            self._gae()

    def _host_alloc(self):
        L = self.L
        sizes = [4 * self.n] * 3 + [self.n] * 2 + [4 * self.n] * 2
        self.host = [L.ppo_b200_malloc_host(s) for s in sizes]
        reps = self.N // self.BLOCK
        for h, a in zip(self.host[:5], self.blk):
            for i in range(reps):
                C.memmove(h + i * a.nbytes, a.ctypes.data, a.nbytes)
        self.host_sizes = sizes

    def step_e2e(self, k):
        L = self.L
        if self.host is None:
            self._host_alloc()
        for _ in range(k):
            for d, h, s in zip(self.dev, self.host[:5], self.host_sizes[:5]):
                L.ppo_b200_h2d(d.ptr, h, s)
            self._gae()
            L.ppo_b200_d2h(self.host[5], self.adv.ptr, 4 * self.n)
            L.ppo_b200_d2h(self.host[6], self.tgt.ptr, 4 * self.n)

    def units_per_step(self):
        return self.n

    def e2e_bytes(self):
        return 14 * self.n, 8 * self.n

    def config(self):
        return {"workload": "c5: GAE/returns + advantage normalisation on synthetic rewards/values/dones, T=%d x N=%d per GPU "
                            "(BASELINE.json configs[4])" % (self.T, self.N),
                "elements_per_step_per_gpu": self.n, "parallelism": "env-sharded x%d, no collective in the scan" % self.world,
                "l2": "inputs larger than L2: 3.9 GB read + 1.1 GB written per step",
                "e2e_call": "ppo_b200_gae on device staging of pinned HOST arrays: 5 H2D copies (14 B/element) + scan + "
                            "normalise + 2 D2H copies (8 B/element)"}

    def roofline_work(self, kernels):
        return {"gae_scan": ("hbm", 22.0 * self.n), "gae_normalize_kernel": ("hbm", 8.0 * self.n)}

    def teardown(self):
        for d in self.dev + [self.adv, self.tgt, self.st]:
            d.free()
        if self.host:
            for h in self.host:
                self.L.ppo_b200_free_host(h)

    def cpu_sample(self, n_iters, seed=1):
        import oracle
        blk = self._block(np.random.default_rng(seed), self.CPU_N)
        times = []
        for _ in range(n_iters):
            t0 = time.perf_counter()
            oracle.gae(*blk, 0.99, 0.95)
            times.append(time.perf_counter() - t0)
        return times, self.T * self.CPU_N

    def cpu_sample_desc(self):
        return ("oracle port of compute_gae's delta / recurrence / returns / normalisation (src/ppo.cu:338-368) on a "
                "T=%d x N=%d sample (%d elements)" % (self.T, self.CPU_N, self.T * self.CPU_N))


class StageWorkload(Workload):
    """Single-kernel stage benchmark on flat synthetic arrays (HBM roofline evidence for north-star items d, g)."""
    host_arrays = ()

    def _mk(self, shape, dtype=f32, rng=None, scale=1.0):
        import b200
        a = (rng.standard_normal(shape, dtype=f32) * scale).astype(dtype) if rng is not None else np.zeros(shape, dtype)
        return b200.dev(a), a

    def step_device(self, k):
        for _ in range(k):
            self._call()

    def teardown(self):
        for d in self.devs:
            d.free()


class AdamStage(StageWorkload):
    """Adam over one flat fp32 vector of 2^26 parameters (src/adam.cu:53-74): 28 B/parameter."""
    name = "adam"
    metric, unit = "adam_params_per_s", "params/s"
    P = 1 << 26
    CPU_P = 1 << 22

    def setup(self):
        rng = np.random.default_rng(3 + self.rank)
        (self.w, hw), (self.g, hg) = self._mk(self.P, rng=rng, scale=0.1), self._mk(self.P, rng=rng, scale=0.01)
        (self.m, _), (self.v, _) = self._mk(self.P), self._mk(self.P)
        self.devs = [self.w, self.g, self.m, self.v]
        self.hw, self.hg, self.t = hw, hg, 0

    def _call(self):
        self.t += 1
        self.L.ppo_b200_adam_flat(self.w.ptr, self.g.ptr, self.m.ptr, self.v.ptr, self.P, 3e-4, 0.9, 0.999, self.t)

    def step_e2e(self, k):
        for _ in range(k):          # gradient arrives from the host, updated weights go back
            self.L.ppo_b200_h2d(self.g.ptr, self.hg.ctypes.data, self.hg.nbytes)
            self._call()
            self.L.ppo_b200_d2h(self.hw.ctypes.data, self.w.ptr, self.hw.nbytes)

    def units_per_step(self):
        return self.P

    def e2e_bytes(self):
        return 4 * self.P, 4 * self.P

    def config(self):
        return {"workload": "adam: one multi-tensor Adam launch over a flat vector of %d fp32 parameters (g, m, v, w resident; 1.07 GB "
                            "of state, larger than L2)" % self.P, "parallelism": "replicated x%d" % self.world,
                "l2": "inputs larger than L2 (4 x 268 MB)", "e2e_call": "ppo_b200_adam_flat with the gradient uploaded from and the "
                "weights downloaded to pageable host memory each step"}

    def roofline_work(self, kernels):
        return {"adam_flat_kernel": ("hbm", 28.0 * self.P)}

    def cpu_sample(self, n_iters, seed=1):
        import oracle
        rng = np.random.default_rng(seed)
        w, g = rng.standard_normal(self.CPU_P, dtype=f32), rng.standard_normal(self.CPU_P, dtype=f32)
        m, v = np.zeros(self.CPU_P, f32), np.zeros(self.CPU_P, f32)
        times, t = [], 0
        for _ in range(n_iters):
            t0 = time.perf_counter()
            t = oracle.adam(w, g, m, v, 3e-4, t)
            times.append(time.perf_counter() - t0)
        return times, self.CPU_P

    def cpu_sample_desc(self):
        return "oracle port of adam_update (src/adam.cu:53-74) over %d parameters" % self.CPU_P


class GatherStage(StageWorkload):
    """Minibatch gather by permutation on the C3-shaped buffer (src/trajectory_buffer.cu:168-220): 212 B/sample.
    One step = what one PPO iteration of the layer-wise update path does: ONE streaming pass that builds the row-packed mirror
    (csrc/buffer.cu) + one gather of every row per epoch, 14 epochs (10 value + 4 policy, src/main.c:36-37)."""
    name = "gather"
    metric, unit = "gather_samples_per_s", "samples/s"
    S, A, B = 17, 6, 2048 * 512
    EPOCHS = 14
    PACKED = True
    CPU_B = 1 << 18

    def setup(self):
        import b200
        rng = np.random.default_rng(5 + self.rank)
        S, A, B = self.S, self.A, self.B
        self.src = [self._mk((B, S), rng=rng)[0], self._mk((B, A), rng=rng)[0]] + [self._mk(B, rng=rng)[0] for _ in range(3)]
        self.dst = [self._mk((B, S))[0], self._mk((B, A))[0]] + [self._mk(B)[0] for _ in range(3)]
        self.idx = [b200.dev_empty(B, np.int32) for _ in range(self.EPOCHS)]
        for e, d in enumerate(self.idx):
            self.L.ppo_b200_permutation(d.ptr, B, 17, e)
        self.packed = b200.dev_empty((B, self.L.ppo_b200_packed_row_floats(S, A)))
        self.devs = self.src + self.dst + self.idx + [self.packed]

    def _call(self):
        L = self.L
        if self.PACKED:
            L.ppo_b200_pack_rows(self.packed.ptr, self.B, self.S, self.A, *[d.ptr for d in self.src])
        for e in range(self.EPOCHS):
            if self.PACKED:
                L.ppo_b200_gather_packed(self.idx[e].ptr, 0, self.B, self.B, self.S, self.A, self.packed.ptr, *[d.ptr for d in self.dst])
            else:
                L.ppo_b200_gather(self.idx[e].ptr, 0, self.B, self.B, self.S, self.A, *[d.ptr for d in self.src], *[d.ptr for d in self.dst])

    def step_e2e(self, k):
        out = [np.empty(d.shape, f32) for d in self.dst]
        for _ in range(k):
            self._call()
            for d, o in zip(self.dst, out):
                self.L.ppo_b200_d2h(o.ctypes.data, d.ptr, o.nbytes)

    def units_per_step(self):
        return self.B * self.EPOCHS

    def e2e_bytes(self):
        return 0, 4 * self.B * (self.S + self.A + 3)

    def config(self):
        how = ("from the row-packed mirror (one contiguous 128-byte row per sample; the mirror is rebuilt once per step)" if self.PACKED
               else "from the five SoA arrays (random 68-byte state rows, 24-byte action rows, three 4-byte scalars)")
        return {"workload": "%s: all %d rows of a HalfCheetah-shaped buffer (S=17, A=6) gathered by %d device permutations %s"
                            % (self.name, self.B, self.EPOCHS, how),
                "parallelism": "replicated x%d" % self.world, "l2": "buffer 109 MB (+ 134 MB mirror) + 109 MB output; random rows",
                "e2e_call": "pack + %d gathers on the resident buffer + download of the last gathered minibatch" % self.EPOCHS}

    def roofline_work(self, kernels):
        per = float(self.B) * (4 + 2 * 4 * (self.S + self.A + 3))
        return {"gather_packed_kernel": ("hbm", self.EPOCHS * per), "gather_kernel": ("hbm", self.EPOCHS * per),
                "pack_rows_kernel": ("hbm", float(self.B) * 4 * ((self.S + self.A + 3) + 32))}

    def cpu_sample(self, n_iters, seed=1):
        import oracle
        import cabi
        rng = np.random.default_rng(seed)
        n = self.CPU_B
        st, ac = rng.standard_normal((n, self.S), dtype=f32), rng.standard_normal((n, self.A), dtype=f32)
        lp, ad, at = (rng.standard_normal(n, dtype=f32) for _ in range(3))
        cabi.srand(seed)
        idx = oracle.shuffle(n)
        times = []
        for _ in range(n_iters):
            t0 = time.perf_counter()
            oracle.get_batch(idx, 0, n, st, ac, lp, ad, at)
            times.append(time.perf_counter() - t0)
        return times, n

    def cpu_sample_desc(self):
        return "oracle port of get_batch (src/trajectory_buffer.cu:202-220) over %d rows" % self.CPU_B


class GatherSoA(GatherStage):
    """The same gathers straight from the SoA arrays (the reference's layout; get_batch_cuda / ppo_b200_gather)."""
    name = "gather_soa"
    PACKED = False


WORKLOADS = {"c1": C1, "c2": C2, "c3": C3, "c3x3": C3X3, "c4": C4, "c4bf16": C4BF16, "c5": C5, "adam": AdamStage, "gather": GatherStage, "gather_soa": GatherSoA}


# ======================================================================================= CPU arm
def _cpu_worker(args):
    name, n_iters, seed = args
    wl = WORKLOADS[name](None, 0, 1)
    return wl.cpu_sample(n_iters, seed)


CPU_BUILD = ("oracle port, timing-only build: gcc -O3 -march=native -ffast-math, GEMM loops in vectorisable order "
             "(`make -C oracle fast`, compiled on this host); the parity tests use the -O2 -ffp-contract=off build")


def use_fast_oracle():
    import oracle
    oracle.use_fast_build()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cls = WORKLOADS[args.workload]
    kind = getattr(cls, "ref_kind", "port")
    if kind == "port":
        use_fast_oracle()                 # before the fork: the workers inherit the selected build
    cores = max(1, min(len(os.sched_getaffinity(0)), 128))
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(args.workload, args.warmup + args.steps, 1 + i) for i in range(cores)])
    units = res[0][1]
    # every worker ran its own replica; a "step" of the arm = all replicas doing one sample step
    per_step = [max(r[0][args.warmup + k] for r in res) for k in range(args.steps)]
    total = sum(per_step)
    value = cores * units * args.steps / total
    wl = cls(None, 0, args.gpus)
    line = {
        "impl": "reference", "metric": cls.metric, "value": value, "unit": cls.unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": value, "unit": cls.unit, "cores": cores, "kind": kind,
                         "sample": wl.cpu_sample_desc() + "; %d independent replicas, one per host core" % cores,
                         "build": CPU_BUILD if kind == "port" else "unmodified reference sources, nvcc -O3 (oracle/Makefile)"},
        "e2e": {"value": value, "unit": cls.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ======================================================================================= helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; mark() brackets the timed region."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.lo = self.hi = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def _lines(self):
        try:
            return sum(1 for _ in open(self.path))
        except OSError:
            return 0

    def wait_first_sample(self, timeout=3.0):
        t0 = time.time()
        while self.proc and self._lines() == 0 and time.time() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        if self.lo is None:
            self.lo = self._lines()
        else:
            self.hi = self._lines()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    rows.append((float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except OSError:
            pass
        window = "timed region"
        sel = rows[self.lo:(self.hi + 1 if self.hi is not None else None)] if self.lo is not None else rows
        if not sel:                       # timed region shorter than one sampling period: use every sample of the run
            sel, window = rows, "warm-up + timed region (timed region shorter than the 50 ms sampling period)"
        if sel:
            out = {"sm_mhz": statistics.median(r[0] for r in sel), "sm_max_mhz": max(r[1] for r in sel),
                   "reasons": sorted({n for r in sel for n in r[2]}), "samples": len(sel), "window": window}
        return out


def build_roofline(kernels, work, traffic_file=None):
    """Algorithmic work of one step per kernel over its CUDA-event time in the profiled step.
    `work` maps a kernel-name prefix to (bound, amount): bytes for "hbm", FLOPs otherwise."""
    pk = load_peaks()
    agg = {}
    for name, k in kernels.items():
        for prefix in work:                     # a work key matches a kernel when it is a substring of its name
            if prefix in name.replace(" ", ""):
                a = agg.setdefault(prefix, {"launches": 0, "total_ms": 0.0})
                a["launches"] += k["launches"]
                a["total_ms"] += k["total_ms"]
    out = {}
    for prefix, (bound, amount) in work.items():
        if prefix not in agg or agg[prefix]["total_ms"] <= 0:
            continue
        sec = agg[prefix]["total_ms"] * 1e-3
        if bound == "hbm":
            ach, peak, unit, psrc = amount / sec / 1e9, pk["hbm"], "GB/s", pk["src"] + ": HBM copy read+write"
        elif bound == "tensor_bf16":
            ach, peak, unit = amount / sec / 1e12, pk["bf16"], "TFLOP/s"
            psrc = pk["src"] + ": sustained cuBLAS bf16"
            bound = "tensor"
        elif bound == "tensor_tf32":
            ach, peak, unit = amount / sec / 1e12, pk["bf16"] / 2, "TFLOP/s"
            psrc = pk["src"] + ": sustained cuBLAS bf16 / 2 (TF32 dense rate is half of bf16)"
            bound = "tensor"
        elif bound == "tensor_3xtf32":
            # fp32-equivalent FLOPs (2 m n k per contraction); every product costs three TF32 MMAs, so the tensor-pipe peak for
            # this mode is a third of the TF32 rate
            ach, peak, unit = amount / sec / 1e12, pk["bf16"] / 6, "TFLOP/s"
            psrc = pk["src"] + ": sustained cuBLAS bf16 / 2 (TF32) / 3 (three MMAs per product in the 3xTF32 split mode)"
            bound = "tensor"
        else:
            ach, unit = amount / sec / 1e12, "TFLOP/s"
            peak = FP32_MEASURED.get("ffma_const_operands") or FP32_PEAK_TFLOPS
            psrc = ("measured in this run (csrc/ubench.cu): FFMA stream with constant-bank operands = the fp32 pipe's peak"
                    if FP32_MEASURED else "nominal fp32 FFMA: 148 SMs x 128 lanes x 2 x 1.965 GHz")
        out[prefix] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                       "launches": agg[prefix]["launches"], "avg_us": 1e3 * agg[prefix]["total_ms"] / agg[prefix]["launches"],
                       "peak_source": psrc}
        if bound == "fp32":
            out[prefix]["peak_nominal"] = FP32_PEAK_TFLOPS
            if FP32_MEASURED.get("ffma2_register_operands"):
                # what a GEMM-like inner loop with three REGISTER operands can issue (weights change every minibatch)
                out[prefix]["peak_register_operand_ffma2"] = FP32_MEASURED["ffma2_register_operands"]
                out[prefix]["frac_of_register_operand_ceiling"] = ach / FP32_MEASURED["ffma2_register_operands"]
    total = sum(k["total_ms"] for k in kernels.values())
    if not out:
        return None
    dom = max(out, key=lambda p: agg[p]["total_ms"])
    r = dict(out[dom])
    traffic = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ran = [name.replace(" ", "") for name in kernels if dom in name.replace(" ", "")]
        for key, val in tj.items():
            if dom in key and any(key in name for name in ran):
                traffic = val.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    r.update({"kernel": dom, "share_of_step": agg[dom]["total_ms"] / total if total > 0 else None, "traffic": traffic,
              "sum_of_kernel_ms": total, "per_kernel": out})
    return r


# ======================================================================================= GPU arm
def run_gpu_arm(args):
    import torch
    import b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    L = b200.lib()
    L.ppo_b200_set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        b200.package().dist_init_from_torch(L)
        L.ppo_b200_dist_set_shard_mode(1)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    L.ppo_b200_set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0 and not args.only_value:
        FP32_MEASURED["ffma_const_operands"] = L.ppo_b200_measure_fp32_peak(0)
        FP32_MEASURED["ffma_register_operands"] = L.ppo_b200_measure_fp32_peak(1)
        FP32_MEASURED["ffma2_register_operands"] = L.ppo_b200_measure_fp32_peak(2)
    wl = WORKLOADS[args.workload](L, rank, world)
    if args.mb > 0 and hasattr(wl, "MB"):
        wl.MB = args.mb
    wl.setup()

    # ---- value: inputs resident in HBM ------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first_sample()
    wl.step_device(args.warmup)
    barrier()
    if sampler:
        sampler.mark()
    launches0 = L.ppo_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    wl.step_device(args.steps)
    e1.record(stream)
    barrier()
    if sampler:
        sampler.mark()
    ms = e0.elapsed_time(e1)
    launches = L.ppo_b200_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    extra = wl.extra(ms / args.steps) if rank == 0 else {}

    if args.only_value:
        if rank == 0:
            emit({"only_value": True, "workload": args.workload, "ms_per_step": ms / args.steps, "gpu_launches": int(launches)})
        wl.teardown()
        if world > 1:
            L.ppo_b200_dist_finalize()
            dist.destroy_process_group()
        return

    # ---- e2e: the host-buffer call, copies inside the timed region -----------------------------------
    e2e_steps = args.steps if args.e2e_steps <= 0 else args.e2e_steps
    wl.step_e2e(1)                                       # warm the pinned mirrors
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    wl.step_e2e(e2e_steps)
    f1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(f0.elapsed_time(f1), 0.0)
    api_ms = None
    if hasattr(wl, "step_api"):
        wl.step_api(1)
        barrier()
        ta = time.perf_counter()
        wl.step_api(e2e_steps)
        barrier()
        api_ms = 1e3 * (time.perf_counter() - ta)
    if world > 1:
        if api_ms is not None:
            ta_t = torch.tensor([api_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(ta_t, op=dist.ReduceOp.MAX)
            api_ms = float(ta_t.cpu()[0])
        tt = torch.tensor([ms, e2e_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms, wall_ms = (float(x) for x in tt.cpu())
    e2e_ms = max(e2e_ms, wall_ms)                        # host-blocking copies: the wall clock is the honest one

    # ---- per-kernel timing of one profiled step (rank 0; other ranks run the same step unprofiled) -----
    kernels, roofline = {}, None
    if rank == 0:
        L.ppo_b200_profile_begin()
    wl.step_device(1)
    if rank == 0:
        buf = C.create_string_buffer(1 << 16)
        L.ppo_b200_profile_end(buf, len(buf))
        for ln in buf.value.decode().splitlines():
            name, cnt, tot = ln.rsplit(" ", 2)
            kernels[name] = {"launches": int(cnt), "total_ms": float(tot)}
        roofline = build_roofline(kernels, wl.roofline_work(kernels))
    barrier()

    units = world * wl.units_per_step()
    h2d, d2h = wl.e2e_bytes()
    cfg = wl.config()
    wl.teardown()

    # ---- the other named shapes, short runs, so that ONE default invocation carries every BASELINE.json config -----------
    secondary = None
    if args.workload == "c2" and not args.no_secondary:
        secondary = run_secondary(L, rank, world, stream, barrier, torch, dist)

    if rank == 0:
        line = {
            "metric": wl.metric, "value": units * args.steps / (ms * 1e-3), "unit": wl.unit,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": cfg,
            "e2e": {"value": units * e2e_steps / (e2e_ms * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
            "api": None if api_ms is None else {"value": units * e2e_steps / (api_ms * 1e-3), "unit": wl.unit,
                                                 "ms_per_step": api_ms / e2e_steps, "call": "train_ppo_epoch"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(extra)
        line["roofline"] = roofline
        line["kernels"] = kernels
        line["fp32_peaks_measured_tflops"] = dict(FP32_MEASURED, nominal=FP32_PEAK_TFLOPS)
        if world == 1:
            kind = getattr(wl, "ref_kind", "port")
            parity_value = None
            if kind == "port":
                t_par, u_par = wl.cpu_sample(2)              # the bit-exact parity build, for the record
                parity_value = u_par / t_par[-1]
                use_fast_oracle()
            times, cpu_units = wl.cpu_sample(args.cpu_iters + 1)
            times = times[1:]
            line["cpu_baseline"] = {"value": cpu_units / statistics.mean(times), "unit": wl.unit, "cores": 1, "kind": kind,
                                    "sample": wl.cpu_sample_desc() + "; %d timed steps, one thread (the reference is "
                                              "single-threaded, src/main.c:18)" % len(times),
                                    "build": CPU_BUILD if kind == "port" else "unmodified reference sources, nvcc -O3 (oracle/Makefile)"}
            if parity_value is not None:
                line["cpu_baseline"]["parity_build_value"] = parity_value
        if secondary:
            line["secondary"] = secondary
        emit(line)
    if world > 1:
        L.ppo_b200_dist_finalize()
        dist.destroy_process_group()


def run_secondary(L, rank, world, stream, barrier, torch, dist):
    """Short runs of the other BASELINE.json configs after the headline workload: value (device-resident, CUDA events, max over
    ranks) + the roofline of the dominant kernel from one profiled step.  c1 (host env, not collective) only at N = 1."""
    out = {}
    for name, steps in (("c3", 2), ("c3x3", 2), ("c4", 2), ("c4bf16", 2), ("c5", 5), ("gather", 3)):
        wl = WORKLOADS[name](L, rank, world)
        wl.setup()
        wl.step_device(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        wl.step_device(steps)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.cpu()[0])
        roof = None
        if rank == 0:
            L.ppo_b200_profile_begin()
        wl.step_device(1)
        if rank == 0:
            buf = C.create_string_buffer(1 << 16)
            L.ppo_b200_profile_end(buf, len(buf))
            kernels = {}
            for ln in buf.value.decode().splitlines():
                kname, cnt, tot = ln.rsplit(" ", 2)
                kernels[kname] = {"launches": int(cnt), "total_ms": float(tot)}
            r = build_roofline(kernels, wl.roofline_work(kernels))
            if r:
                roof = {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "share_of_step", "peak_source")}
                roof["whole_step_frac"] = whole_step_fraction(wl, ms / steps)
        barrier()
        out[name] = {"metric": wl.metric, "value": world * wl.units_per_step() * steps / (ms * 1e-3), "unit": wl.unit, "n_gpus": world,
                     "steps": steps, "warmup": 1, "ms_per_step": ms / steps, "dtype": wl.dtype, "config": wl.config()["workload"],
                     "roofline": roof}
        wl.teardown()
    if world == 1:
        wl = WORKLOADS["c1"](L, rank, world)
        wl.setup()
        wl.step_device(1)
        L.ppo_b200_sync()
        t0 = time.perf_counter()
        wl.step_device(2)
        L.ppo_b200_sync()
        sec = (time.perf_counter() - t0) / 2
        wl.teardown()
        c1 = {"metric": wl.metric, "value": wl.STEPS / sec, "unit": wl.unit, "n_gpus": 1, "steps": 2, "warmup": 1,
              "s_per_epoch": sec, "config": wl.config()["workload"], "timing": "wall clock around train_ppo_epoch (host-driven env loop)"}
        try:
            t_blas, _ = wl.cpu_sample(2, blas=True)
            c1["reference_cpu"] = {"value": wl.STEPS / t_blas[-1], "s_per_epoch": t_blas[-1], "kind": "reference", "cores": 1, "blas": wl.cpu_blas}
            if "OpenBLAS" in wl.cpu_blas:
                t_nv, _ = wl.cpu_sample(1, blas=False)
                c1["reference_cpu_naive_cblas"] = {"value": wl.STEPS / t_nv[-1], "s_per_epoch": t_nv[-1], "kind": "reference", "cores": 1,
                                                   "blas": wl.cpu_blas}
        except Exception as exc:                                    # the reference .so did not travel / cannot load
            c1["reference_cpu"] = {"unavailable": repr(exc)}
        out["c1"] = c1
    return out


def whole_step_fraction(wl, ms_per_step):
    """Whole-step arithmetic rate of a GEMM-bound workload against its peak (c4: TF32 tensor; c3: fp32), or None."""
    if not hasattr(wl, "SIZES") or not hasattr(wl, "N_VAL") or not hasattr(wl, "cap"):
        return None
    nb = wl.cap // wl.MB
    sv = wl.SIZES[:-1] + [1]
    flops = (wl.N_VAL * nb * train_flops(sv, wl.MB) + wl.N_POL * nb * train_flops(wl.SIZES, wl.MB) + 2 * 2 * wl.cap * mlp_weights(sv))
    tfs = flops / (ms_per_step * 1e-3) / 1e12
    pk = load_peaks()
    prec = getattr(wl, "TF32", 0)
    peak = (pk["bf16"] if prec == 2 else pk["bf16"] / 6 if prec == 3 else pk["bf16"] / 2 if prec
            else (FP32_MEASURED.get("ffma_const_operands") or FP32_PEAK_TFLOPS))
    return {"tflops": tfs, "peak": peak, "frac": tfs / peak}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints lands on stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    # NCCL (NCCL_DEBUG=VERSION/INFO), torchrun banners etc. write to fd 1: keep stdout clean for the JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (default: --steps)")
    ap.add_argument("--cpu-iters", type=int, default=3)
    ap.add_argument("--mb", type=int, default=0, help="override the workload's minibatch size per GPU (c2/c3/c4)")
    ap.add_argument("--no-secondary", action="store_true", help="c2 only: skip the short c3 / c4 / c5 / c1 runs that fill `secondary`")
    ap.add_argument("--only-value", action="store_true",
                    help="skip the e2e / per-kernel / CPU legs (short command for ncu passes; not a bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
