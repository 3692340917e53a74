#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native PPO training path (contract: see DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5] [--impl reference]

Default workload = BASELINE.json configs[1] ("c2"): Pendulum-v1 PPO with 4096 vectorised device
envs per GPU, 2x64 tanh actor-critic, fp32.  One "step" = one PPO iteration = a 200-step rollout of
every env (819 200 env-steps per GPU) + GAE + 10 value epochs + 4 policy epochs of minibatch 16 384
(the reference's schedule, src/main.c:33-43, at a vectorised minibatch size).

  value : env-steps/s (rollout+update), whole job, device-resident (ppo_b200_train_iterations)
  e2e   : the same through the reference-facing C-ABI call train_ppo_epoch(), which also refreshes
          every host mirror (buffer, policy, V) each iteration like src/ppo.cu:536-538
  roofline / kernels : per-kernel CUDA-event timing of one profiled step (ppo_b200_profile_*)
  cpu_baseline : the plain-C restatement of the reference path (oracle/, "port") on host cores

N > 1: launched by torchrun, one process per GPU, weak scaling (4096 envs per GPU, the global
minibatch is the union of the rank-local ones), NCCL all-reduce of the flat gradients.
`--impl reference` times the CPU path only (rank 0), on all host cores as independent replicas.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIZES = [3, 64, 64, 1]
ACTS = ["tanh", "tanh", "none"]
N_ENVS, T = 4096, 200
MB, N_POL, N_VAL = 16384, 4, 10
CPU_SAMPLE_STEPS, CPU_SAMPLE_MB = 8200, 2050     # 41 episodes of 200 steps, 4 minibatches per epoch


# --------------------------------------------------------------------------------------- CPU arm
def cpu_iteration_seconds(n_iters, seed=1):
    """Plain-C path (oracle port of src/ppo.cu:373-448 + C Pendulum), single thread.  Returns the list
    of per-iteration wall times; one iteration = CPU_SAMPLE_STEPS env-steps (rollout + update)."""
    import cabi
    import oracle
    cabi.srand(seed)
    tr = oracle.Trainer(SIZES, ACTS, batch_size=CPU_SAMPLE_MB, n_epochs_policy=N_POL, n_epochs_value=N_VAL)
    buf = tr.make_buffer(CPU_SAMPLE_STEPS)
    times = []
    for _ in range(n_iters):
        t0 = time.perf_counter()
        tr.collect(buf, CPU_SAMPLE_STEPS, 1)
        tr.update(buf)
        times.append(time.perf_counter() - t0)
    return times


def _cpu_worker(args):
    n_iters, seed = args
    return cpu_iteration_seconds(n_iters, seed)


def cpu_sample_desc():
    return ("oracle port of the reference plain-C path: 1 env Pendulum, %d env-steps per step "
            "(rollout + GAE + %d value / %d policy epochs, minibatch %d), 2x64 tanh"
            % (CPU_SAMPLE_STEPS, N_VAL, N_POL, CPU_SAMPLE_MB))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    cores = max(1, min(cores, 128))
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(args.warmup + args.steps, 1 + i) for i in range(cores)])
    # every worker ran its own replica; a "step" of the arm = all replicas doing one iteration
    per_step = [max(r[args.warmup + k] for r in res) for k in range(args.steps)]
    total = sum(per_step)
    value = cores * CPU_SAMPLE_STEPS * args.steps / total
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": cpu_sample_desc() + "; %d independent replicas, one per host core" % cores},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- helpers
def workload_config(n_gpus):
    return {"workload": "c2: Pendulum-v1 PPO, %d device envs/GPU x T=%d, 2x64 tanh MLP, fp32, minibatch %d/GPU, "
                        "%d value + %d policy epochs (BASELINE.json configs[1])" % (N_ENVS, T, MB, N_VAL, N_POL),
            "env_steps_per_step_per_gpu": N_ENVS * T, "parallelism": "dp%d" % n_gpus,
            "l2": "working set > L2: each step streams the 819200-row buffer (38 MB) and 420 MB of V-net activations",
            "permutation": "device (auto mode after a device rollout)"}


class ClockSampler:
    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons)}
        return out


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except (OSError, KeyError, ValueError):
        return 6650.0, 1400.0, "fallback"


def mlp_weights(sizes):
    return sum(sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1))


# --------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import b200
    import cabi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    L = b200.lib()
    L.ppo_b200_set_device(local_rank)
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

    cabi.srand(1234)                                  # same initial weights on every rank
    env = L.create_pendulum_env_cuda(N_ENVS, 100 + rank)
    cap = N_ENVS * T
    ppo = L.create_ppo(cabi.cstr_array(ACTS), cabi.int_array(SIZES), len(SIZES), cap, 3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
    L.ppo_b200_set_permutation_mode(ppo, -1, 7 + rank)

    # ---- value: device-resident iterations ------------------------------------------------------
    L.ppo_b200_train_iterations(ppo, env, args.warmup, MB, N_POL, N_VAL)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = L.ppo_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    L.ppo_b200_train_iterations(ppo, env, args.steps, MB, N_POL, N_VAL)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.ppo_b200_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    mean_return = L.ppo_b200_last_mean_return(ppo)

    # ---- e2e: the reference-facing call, host mirrors refreshed every iteration --------------------
    L.train_ppo_epoch(ppo, env, cap, MB, N_POL, N_VAL)      # warm the pinned mirrors
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    L.train_ppo_epoch(ppo, env, cap * args.steps, MB, N_POL, N_VAL)
    f1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(f0.elapsed_time(f1), 0.0)
    S, A = SIZES[0], SIZES[-1]
    p_mu = mlp_weights(SIZES) + sum(SIZES[1:])
    p_v = p_mu - (SIZES[-2] + 1) * (A - 1)
    d2h = cap * (4 * (2 * S + A + 4) + 2) + 4 * (p_mu + p_v + A)

    if world > 1:
        tt = torch.tensor([ms, e2e_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms, wall_ms = (float(x) for x in tt.cpu())

    # ---- per-kernel timing of one profiled step (rank 0) ----------------------------------------------
    kernels, roofline = {}, None
    if rank == 0:
        L.ppo_b200_profile_begin()
        L.ppo_b200_train_iterations(ppo, env, 1, MB, N_POL, N_VAL)
        buf = C.create_string_buffer(1 << 16)
        L.ppo_b200_profile_end(buf, len(buf))
        for ln in buf.value.decode().splitlines():
            name, cnt, tot = ln.rsplit(" ", 2)
            kernels[name] = {"launches": int(cnt), "total_ms": float(tot)}
        roofline = build_roofline(kernels)
    barrier()

    if rank == 0:
        cpu_times = cpu_iteration_seconds(4)[1:] if world == 1 else None
        steps_total = world * cap * args.steps
        line = {
            "metric": "env_steps_per_s", "value": steps_total / (ms * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": steps_total / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                    "call": "train_ppo_epoch (reference API): rollout + update + buffer_to_host/policy_to_host/nn_write_weights_to_host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "update_samples_per_s": world * (N_POL + N_VAL) * (cap // MB) * MB * args.steps / (ms * 1e-3),
            "mean_episode_return": mean_return,
            "roofline": roofline,
            "kernels": kernels,
        }
        if cpu_times:
            line["cpu_baseline"] = {"value": CPU_SAMPLE_STEPS / statistics.mean(cpu_times), "unit": "env-steps/s", "cores": 1,
                                    "kind": "port", "sample": cpu_sample_desc() + "; %d timed steps" % len(cpu_times)}
        print(json.dumps(line), flush=True)
    L.free_ppo(ppo)
    env.contents.free_env()
    if world > 1:
        L.ppo_b200_dist_finalize()
        dist.destroy_process_group()


def build_roofline(kernels):
    """Algorithmic work of one c2 step per kernel class over its CUDA-event time."""
    hbm, tens, src = load_peaks()
    B, nb = N_ENVS * T, (N_ENVS * T) // MB
    steps = (N_POL + N_VAL) * nb
    W = mlp_weights(SIZES)
    P = W + sum(SIZES[1:])
    H = SIZES[1]
    # fp32 FLOPs through the tiled kernels per minibatch (skinny last layer of the forward excluded):
    fwd_tiled = 2 * MB * (SIZES[0] * H + H * H)                     # sgemm_kernel<kFwd>
    bwd_in = 2 * MB * (H * H + H * 1)                               # sgemm_kernel<kBwdInput>: layers 2,1
    bwd_par = 2 * MB * W                                            # sgemm_kernel<kBwdParam>
    gae_fwd = 2 * 2 * B * (SIZES[0] * H + H * H)                    # two V forwards over the buffer
    work = {
        "gae_scan_kernel": ("hbm", 22.0 * B),
        "gae_normalize_kernel": ("hbm", 8.0 * B),
        "adam_flat_kernel": ("hbm", 28.0 * (steps * P + N_POL * nb * 1)),
        "gather_kernel": ("hbm", steps * MB * (4 + 2 * 4 * (SIZES[0] + SIZES[-1] + 3))),
        "sgemm_kernel<kFwd>": ("fp32", steps * fwd_tiled + gae_fwd),
        "sgemm_kernel<kBwdInput>": ("fp32", steps * bwd_in),
        "sgemm_kernel<kBwdParam>": ("fp32", steps * bwd_par),
    }
    out = {}
    for name, (bound, amount) in work.items():
        if name not in kernels or kernels[name]["total_ms"] <= 0:
            continue
        sec = kernels[name]["total_ms"] * 1e-3
        if bound == "hbm":
            ach, peak, unit = amount / sec / 1e9, hbm, "GB/s"
        else:
            ach, peak, unit = amount / sec / 1e12, 148 * 128 * 2 * 1.965e9 / 1e12, "TFLOP/s"
        out[name] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                     "avg_us": 1e3 * kernels[name]["total_ms"] / kernels[name]["launches"]}
    total = sum(k["total_ms"] for k in kernels.values())
    dom = max(kernels, key=lambda k: kernels[k]["total_ms"])
    r = dict(out.get(dom, {"bound": "hbm", "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None}))
    r.update({"kernel": dom, "share_of_step": kernels[dom]["total_ms"] / total, "traffic": None,
              "peak_source": src + (" HBM copy" if r["bound"] == "hbm" else "; fp32 peak = 148 SMs x 128 FMA x 2 x 1.965 GHz (nominal, no measured fp32 figure)"),
              "per_kernel": out})
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
