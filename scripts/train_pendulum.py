#!/usr/bin/env python
"""Learning-curve run of the c2 configuration (BASELINE.json configs[1]): Pendulum-v1, 4096 device envs x T=200,
2x64 MLP, the reference's hyper-parameters (src/main.c:33-43: lr 3e-4 both, lambda 0.95, eps 0.2, ent 0, init std 1,
4 policy / 10 value epochs, gamma 0.99) at minibatch 18944 (= 148 SMs x 2 CTAs x 64 rows).  Prints / writes the mean undiscounted episode return
per iteration (eval_ppo's "R", src/ppo.cu:581, over the 4096 training episodes of that iteration).

    python scripts/train_pendulum.py [--iters 80] [--act tanh|relu] [--obs-norm] [--mb 16384] [--out gpurun_out/learning_curve.json]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import b200, cabi

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=80)
ap.add_argument("--act", default="tanh")
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--T", type=int, default=200)
ap.add_argument("--mb", type=int, default=18944)
ap.add_argument("--lr", type=float, default=3e-4)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--obs-norm", action="store_true")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "learning_curve.json"))
args = ap.parse_args()

L = b200.lib()
L.ppo_b200_set_device(0)
cabi.srand(args.seed)
env = L.create_pendulum_env_cuda(args.envs, args.seed)
acts = [args.act, args.act, "none"]
ppo = L.create_ppo(cabi.cstr_array(acts), cabi.int_array([3, 64, 64, 1]), 4, args.envs * args.T, args.lr, args.lr, 0.95, 0.2, 0.0, 1.0, True)
if args.obs_norm:
    L.ppo_b200_set_obs_norm(ppo, 1)
curve, t0 = [], time.perf_counter()
solved_at = None
for it in range(args.iters):
    L.ppo_b200_train_iterations(ppo, env, 1, args.mb, 4, 10)
    r = L.ppo_b200_last_mean_return(ppo)          # return of the rollout that fed this iteration's update
    curve.append(r)
    if solved_at is None and r > -200:
        solved_at = it
    if it % 5 == 0 or it == args.iters - 1:
        print("iter %3d  env-steps %9d  mean return %9.1f  wall %.2fs" % (it, (it + 1) * args.envs * args.T, r, time.perf_counter() - t0), flush=True)
wall = time.perf_counter() - t0
res = {"config": vars(args), "mean_return_per_iteration": curve, "solved_threshold": -200, "first_iteration_above_threshold": solved_at,
       "env_steps_to_threshold": None if solved_at is None else (solved_at + 1) * args.envs * args.T,
       "best": max(curve), "final": curve[-1], "wall_s": wall, "env_steps_per_s_incl_readback": args.iters * args.envs * args.T / wall}
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(res, open(args.out, "w"), indent=1)
print("best %.1f final %.1f solved_at %s wall %.2fs" % (res["best"], res["final"], solved_at, wall))
L.free_ppo(ppo)
env.contents.free_env()
