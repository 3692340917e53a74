"""PPO_B200_PHASE_DEBUG=1 python scripts/phase_debug.py [mb] — globaltimer stamps of the LAST minibatch of the last
fused_phase_kernel launch (persistent phase kernel, csrc/fused_mlp.cu), per CTA:
  3 tile start (gather complete) | 4 tiles done [A] | 5 past grid barrier [B] | 6 slice reduce + Adam done [C] | 7 (previous
  step) past grid barrier [D] + image re-staged [E]."""
import ctypes as C, os, sys
os.environ["PPO_B200_PHASE_DEBUG"] = "1"
sys.path.insert(0, "tests")
import numpy as np, b200, cabi
L = b200.lib()
L.ppo_b200_set_device(0)
cabi.srand(1)
N, T = 4096, 200
MB = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
env = L.create_pendulum_env_cuda(N, 1)
ppo = L.create_ppo(cabi.cstr_array(["tanh", "tanh", "none"]), cabi.int_array([3, 64, 64, 1]), 4, N * T, 3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
L.ppo_b200_set_permutation_mode(ppo, 1, 7)
L.ppo_b200_train_iterations(ppo, env, 2, MB, 4, 10)
L.ppo_b200_sync()
grid = min(148, -(-MB // 128)) if MB > 148 * 64 else -(-MB // 64)
out = np.zeros((grid, 16), np.uint64)
L.ppo_b200_debug_phase_stamps.argtypes = [C.c_void_p, C.c_int]
L.ppo_b200_debug_phase_stamps(out.ctypes.data, grid)
t = out.astype(np.int64)
us = lambda a, b: (t[:, a] - t[:, b]) / 1e3
print("CTAs", grid, "minibatch", MB)
for name, d in [("[E]->tile start (gather wait)", us(3, 7)), ("[A] tiles (fwd+head+bwd)", us(4, 3)), ("[B] grid barrier wait", us(5, 4)),
                ("[C] slice reduce + Adam", us(6, 5))]:
    print("%-34s median %.2f us  min %.2f  max %.2f" % (name, np.median(d), d.min(), d.max()))
for name, a, b in [("  fwd L0", 8, 3), ("  fwd L1", 9, 8), ("  last layer + head (+dX)", 11, 9), ("  bwd l=1 (dW | dX)", 13, 11),
                   ("  bwd l=0", 12, 13), ("  next gather issue -> [A] end", 4, 12)]:
    d = us(a, b)
    print("%-34s median %.2f us  min %.2f  max %.2f" % (name, np.median(d), d.min(), d.max()))
if t[:, 10].any():      # specialised kernel: finer stamps
    for name, a, b in [("  spec64: head (128 thr)", 10, 9), ("  spec64: dPre2 + dW2 + db1", 11, 10), ("  spec64: dX1 loop + epilogue (warps 0-3)", 14, 11),
                       ("  spec64: dW0/db0 + loss sums (warps 0-3)", 15, 14), ("  spec64: dW1 loop (warps 4-7)", 1, 11),
                       ("  spec64: bwd start -> end barrier", 13, 11), ("  spec64: dW1 add + store", 12, 13)]:
        d = us(a, b)
        print("%-34s median %.2f us  min %.2f  max %.2f" % (name, np.median(d), d.min(), d.max()))
print("step period estimate (t6 last - t7 prev): median %.2f us" % np.median(us(6, 7)))
