"""PPO_B200_PHASE_DEBUG=1 python scripts/phase_debug.py — per-phase globaltimer stamps of fused_tile64_kernel."""
import ctypes as C, os, sys
os.environ["PPO_B200_PHASE_DEBUG"] = "1"
sys.path.insert(0, "tests")
import numpy as np, b200, cabi
L = b200.lib()
L.ppo_b200_set_device(0)
cabi.srand(1)
N, T, MB = 4096, 200, 16384
env = L.create_pendulum_env_cuda(N, 1)
ppo = L.create_ppo(cabi.cstr_array(["tanh", "tanh", "none"]), cabi.int_array([3, 64, 64, 1]), 4, N * T, 3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
L.ppo_b200_train_iterations(ppo, env, 2, MB, 4, 10)
L.ppo_b200_sync()
nb = MB // 64
out = np.zeros((nb, 16), np.uint64)
L.ppo_b200_debug_phase_stamps.argtypes = [C.c_void_p, C.c_int]
L.ppo_b200_debug_phase_stamps(out.ctypes.data, nb)
t = out.astype(np.int64)
t0 = t[:, 0].min()
names = {0: "start", 1: "gather done", 2: "dep wait done", 3: "image in smem", 4: "fwd L0", 5: "fwd L1", 6: "fwd L2", 9: "head done", 12: "bwd l=2", 11: "bwd l=1", 10: "bwd l=0 (end)"}
order = [0, 1, 2, 3, 4, 5, 6, 9, 12, 11, 10]
print("blocks", nb, "kernel span (first start -> last end): %.2f us" % ((t[:, 10].max() - t0) / 1e3))
print("block start spread: %.2f us" % ((t[:, 0].max() - t0) / 1e3))
prev = None
for k in order:
    rel = (t[:, k] - t[:, 0]) / 1e3
    d = "" if prev is None else "  delta median %.2f" % np.median((t[:, k] - t[:, prev]) / 1e3)
    print("%-16s since block start: median %.2f us  max %.2f%s" % (names[k], np.median(rel), rel.max(), d))
    prev = k
