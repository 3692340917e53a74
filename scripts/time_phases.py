"""python scripts/time_phases.py — wall-clock per iteration of the c2 shape with different (policy, value) epoch mixes."""
import os, sys, time
sys.path.insert(0, "tests")
import b200, cabi
L = b200.lib()
L.ppo_b200_set_device(0)
cabi.srand(1)
N, T, MB = 4096, 200, 18944
env = L.create_pendulum_env_cuda(N, 1)
ppo = L.create_ppo(cabi.cstr_array(["tanh", "tanh", "none"]), cabi.int_array([3, 64, 64, 1]), 4, N * T, 3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
L.ppo_b200_set_permutation_mode(ppo, 1, 7)
for npol, nval in [(4, 10), (14, 0), (0, 14), (0, 0)]:
    L.ppo_b200_train_iterations(ppo, env, 2, MB, npol, nval)
    L.ppo_b200_sync()
    t0 = time.perf_counter()
    L.ppo_b200_train_iterations(ppo, env, 10, MB, npol, nval)
    L.ppo_b200_sync()
    print("policy epochs %2d value epochs %2d: %.3f ms / iteration" % (npol, nval, (time.perf_counter() - t0) * 100))
