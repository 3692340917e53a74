"""Update-phase time of the reference's DEFAULT width (2x128, src/main.c:20) on 4096 device envs: old vs tile kernel."""
import sys, time; sys.path.insert(0, "tests")
import b200, cabi
L = b200.lib(); L.ppo_b200_set_device(0)
cabi.srand(1)
N, T, MB = 4096, 200, 16384
env = L.create_pendulum_env_cuda(N, 1)
ppo = L.create_ppo(cabi.cstr_array(["relu", "relu", "none"]), cabi.int_array([3, 128, 128, 1]), 4, N * T, 3e-4, 3e-4, 0.95, 0.2, 0.0, 1.0, True)
L.ppo_b200_train_iterations(ppo, env, 2, MB, 4, 10); L.ppo_b200_sync()
t0 = time.perf_counter(); L.ppo_b200_train_iterations(ppo, env, 5, MB, 4, 10); L.ppo_b200_sync()
print("2x128 relu: %.2f ms per iteration, mean return %.1f" % (1e3 * (time.perf_counter() - t0) / 5, L.ppo_b200_last_mean_return(ppo)))
