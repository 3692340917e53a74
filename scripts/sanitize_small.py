"""Small end-to-end pass for compute-sanitizer (racecheck / memcheck): fused update on three net shapes, a short device
rollout with obs-norm, GAE on a ragged buffer, gather, Adam."""
import sys; sys.path.insert(0, "tests")
import numpy as np
import b200, cabi, oracle
from test_gpu_train import fill_host_buffer, make_ppo, synthetic_buffer
L = b200.lib(); L.ppo_b200_set_device(0)
for sizes, acts, n, mb in [([3, 64, 64, 1], ["tanh", "tanh", "none"], 512, 128), ([17, 32, 32, 6], ["relu", "relu", "none"], 384, 128),
                           ([3, 128, 128, 1], ["relu", "relu", "none"], 256, 128)]:
    cabi.srand(3)
    ppo = make_ppo(L, sizes, acts, n)
    T = oracle.Trainer(sizes, acts, batch_size=mb, n_epochs_policy=1, n_epochs_value=1, ref_index=False)
    b = synthetic_buffer(T, np.random.default_rng(1), sizes, acts, n)
    fill_host_buffer(ppo, b)
    L.ppo_b200_update(ppo, 0.99, mb, 1, 1)
    L.free_ppo(ppo)
env = L.create_pendulum_env_cuda(64, 3)
ppo = make_ppo(L, [3, 64, 64, 1], ["tanh", "tanh", "none"], 64 * 8)
L.ppo_b200_set_obs_norm(ppo, 1)
L.ppo_b200_train_iterations(ppo, env, 2, 128, 1, 1)
L.ppo_b200_sync()
L.free_ppo(ppo); env.contents.free_env()
n = 5000
rng = np.random.default_rng(0)
d = [b200.dev(rng.standard_normal(n).astype(np.float32)) for _ in range(3)] + [b200.dev((rng.random(n) < 0.01).astype(np.uint8)), b200.dev((rng.random(n) < 0.01).astype(np.uint8))]
adv, tgt, st = b200.dev_empty(n), b200.dev_empty(n), b200.dev_empty(2)
L.ppo_b200_gae(*[x.ptr for x in d], n, 0.99, 0.95, adv.ptr, tgt.ptr, 1, st.ptr)
n2 = 512 * 64 * 2
d2 = [b200.dev(rng.standard_normal(n2).astype(np.float32)) for _ in range(3)] + [b200.dev((rng.random(n2) < 0.01).astype(np.uint8)), b200.dev((rng.random(n2) < 0.01).astype(np.uint8))]
adv2, tgt2 = b200.dev_empty(n2), b200.dev_empty(n2)
L.ppo_b200_gae(*[x.ptr for x in d2], n2, 0.99, 0.95, adv2.ptr, tgt2.ptr, 1, st.ptr)      # TMA-staged variant
L.ppo_b200_sync()
print("sanitize_small done")
