/* examples/pendulum_b200.c — a plain-C caller of the drop-in library, written against the reference's API.
 *
 * Same calls as the reference's src/main.c (create_ppo / train_ppo_epoch / eval_ppo / save_ppo), but with the additive
 * vectorised device env (create_pendulum_env_cuda) so that train_ppo_epoch takes the fused rollout path.
 *
 *   gcc -std=gnu11 -O2 -I include examples/pendulum_b200.c -o build/pendulum_b200 \
 *       -L ppo.c_b200 -lppo_b200 -lm -Wl,-rpath,$PWD/ppo.c_b200
 *   ./build/pendulum_b200 [epochs=40] [n_envs=4096]
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "ppo.h"
#include "env.h"

int main(int argc, char** argv) {
    const int epochs = argc > 1 ? atoi(argv[1]) : 40;
    const int n_envs = argc > 2 ? atoi(argv[2]) : 4096;
    const int T = 200;                                   /* one full Pendulum episode per env and iteration */
    srand(1);                                            /* the library draws the initial weights from glibc rand(), like the reference */
    Env* env = create_pendulum_env_cuda(n_envs, 1);
    int layer_sizes[] = {env->state_size, 64, 64, env->action_size};
    char* activation_functions[] = {"tanh", "tanh", "none"};
    /* hyper-parameters of the reference's main.c:33-43 */
    PPO* ppo = create_ppo(activation_functions, layer_sizes, 4, n_envs * T, 3e-4f, 3e-4f, 0.95f, 0.2f, 0.0f, 1.0f, true);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    float best = -1e9f;
    for (int i = 0; i < epochs; i++) {
        train_ppo_epoch(ppo, env, n_envs * T, 16384, 4, 10);         /* rollout + GAE + 10 value / 4 policy epochs */
        const float r = ppo_b200_last_mean_return(ppo);              /* eval_ppo's "R" over this iteration's episodes */
        if (r > best) best = r;
        if (i % 5 == 0 || i == epochs - 1) printf("Epoch: %d mean return: %.1f\n", i, r);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    printf("best mean return %.1f, %.2f s, %.1f M env-steps/s (with host mirrors every epoch)\n", best, s,
           1e-6 * (double)epochs * n_envs * T / s);
    eval_ppo(ppo, env, n_envs * T);                                  /* prints J / R / Episodes like the reference */
    save_ppo(ppo, "/tmp/ppo_model_b200.bin");                        /* reference byte format (src/ppo.cu:585-607) */
    free_ppo(ppo);
    env->free_env();
    return best > -200.f ? 0 : 1;
}
