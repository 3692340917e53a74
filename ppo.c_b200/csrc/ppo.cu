// ppo.cu — the PPO object and the training schedule (§8f eval/checkpoint).
//
// Mirrors reference src/ppo.cu: create/free (:6-51), collect_trajectories (:54-79), compute_gae[_cuda]
// (:261-369), _train_ppo_epoch[_cuda] (:373-558), eval_ppo (:560-583), save/load (:585-648).
//
// Design differences from the reference's CUDA twin (all behind the same signatures):
//   * device-resident: nothing is copied to the host inside an iteration, no cudaMalloc, no
//     blocking scalar read per minibatch (the reference does 1-3 syncs + 1-2 cudaMallocs each,
//     src/loss.cu:51-57, src/ppo.cu:150-156); losses accumulate in device scalars;
//   * per minibatch: gather -> fused forward layers -> fused loss head -> backward with split-K slabs
//     -> Adam with the slab reduction fused in (or slab-reduce -> NCCL all-reduce -> Adam under DP);
//   * permutations: the reference's rand() swap chain stays on the host for bit-exact indices, but is
//     produced one epoch AHEAD into pinned memory and uploaded asynchronously (double-buffered), so
//     it overlaps the previous epoch's kernels; an optional device generator removes it altogether;
//   * rollout: fused device kernel for the vectorised Pendulum, per-step kernel for opaque host envs.
#include <unordered_map>

#include "common.cuh"
#include "internal.h"

namespace b200 {

struct Trainer {
    int mb_cap = 0;
    float *states = nullptr, *actions = nullptr, *lp_old = nullptr, *adv = nullptr, *advt = nullptr;
    float *lp = nullptr, *gl = nullptr, *gmu = nullptr;
    int* h_perm[2] = {nullptr, nullptr};
    int* d_perm[2] = {nullptr, nullptr};
    cudaEvent_t perm_evt[2] = {nullptr, nullptr};
    int perm_cap = 0, perm_slot = 0;
    int* h_perm_all = nullptr;   // all permutations of one phase (persistent phase kernel): [epochs][limit]
    int* d_perm_all = nullptr;
    size_t perm_all_cap = 0;
    cudaEvent_t perm_all_evt = nullptr;
    float* d_align = nullptr;    // one float all-reduced at the start of a data-parallel update (rank alignment)
    int perm_mode = -1;          // -1 auto: device generator after a device rollout, reference rand() chain otherwise
    bool last_rollout_on_device = false;
    unsigned long long perm_seed = 0, perm_epoch = 0;
    float* d_scalars = nullptr;   // [0] value-loss sum, [1] policy-loss sum, [2..3] return stats, [4..] dist triples
    int n_v_steps = 0, n_p_steps = 0;
    bool obs_norm = false;
    double* d_dist_triples = nullptr;
    float* packed = nullptr;     // row-packed mirror of the gathered fields (layer-wise update path), [packed_cap][PW]
    size_t packed_cap = 0;
    float eval_J = 0.f, eval_R = 0.f;   // what the last eval_ppo printed (src/ppo.cu:581)
    int eval_episodes = 0;
};
static std::unordered_map<PPO*, Trainer*> g_trainers;

__global__ void add_scalar_kernel(float* x, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += v;
}
static void add_scalar(float* x, int n, float v) { B200_LAUNCH(add_scalar_kernel, div_up(n, 64), 64, 0, x, n, v); }

static Trainer* trainer(PPO* ppo) {
    auto it = g_trainers.find(ppo);
    if (it == g_trainers.end()) B200_FATAL("PPO %p was not created by this library", (void*)ppo);
    return it->second;
}

static Trainer* attach_trainer(PPO* ppo) {
    ensure_device();
    Trainer* t = new Trainer();
    t->d_scalars = dmalloc<float>(64);
    CUDA_CHECK(cudaMemset(t->d_scalars, 0, 64 * sizeof(float)));
    t->d_dist_triples = dmalloc<double>(3 * 64);
    CUDA_CHECK(cudaEventCreateWithFlags(&t->perm_evt[0], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&t->perm_evt[1], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&t->perm_all_evt, cudaEventDisableTiming));
    g_trainers[ppo] = t;
    return t;
}

static void ensure_minibatch(Trainer* t, int mb, int S, int A) {
    if (mb <= t->mb_cap) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    float** ptrs[] = {&t->states, &t->actions, &t->lp_old, &t->adv, &t->advt, &t->lp, &t->gl, &t->gmu};
    for (float** p : ptrs) if (*p) CUDA_CHECK(cudaFree(*p));
    t->states = dmalloc<float>((size_t)mb * S);
    t->actions = dmalloc<float>((size_t)mb * A);
    t->lp_old = dmalloc<float>(mb);
    t->adv = dmalloc<float>(mb);
    t->advt = dmalloc<float>(mb);
    t->lp = dmalloc<float>(mb);
    t->gl = dmalloc<float>(mb);
    t->gmu = dmalloc<float>((size_t)mb * A);
    t->mb_cap = mb;
}

static void ensure_perm(Trainer* t, int n) {
    if (n <= t->perm_cap) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    for (int s = 0; s < 2; s++) {
        if (t->h_perm[s]) CUDA_CHECK(cudaFreeHost(t->h_perm[s]));
        if (t->d_perm[s]) CUDA_CHECK(cudaFree(t->d_perm[s]));
        t->h_perm[s] = hmalloc_pinned<int>(n);
        t->d_perm[s] = dmalloc<int>(n);
    }
    t->perm_cap = n;
}

// Next permutation of [0,n) on the device; returns the device index array.
static const int* next_permutation(Trainer* t, int n) {
    ensure_perm(t, n);
    const int s = t->perm_slot;
    t->perm_slot ^= 1;
    const bool on_device = t->perm_mode == 1 || (t->perm_mode < 0 && t->last_rollout_on_device);
    if (on_device) {
        ppo_b200_permutation(t->d_perm[s], n, t->perm_seed, t->perm_epoch++);
    } else {
        CUDA_CHECK(cudaEventSynchronize(t->perm_evt[s]));   // the upload that last used this pinned slot
        host_shuffle(t->h_perm[s], n);                      // glibc rand(), reference order
        CUDA_CHECK(cudaMemcpyAsync(t->d_perm[s], t->h_perm[s], (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream()));
        CUDA_CHECK(cudaEventRecord(t->perm_evt[s], stream()));
    }
    return t->d_perm[s];
}

// The permutations of `n_epochs` consecutive epochs, [n_epochs][n] on the device, generated up front: nothing but the
// shuffles draws from rand() during the update phase, so producing them before the first minibatch of
// the phase leaves the reference's rand() stream untouched.
static const int* phase_permutations(Trainer* t, int n, int n_epochs) {
    const size_t need = (size_t)n * n_epochs;
    if (need > t->perm_all_cap) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (t->h_perm_all) CUDA_CHECK(cudaFreeHost(t->h_perm_all));
        if (t->d_perm_all) CUDA_CHECK(cudaFree(t->d_perm_all));
        t->h_perm_all = nullptr;
        t->d_perm_all = dmalloc<int>(need);
        t->perm_all_cap = need;
    }
    const bool on_device = t->perm_mode == 1 || (t->perm_mode < 0 && t->last_rollout_on_device);
    if (on_device) {
        for (int e = 0; e < n_epochs; e++) ppo_b200_permutation(t->d_perm_all + (size_t)e * n, n, t->perm_seed, t->perm_epoch++);
    } else {
        if (!t->h_perm_all) t->h_perm_all = hmalloc_pinned<int>(t->perm_all_cap);
        CUDA_CHECK(cudaEventSynchronize(t->perm_all_evt));     // the upload that last read this pinned buffer
        for (int e = 0; e < n_epochs; e++) host_shuffle(t->h_perm_all + (size_t)e * n, n);   // glibc rand(), reference order
        CUDA_CHECK(cudaMemcpyAsync(t->d_perm_all, t->h_perm_all, need * sizeof(int), cudaMemcpyHostToDevice, stream()));
        CUDA_CHECK(cudaEventRecord(t->perm_all_evt, stream()));
    }
    return t->d_perm_all;
}

// One phase (all value epochs or all policy epochs) through the persistent kernel, in launches of as many epochs as fit
// the permutation budget.
static void phase_epochs(Trainer* t, NeuralNetwork* nn, GaussianPolicy* pol, Adam* adam_net, Adam* adam_ls, float lr, int limit,
                         int mb_local, int mb_total, int row0, int batch_size, int num_batches, int n_epochs,
                         const TrajectoryBuffer* b, float epsilon, float ent_coeff, float* loss_slot, bool peer) {
    const int per_launch = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_epochs, (size_t(256) << 20) / ((size_t)limit * sizeof(int))));
    for (int e0 = 0; e0 < n_epochs; e0 += per_launch) {
        const int ne = std::min(per_launch, n_epochs - e0);
        const int* perms = phase_permutations(t, limit, ne);
        fused_phase_update(nn, pol, adam_net, adam_ls, lr, perms, limit, mb_local, mb_total, row0, batch_size, num_batches, ne, b,
                           epsilon, ent_coeff, loss_slot, peer);
    }
}

// PPO_B200_FUSED=0 forces the generic layer-wise kernels (A/B testing, parity tests of both paths).
static bool use_fused_env() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("PPO_B200_FUSED"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}
static bool packed_gather_enabled() {      // PPO_B200_PACKED_GATHER=0 gathers from the SoA arrays (A/B runs)
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("PPO_B200_PACKED_GATHER"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}
static int g_force_path = -1;   // tests: -1 env default, 0 layer-wise, 1 fused
static bool use_fused() { return g_force_path < 0 ? use_fused_env() : g_force_path == 1; }

// compute_gae on device arrays (src/ppo.cu:261-323): two V forwards, scan, global stats, normalise.
static void gae_device(NeuralNetwork* V, TrajectoryBuffer* b, Trainer* t, int limit, float gamma, float lambda) {
    float* v_next = static_cast<float*>(scratch(kScratchGaeVNext, (size_t)limit * sizeof(float)));
    const float* v;
    if (use_fused() && fused_supported(V)) {   // one launch per pass, activations stay in shared memory
        float* vbuf = static_cast<float*>(scratch(kScratchGaeV, (size_t)limit * sizeof(float)));
        fused_forward(V, b->d_next_state_p, limit, v_next);
        fused_forward(V, b->d_state_p, limit, vbuf);
        v = vbuf;
    } else {
        net_forward(V, b->d_next_state_p, limit, true);
        CUDA_CHECK(cudaMemcpyAsync(v_next, V->d_output, (size_t)limit * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
        net_forward(V, b->d_state_p, limit, true);
        v = V->d_output;
    }
    GaeWork w = gae_scan(b->d_reward_p, v, v_next, b->d_terminated_p, b->d_truncated_p, limit, gamma, lambda,
                         b->d_advantage_p, b->d_adv_target_p);
    if (dist_active() && dist_shard_mode() == 1 && t) {
        // every rank scanned its own envs: merge the Welford triples of all ranks in rank order
        dist_allgather_doubles(w.stats_d, t->d_dist_triples, 3);
        gae_merge_ranks(t->d_dist_triples, dist_world(), w.stats_f);
    }
    gae_normalize(b->d_advantage_p, limit, w.stats_f);
}

static void update_device(PPO* ppo, float gamma, int batch_size, int n_epochs_policy, int n_epochs_value) {
    Trainer* t = trainer(ppo);
    TrajectoryBuffer* b = ppo->buffer;
    GaussianPolicy* pol = ppo->policy;
    const int S = b->state_size, A = b->action_size;
    const int limit = b->full ? b->capacity : b->idx;
    if (limit <= 0 || batch_size <= 0) return;

    gae_device(ppo->V, b, t, limit, gamma, ppo->lambda);

    // data-parallel split of each minibatch
    const int G = dist_active() ? dist_world() : 1, rank = dist_active() ? dist_rank() : 0;
    int mb_local = batch_size, mb_total = batch_size, row0 = 0;
    if (G > 1) {
        if (dist_shard_mode() == 0) {
            if (batch_size % G) B200_FATAL("batch_size %d not divisible by world size %d", batch_size, G);
            mb_local = batch_size / G;
            row0 = rank * mb_local;
        } else {
            mb_total = batch_size * G;
        }
    }
    ensure_minibatch(t, mb_local, S, A);
    const int num_batches = limit / batch_size;   // src/ppo.cu:387-388 (ceilf of an integer quotient)
    NetDev* ndV = net_dev(ppo->V);
    NetDev* ndP = net_dev(pol->mu);
    const bool fusedV = use_fused() && fused_supported(ppo->V);
    const bool fusedP = use_fused() && fused_supported(pol->mu);
    // small-net gradients cross the GPUs inside the update kernel over NVLink peer memory when it can be mapped and the
    // net's slab fits a receive lane; otherwise that net goes slab-reduce -> NCCL all-reduce -> Adam
    const bool peer_ok = G > 1 && (fusedV || fusedP) && dist_peer_ready();
    const bool peerV = peer_ok && fusedV && ndV->param_count + 2 <= kPeerCap;
    const bool peerP = peer_ok && fusedP && ndP->param_count + A + 1 <= kPeerCap;
    if (peerV || peerP) {
        // ranks enter the exchange loop together: benign skew (a slow host shuffle, first-iteration allocations) is absorbed
        // by a stream-ordered NCCL collective instead of the in-kernel poll
        if (!t->d_align) { t->d_align = dmalloc<float>(1); CUDA_CHECK(cudaMemsetAsync(t->d_align, 0, sizeof(float), stream())); }
        dist_allreduce_sum(t->d_align, 1);
    }
    // layer-wise nets gather every row once per epoch from arrays that are final after GAE: build the row-packed mirror once
    const bool packed = (!fusedV || !fusedP) && num_batches > 0 && packed_gather_enabled() && packed_row_floats(S, A) <= 128;
    if (packed) {
        const size_t need = (size_t)limit * packed_row_floats(S, A);
        if (need > t->packed_cap) {
            CUDA_CHECK(cudaStreamSynchronize(stream()));
            if (t->packed) CUDA_CHECK(cudaFree(t->packed));
            t->packed = dmalloc<float>(need);
            t->packed_cap = need;
        }
        launch_pack_rows(t->packed, limit, S, A, b->d_state_p, b->d_action_p, b->d_logprob_p, b->d_advantage_p, b->d_adv_target_p);
    }
    CUDA_CHECK(cudaMemsetAsync(t->d_scalars, 0, 2 * sizeof(float), stream()));
    t->n_v_steps = n_epochs_value * num_batches;
    t->n_p_steps = n_epochs_policy * num_batches;
    const bool phaseV = fusedV && (G == 1 || peerV) && fused_phase_supported(ppo->V);
    const bool phaseP = fusedP && (G == 1 || peerP) && fused_phase_supported(pol->mu);

    // ---- value epochs, src/ppo.cu:491-510
    if (phaseV && num_batches > 0)
        phase_epochs(t, ppo->V, nullptr, ppo->adam_V, nullptr, ppo->lr_V, limit, mb_local, mb_total, row0, batch_size, num_batches,
                     n_epochs_value, b, 0.f, 0.f, t->d_scalars + 0, peerV);
    for (int j = 0; j < (phaseV ? 0 : n_epochs_value); j++) {
        const int* perm = next_permutation(t, limit);
        for (int k = 0; k < num_batches; k++) {
            if (fusedV) {
                const bool peer = peerV;
                float* red = (G > 1 && !peer) ? static_cast<float*>(scratch(kScratchMisc, (ndV->param_count + 2) * sizeof(float))) : nullptr;
                fused_minibatch_update(ppo->V, nullptr, ppo->adam_V, nullptr, ppo->lr_V, perm, k * batch_size + row0, limit,
                                       mb_local, mb_total, b, 0.f, 0.f, t->d_scalars + 0, red, k > 0, peer);
                if (G > 1 && !peer) {   // slab-reduce -> NCCL all-reduce -> Adam
                    dist_allreduce_sum(red, ndV->param_count + 2);
                    ppo->adam_V->time_step += 1;
                    adam_flat(ndV->params, red, ppo->adam_V->m, ppo->adam_V->v, (int)ndV->param_count, ppo->lr_V,
                              ppo->adam_V->beta1, ppo->adam_V->beta2, ppo->adam_V->time_step, nullptr, 0, 0);
                }
                continue;
            }
            ndV->image_dirty = true;
            if (packed) launch_gather_packed(perm, k * batch_size + row0, limit, mb_local, S, A, t->packed, t->states, t->actions, t->lp_old, t->adv, t->advt);
            else launch_gather(perm, k * batch_size + row0, limit, mb_local, S, A, b->d_state_p, b->d_action_p, b->d_logprob_p,
                               b->d_advantage_p, b->d_adv_target_p, t->states, t->actions, t->lp_old, t->adv, t->advt);
            net_forward(ppo->V, t->states, mb_local, true);
            launch_value_head(ppo->V->d_output, t->advt, t->gl, mb_local, mb_total, t->d_scalars + 0);
            net_backward_partials(ppo->V, t->gl, mb_local);
            ppo->adam_V->time_step += 1;
            if (G > 1) {
                net_reduce_grads(ppo->V);
                dist_allreduce_sum(ndV->grads, ndV->param_count);
                adam_flat(ndV->params, ndV->grads, ppo->adam_V->m, ppo->adam_V->v, (int)ndV->param_count, ppo->lr_V,
                          ppo->adam_V->beta1, ppo->adam_V->beta2, ppo->adam_V->time_step, nullptr, 0, 0);
            } else {
                adam_flat(ndV->params, ndV->grads, ppo->adam_V->m, ppo->adam_V->v, (int)ndV->param_count, ppo->lr_V,
                          ppo->adam_V->beta1, ppo->adam_V->beta2, ppo->adam_V->time_step, ndV->partials,
                          ndV->last_splits, ndV->slab_stride());
            }
        }
    }
    // ---- policy epochs, src/ppo.cu:512-533
    if (phaseP && num_batches > 0)
        phase_epochs(t, pol->mu, pol, ppo->adam_policy, ppo->adam_entropy, ppo->lr_policy, limit, mb_local, mb_total, row0, batch_size,
                     num_batches, n_epochs_policy, b, ppo->epsilon, ppo->ent_coeff, t->d_scalars + 1, peerP);
    for (int j = 0; j < (phaseP ? 0 : n_epochs_policy); j++) {
        const int* perm = next_permutation(t, limit);
        for (int k = 0; k < num_batches; k++) {
            if (fusedP) {
                const bool peer = peerP;
                float* red = (G > 1 && !peer) ? static_cast<float*>(scratch(kScratchMisc, (ndP->param_count + A + 1) * sizeof(float))) : nullptr;
                fused_minibatch_update(pol->mu, pol, ppo->adam_policy, ppo->adam_entropy, ppo->lr_policy, perm,
                                       k * batch_size + row0, limit, mb_local, mb_total, b, ppo->epsilon, ppo->ent_coeff,
                                       t->d_scalars + 1, red, k > 0, peer);
                if (G > 1 && !peer) {
                    dist_allreduce_sum(red, ndP->param_count + A + 1);
                    // entropy gradient (src/ppo.cu:436-438): added once, AFTER the cross-rank sum
                    if (ppo->ent_coeff != 0.f) add_scalar(red + ndP->param_count, A, -ppo->ent_coeff);
                    ppo->adam_entropy->time_step += 1;
                    ppo->adam_policy->time_step += 1;
                    adam_flat(pol->d_log_std, red + ndP->param_count, ppo->adam_entropy->m, ppo->adam_entropy->v, A, ppo->lr_policy,
                              ppo->adam_entropy->beta1, ppo->adam_entropy->beta2, ppo->adam_entropy->time_step, nullptr, 0, 0);
                    adam_flat(ndP->params, red, ppo->adam_policy->m, ppo->adam_policy->v, (int)ndP->param_count, ppo->lr_policy,
                              ppo->adam_policy->beta1, ppo->adam_policy->beta2, ppo->adam_policy->time_step, nullptr, 0, 0);
                }
                continue;
            }
            ndP->image_dirty = true;
            if (packed) launch_gather_packed(perm, k * batch_size + row0, limit, mb_local, S, A, t->packed, t->states, t->actions, t->lp_old, t->adv, t->advt);
            else launch_gather(perm, k * batch_size + row0, limit, mb_local, S, A, b->d_state_p, b->d_action_p, b->d_logprob_p,
                               b->d_advantage_p, b->d_adv_target_p, t->states, t->actions, t->lp_old, t->adv, t->advt);
            net_forward(pol->mu, t->states, mb_local, true);
            launch_policy_head(pol->mu->d_output, pol->d_log_std, t->actions, t->lp_old, t->adv, mb_local, A, mb_total,
                               ppo->epsilon, ppo->ent_coeff, t->lp, t->gmu, pol->d_log_std_grad, t->d_scalars + 1);
            net_backward_partials(pol->mu, t->gmu, mb_local);
            ppo->adam_entropy->time_step += 1;
            ppo->adam_policy->time_step += 1;
            if (G > 1) {
                net_reduce_grads(pol->mu);
                dist_allreduce_sum(ndP->grads, ndP->param_count);
                dist_allreduce_sum(pol->d_log_std_grad, A);
                // launch_policy_head added -ent_coeff on every rank: the sum carries it G times, the reference once
                if (ppo->ent_coeff != 0.f) add_scalar(pol->d_log_std_grad, A, (G - 1) * ppo->ent_coeff);
            }
            // log_std first, then the mu-net (src/ppo.cu:529-531)
            adam_flat(pol->d_log_std, pol->d_log_std_grad, ppo->adam_entropy->m, ppo->adam_entropy->v, A, ppo->lr_policy,
                      ppo->adam_entropy->beta1, ppo->adam_entropy->beta2, ppo->adam_entropy->time_step, nullptr, 0, 0);
            if (G > 1)
                adam_flat(ndP->params, ndP->grads, ppo->adam_policy->m, ppo->adam_policy->v, (int)ndP->param_count,
                          ppo->lr_policy, ppo->adam_policy->beta1, ppo->adam_policy->beta2, ppo->adam_policy->time_step,
                          nullptr, 0, 0);
            else
                adam_flat(ndP->params, ndP->grads, ppo->adam_policy->m, ppo->adam_policy->v, (int)ndP->param_count,
                          ppo->lr_policy, ppo->adam_policy->beta1, ppo->adam_policy->beta2, ppo->adam_policy->time_step,
                          ndP->partials, ndP->last_splits, ndP->slab_stride());
        }
    }
}

static void activate_device(TrajectoryBuffer* b) {
    b->action_p = b->d_action_p; b->state_p = b->d_state_p; b->next_state_p = b->d_next_state_p;
    b->reward_p = b->d_reward_p; b->logprob_p = b->d_logprob_p; b->advantage_p = b->d_advantage_p;
    b->adv_target_p = b->d_adv_target_p; b->terminated_p = b->d_terminated_p; b->truncated_p = b->d_truncated_p;
}

static void collect_device(TrajectoryBuffer* buffer, DeviceEnv* e, GaussianPolicy* policy, int steps, float* ret_stats) {
    const int n_envs = device_env_count(e);
    if (steps % n_envs) B200_FATAL("steps %d is not a multiple of n_envs %d", steps, n_envs);
    if (buffer->idx != 0) B200_FATAL("device rollout needs an empty or full buffer (idx == 0)");
    device_rollout(e, policy, buffer, steps / n_envs, nullptr, nullptr, ret_stats);
    activate_device(buffer);
    buffer->idx = (buffer->idx + steps) % buffer->capacity;
    buffer->full = buffer->full || buffer->idx == 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

PPO* create_ppo(char** activation_functions, int* layer_sizes, int num_layers, int buffer_size, float lr_policy,
                float lr_v, float lambda, float epsilon, float ent_coeff, float init_std, bool use_cuda) {
    PPO* ppo = (PPO*)malloc(sizeof(PPO));
    ppo->buffer = create_trajectory_buffer(buffer_size, layer_sizes[0], layer_sizes[num_layers - 1]);
    ppo->policy = create_gaussian_policy(layer_sizes, activation_functions, num_layers, init_std);   // mu-net first (rand() order)
    std::vector<int> sizes_v(layer_sizes, layer_sizes + num_layers);
    sizes_v[num_layers - 1] = 1;
    ppo->V = create_neural_network(sizes_v.data(), activation_functions, num_layers);
    // The optimiser state always lives on the device: use_cuda=false callers get the same kernels
    // (this library has no CPU arithmetic path); the flag is kept for the checkpoint/ABI.
    ppo->adam_policy = create_adam_from_nn_cuda(ppo->policy->mu, 0.9, 0.999);
    ppo->adam_V = create_adam_from_nn_cuda(ppo->V, 0.9, 0.999);
    ppo->adam_entropy = create_adam_cuda(&ppo->policy->d_log_std, &ppo->policy->d_log_std_grad, &ppo->policy->action_size, 1,
                                         ppo->policy->action_size, 0.9, 0.999);
    ppo->lambda = lambda;
    ppo->epsilon = epsilon;
    ppo->ent_coeff = ent_coeff;
    ppo->lr_policy = lr_policy;
    ppo->lr_V = lr_v;
    ppo->use_cuda = use_cuda;
    attach_trainer(ppo);
    return ppo;
}

void free_ppo(PPO* ppo) {
    if (!ppo) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    auto it = g_trainers.find(ppo);
    if (it != g_trainers.end()) {
        Trainer* t = it->second;
        float* ptrs[] = {t->states, t->actions, t->lp_old, t->adv, t->advt, t->lp, t->gl, t->gmu, t->d_scalars};
        for (float* p : ptrs) if (p) CUDA_CHECK(cudaFree(p));
        CUDA_CHECK(cudaFree(t->d_dist_triples));
        if (t->h_perm_all) CUDA_CHECK(cudaFreeHost(t->h_perm_all));
        if (t->d_perm_all) CUDA_CHECK(cudaFree(t->d_perm_all));
        if (t->d_align) CUDA_CHECK(cudaFree(t->d_align));
        if (t->packed) CUDA_CHECK(cudaFree(t->packed));
        CUDA_CHECK(cudaEventDestroy(t->perm_all_evt));
        for (int s = 0; s < 2; s++) {
            if (t->h_perm[s]) CUDA_CHECK(cudaFreeHost(t->h_perm[s]));
            if (t->d_perm[s]) CUDA_CHECK(cudaFree(t->d_perm[s]));
            CUDA_CHECK(cudaEventDestroy(t->perm_evt[s]));
        }
        delete t;
        g_trainers.erase(it);
    }
    free_adam_cuda(ppo->adam_policy);
    free_adam_cuda(ppo->adam_V);
    free_adam_cuda(ppo->adam_entropy);
    free_trajectory_buffer(ppo->buffer, true);
    free_gaussian_policy(ppo->policy);
    free_neural_network(ppo->V);
    free(ppo);
}

void collect_trajectories(TrajectoryBuffer* buffer, Env* env, GaussianPolicy* policy, int steps) {
    if (DeviceEnv* e = as_device_env(env)) {
        collect_device(buffer, e, policy, steps, nullptr);
        return;
    }
    // opaque host env: reference bookkeeping (src/ppo.cu:54-79) around a per-step device sample.
    // The active pointer set must be the host one (it is after create / buffer_to_host / reset).
    if (buffer->state_p != buffer->h_state_p) buffer_to_host(buffer);
    const int S = buffer->state_size, A = buffer->action_size;
    const bool mailbox = host_sampler_supported(policy);      // resident sampler kernel instead of a launch + sync per step
    env->reset_env(buffer->state(buffer, buffer->idx));
    for (int i = 0; i < steps; i++) {
        const int idx = buffer->idx;
        int draws[16];
        int nd = 0;
        if (A == 1) { draws[0] = rand(); draws[1] = rand(); nd = 2; }
        else {
            if (A > 14) B200_FATAL("host-env rollout supports action_size <= 14");
            for (int q = 0; q + 1 < A; q += 2) { draws[nd++] = rand(); draws[nd++] = rand(); }
            if (A & 1) { draws[nd++] = rand(); draws[nd++] = rand(); }
        }
        if (mailbox) {
            host_sampler_step(policy, buffer->h_state_p + (size_t)idx * S, buffer->h_action_p + (size_t)idx * A,
                              buffer->h_logprob_p + idx, draws, nd);
        } else {
            launch_sample_action(policy, buffer->h_state_p + (size_t)idx * S, buffer->h_action_p + (size_t)idx * A,
                                 buffer->h_logprob_p + idx, draws, nd);
            CUDA_CHECK(cudaStreamSynchronize(stream()));
        }
        env->step_env(buffer->action(buffer, idx), buffer->next_state(buffer, idx), buffer->reward(buffer, idx),
                      buffer->terminated(buffer, idx), buffer->truncated(buffer, idx), buffer->action_size);
        const int new_idx = (idx + 1) % buffer->capacity;
        if (i < steps - 1) {
            if (*buffer->truncated(buffer, idx) || *buffer->terminated(buffer, idx)) env->reset_env(buffer->state(buffer, new_idx));
            else memcpy(buffer->state(buffer, new_idx), buffer->next_state(buffer, idx), S * sizeof(float));
        } else if (!*buffer->terminated(buffer, idx)) {
            *buffer->truncated(buffer, idx) = true;
        }
        buffer->idx = new_idx;
        buffer->full = buffer->full || buffer->idx == 0;
    }
    if (mailbox) host_sampler_stop();
}

void compute_gae_cuda(NeuralNetwork* V, TrajectoryBuffer* buffer, float gamma, float lambda, int horizon) {
    (void)horizon;
    const int limit = buffer->full ? buffer->capacity : buffer->idx;
    if (buffer->state_p != buffer->d_state_p) B200_FATAL("compute_gae_cuda: buffer is not on the device (call buffer_to_device)");
    gae_device(V, buffer, nullptr, limit, gamma, lambda);
}

void compute_gae(NeuralNetwork* V, TrajectoryBuffer* buffer, float gamma, float lambda) {
    // host-pointer twin: stage through the buffer's own device arrays
    const bool was_host = buffer->state_p == buffer->h_state_p;
    if (was_host) { nn_write_weights_to_device(V); buffer_to_device(buffer); }
    compute_gae_cuda(V, buffer, gamma, lambda, 0);
    if (was_host) buffer_to_host(buffer);
}

void ppo_b200_buffer_upload(PPO* ppo) {
    buffer_to_device(ppo->buffer);
    ppo->buffer->idx = 0;
    ppo->buffer->full = true;
}

void ppo_b200_sync_host(PPO* ppo) {     // src/ppo.cu:536-538
    buffer_to_host(ppo->buffer);
    policy_to_host(ppo->policy);
    nn_write_weights_to_host(ppo->V);
}

void ppo_b200_update_device(PPO* ppo, float gamma, int batch_size, int n_epochs_policy, int n_epochs_value) {
    update_device(ppo, gamma, batch_size, n_epochs_policy, n_epochs_value);
}

void ppo_b200_update(PPO* ppo, float gamma, int batch_size, int n_epochs_policy, int n_epochs_value) {
    trainer(ppo)->last_rollout_on_device = false;
    buffer_upload_inputs(ppo->buffer);            // the GAE outputs need no upload, the kernels run behind the copies
    ppo->buffer->idx = 0;
    ppo->buffer->full = true;
    update_device(ppo, gamma, batch_size, n_epochs_policy, n_epochs_value);
    buffer_download_outputs(ppo->buffer);         // only advantage / adv_target changed on the device
    policy_to_host(ppo->policy);
    nn_write_weights_to_host(ppo->V);
}

void ppo_b200_train_iterations(PPO* ppo, Env* env, int n_iters, int batch_size, int n_epochs_policy, int n_epochs_value) {
    Trainer* t = trainer(ppo);
    DeviceEnv* e = as_device_env(env);
    for (int i = 0; i < n_iters; i++) {
        t->last_rollout_on_device = e != nullptr;
        if (e) {
            collect_device(ppo->buffer, e, ppo->policy, ppo->buffer->capacity, t->d_scalars + 2);
        } else {
            collect_trajectories(ppo->buffer, env, ppo->policy, ppo->buffer->capacity);
            buffer_to_device(ppo->buffer);
        }
        update_device(ppo, env->gamma, batch_size, n_epochs_policy, n_epochs_value);
    }
}

void train_ppo_epoch(PPO* ppo, Env* env, int steps_per_epoch, int batch_size, int n_epochs_policy, int n_epochs_value) {
    const int iters = steps_per_epoch / ppo->buffer->capacity;     // src/ppo.cu:479
    for (int i = 0; i < iters; i++) {
        ppo_b200_train_iterations(ppo, env, 1, batch_size, n_epochs_policy, n_epochs_value);
        ppo_b200_sync_host(ppo);                                   // src/ppo.cu:536-538, every iteration
    }
}

void eval_ppo(PPO* ppo, Env* env, int steps) {                     // src/ppo.cu:560-583
    reset_buffer(ppo->buffer);
    if (ppo->buffer->state_p != ppo->buffer->h_state_p) buffer_to_host(ppo->buffer);
    collect_trajectories(ppo->buffer, env, ppo->policy, steps);
    if (ppo->buffer->state_p != ppo->buffer->h_state_p) buffer_to_host(ppo->buffer);
    float rewards = *ppo->buffer->reward(ppo->buffer, steps - 1);
    float episode_J = *ppo->buffer->reward(ppo->buffer, steps - 1);
    int n_episodes = 1;
    float sum_J = 0;
    for (int i = steps - 2; i >= 0; i--) {
        rewards += *ppo->buffer->reward(ppo->buffer, i);
        episode_J = *ppo->buffer->reward(ppo->buffer, i) + env->gamma * episode_J;
        if (*ppo->buffer->terminated(ppo->buffer, i) || *ppo->buffer->truncated(ppo->buffer, i)) {
            n_episodes++;
            sum_J += episode_J;
            episode_J = 0;
        }
    }
    printf("J: %f R: %f Episodes: %d\n", sum_J / n_episodes, rewards / n_episodes, n_episodes);
    Trainer* t = trainer(ppo);
    t->eval_J = sum_J / n_episodes;
    t->eval_R = rewards / n_episodes;
    t->eval_episodes = n_episodes;
    reset_buffer(ppo->buffer);
}

void ppo_b200_last_eval(PPO* ppo, float* J, float* R, int* episodes) {
    Trainer* t = trainer(ppo);
    if (J) *J = t->eval_J;
    if (R) *R = t->eval_R;
    if (episodes) *episodes = t->eval_episodes;
}

void save_ppo(PPO* ppo, const char* filename) {                    // src/ppo.cu:585-607, same byte format
    FILE* file = fopen(filename, "wb");
    if (!file) B200_FATAL("save_ppo: cannot open %s", filename);
    policy_to_host(ppo->policy);
    nn_write_weights_to_host(ppo->V);
    fwrite(&ppo->lambda, sizeof(float), 1, file);
    fwrite(&ppo->epsilon, sizeof(float), 1, file);
    fwrite(&ppo->ent_coeff, sizeof(float), 1, file);
    fwrite(&ppo->lr_policy, sizeof(float), 1, file);
    fwrite(&ppo->lr_V, sizeof(float), 1, file);
    fwrite(&ppo->buffer->state_size, sizeof(int), 1, file);
    fwrite(&ppo->buffer->action_size, sizeof(int), 1, file);
    fwrite(&ppo->buffer->capacity, sizeof(int), 1, file);
    save_policy(ppo->policy, file);
    save_neural_network(ppo->V, file);
    save_adam(ppo->adam_policy, file, true);
    save_adam(ppo->adam_V, file, true);
    save_adam(ppo->adam_entropy, file, true);
    fclose(file);
}

PPO* load_ppo(const char* filename, bool use_cuda) {               // src/ppo.cu:610-648
    FILE* file = fopen(filename, "rb");
    if (!file) B200_FATAL("load_ppo: cannot open %s", filename);
    PPO* ppo = (PPO*)malloc(sizeof(PPO));
    ppo->use_cuda = use_cuda;
    int state_size, action_size, capacity;
    bool ok = fread(&ppo->lambda, sizeof(float), 1, file) == 1 && fread(&ppo->epsilon, sizeof(float), 1, file) == 1 &&
              fread(&ppo->ent_coeff, sizeof(float), 1, file) == 1 && fread(&ppo->lr_policy, sizeof(float), 1, file) == 1 &&
              fread(&ppo->lr_V, sizeof(float), 1, file) == 1 && fread(&state_size, sizeof(int), 1, file) == 1 &&
              fread(&action_size, sizeof(int), 1, file) == 1 && fread(&capacity, sizeof(int), 1, file) == 1;
    if (!ok) B200_FATAL("load_ppo: %s truncated", filename);
    ppo->buffer = create_trajectory_buffer(capacity, state_size, action_size);
    ppo->policy = load_policy(file, state_size, action_size);
    ppo->V = load_neural_network(file);
    ppo->adam_policy = load_adam_from_nn(file, ppo->policy->mu, true);
    ppo->adam_V = load_adam_from_nn(file, ppo->V, true);
    ppo->adam_entropy = load_adam(file, &ppo->policy->d_log_std, &ppo->policy->d_log_std_grad, &action_size, true);
    fclose(file);
    attach_trainer(ppo);
    return ppo;
}

void ppo_b200_set_permutation_mode(PPO* ppo, int mode, unsigned long long seed) {
    Trainer* t = trainer(ppo);
    t->perm_mode = mode;
    t->perm_seed = seed;
    t->perm_epoch = 0;
}

void ppo_b200_set_kernel_path(int path) { g_force_path = path; }

void ppo_b200_set_obs_norm(PPO* ppo, int enabled) {
    trainer(ppo)->obs_norm = enabled != 0;
    device_env_set_obs_norm(enabled != 0);
}

static float read_scalar(Trainer* t, int i) {
    float h[4];
    CUDA_CHECK(cudaMemcpyAsync(h, t->d_scalars, sizeof(h), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    return h[i];
}
float ppo_b200_last_mean_return(PPO* ppo) {
    Trainer* t = trainer(ppo);
    const float s = read_scalar(t, 2), n = read_scalar(t, 3);
    return n > 0 ? s / n : 0.f;
}
float ppo_b200_last_value_loss(PPO* ppo) { Trainer* t = trainer(ppo); return t->n_v_steps ? read_scalar(t, 0) / t->n_v_steps : 0.f; }
float ppo_b200_last_policy_loss(PPO* ppo) { Trainer* t = trainer(ppo); return t->n_p_steps ? read_scalar(t, 1) / t->n_p_steps : 0.f; }

}  // extern "C"
