/* oracle/ppo_oracle.c — see ppo_oracle.h.  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Arithmetic notes.  The reference's .cu files are compiled as C++ by nvcc, so `exp(float)`,
 * `sqrt(float)` pick the float overloads (expf, sqrtf), `pow(float, int)` promotes to double, and
 * double literals (0.5, 1e-8, M_PI) promote the surrounding expression to double.  This file is C,
 * so every such choice is spelled out; it is compiled with -ffp-contract=off like oracle/_ref.
 * All citations are relative to /root/reference. */
#include "ppo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846 /* policy.h:7 */

/* ============================== MLP ============================================================ */

int orc_param_count(const int* sizes, int num_layers) {
    int p = 0;
    for (int i = 0; i < num_layers - 1; i++) p += sizes[i] * sizes[i + 1] + sizes[i + 1];
    return p;
}

/* neural_network.cu:40-51 — uniform(-sqrt3*std, sqrt3*std), std = gain*sqrt(2/(in+out)),
 * gain sqrt2 for hidden layers and 1 for the last; bias uniform(+-1/sqrt(in)).  rand() order:
 * per layer all W then all b. */
void orc_init_params(float* params, const int* sizes, int num_layers) {
    float* p = params;
    for (int i = 0; i < num_layers - 1; i++) {
        float gain = (i == num_layers - 2) ? 1.0f : sqrtf(2.0);
        float std = gain * sqrtf(2.0 / (sizes[i] + sizes[i + 1]));
        for (int j = 0; j < sizes[i] * sizes[i + 1]; j++)
            *p++ = (2 * (float)rand() / RAND_MAX - 1) * sqrtf(3.0) * std;
        for (int j = 0; j < sizes[i + 1]; j++)
            *p++ = (2 * (float)rand() / RAND_MAX - 1) * (1. / sqrtf(sizes[i]));
    }
}

int orc_cache_floats(const int* sizes, int num_layers, int m) {
    int s = 0;
    for (int i = 0; i < num_layers; i++) s += sizes[i];
    return s * m;
}

const float* orc_mlp_output(const float* cache, const int* sizes, int num_layers, int m) {
    int s = 0;
    for (int i = 0; i < num_layers - 1; i++) s += sizes[i];
    return cache + (size_t)s * m;
}

/* mat_mul.cu:39-55 with the sequential-k sgemm of oracle/shim: out = (sum_p x*w from 0) + b. */
static void orc_mat_mul(float* out, const float* x, const float* w, const float* b, int m, int n, int l) {
#ifdef ORC_FAST /* timing-only build (oracle/Makefile `fast`): dot products the compiler may vectorise over p (-ffast-math) */
    for (int j = 0; j < m; j++)
        for (int k = 0; k < l; k++) {
            const float* xr = x + (size_t)j * n;
            const float* wr = w + (size_t)k * n;
            float acc = 0.0f;
#pragma GCC ivdep
            for (int p = 0; p < n; p++) acc += xr[p] * wr[p];
            out[(size_t)j * l + k] = acc + b[k];
        }
    return;
#endif
    for (int j = 0; j < m; j++)
        for (int k = 0; k < l; k++) {
            float acc = 0.0f;
            for (int p = 0; p < n; p++) acc += x[(size_t)j * n + p] * w[(size_t)k * n + p];
            out[(size_t)j * l + k] = acc + b[k];
        }
}

static void orc_act(float* x, int count, int act) {
    if (act == ORC_ACT_RELU) { /* activation_function.cu:5-9 */
        for (int i = 0; i < count; i++) x[i] = x[i] > 0 ? x[i] : 0;
    } else if (act == ORC_ACT_TANH) { /* [EXT] */
        for (int i = 0; i < count; i++) x[i] = tanhf(x[i]);
    }
}

/* grad <- grad * act'(post-activation y): activation_function.cu:11-15 for ReLU. */
static void orc_act_derivative(const float* y, float* grad, int count, int act) {
    if (act == ORC_ACT_RELU) {
        for (int i = 0; i < count; i++) grad[i] = y[i] > 0 ? grad[i] : 0;
    } else if (act == ORC_ACT_TANH) { /* [EXT] d tanh = 1 - y^2 */
        for (int i = 0; i < count; i++) grad[i] = grad[i] * (1.0f - y[i] * y[i]);
    }
}

void orc_mlp_forward(const float* params, const int* sizes, const int* acts, int num_layers,
                     const float* x, int m, float* cache) {
    memcpy(cache, x, (size_t)m * sizes[0] * sizeof(float));
    const float* p = params;
    float* in = cache;
    for (int i = 0; i < num_layers - 1; i++) {
        float* out = in + (size_t)m * sizes[i];
        const float* w = p;
        const float* b = p + sizes[i] * sizes[i + 1];
        orc_mat_mul(out, in, w, b, m, sizes[i], sizes[i + 1]);
        orc_act(out, m * sizes[i + 1], acts[i]);
        p = b + sizes[i + 1];
        in = out;
    }
}

void orc_mlp_backward(const float* params, const int* sizes, const int* acts, int num_layers,
                      const float* cache, const float* grad_out, int m, float* grads, float* grad_x0) {
    int L = num_layers - 1; /* number of weight layers */
    /* offsets */
    size_t poff[64], coff[65];
    size_t po = 0, co = 0;
    for (int i = 0; i < L; i++) {
        poff[i] = po;
        coff[i] = co;
        po += (size_t)sizes[i] * sizes[i + 1] + sizes[i + 1];
        co += (size_t)m * sizes[i];
    }
    coff[L] = co;

    float* layer_grad = (float*)malloc((size_t)m * sizes[L] * sizeof(float));
    memcpy(layer_grad, grad_out, (size_t)m * sizes[L] * sizeof(float));
    /* neural_network.cu:199-201: derivative of the last activation w.r.t. nn->output */
    orc_act_derivative(cache + coff[L], layer_grad, m * sizes[L], acts[L - 1]);

    for (int i = L - 1; i >= 0; i--) {
        int n = sizes[i], l = sizes[i + 1];
        const float* w = params + poff[i];
        const float* x = cache + coff[i];
        float* gw = grads + poff[i];
        float* gb = gw + (size_t)n * l;
        /* neural_network.cu:211-215: bias grad, sequential over rows */
        for (int j = 0; j < l; j++) {
            float s = 0.0f;
            for (int k = 0; k < m; k++) s += layer_grad[(size_t)k * l + j];
            gb[j] = s;
        }
        /* mat_mul.cu:57-80: grad_x = g . W ; grad_w = g^T . x (sequential sums from 0) */
        float* gx = (float*)malloc((size_t)m * n * sizeof(float));
#ifdef ORC_FAST /* timing-only build: axpy forms (unit-stride inner loops over j), same products, different summation order */
        memset(gx, 0, (size_t)m * n * sizeof(float));
        for (int r = 0; r < m; r++)
            for (int k = 0; k < l; k++) {
                const float g = layer_grad[(size_t)r * l + k];
                const float* wr = w + (size_t)k * n;
                float* o = gx + (size_t)r * n;
                for (int j = 0; j < n; j++) o[j] += g * wr[j];
            }
        memset(gw, 0, (size_t)l * n * sizeof(float));
        for (int k = 0; k < m; k++)
            for (int a = 0; a < l; a++) {
                const float g = layer_grad[(size_t)k * l + a];
                const float* xr = x + (size_t)k * n;
                float* o = gw + (size_t)a * n;
                for (int j = 0; j < n; j++) o[j] += g * xr[j];
            }
#else
        for (int r = 0; r < m; r++)
            for (int j = 0; j < n; j++) {
                float s = 0.0f;
                for (int k = 0; k < l; k++) s += layer_grad[(size_t)r * l + k] * w[(size_t)k * n + j];
                gx[(size_t)r * n + j] = s;
            }
        for (int a = 0; a < l; a++)
            for (int j = 0; j < n; j++) {
                float s = 0.0f;
                for (int k = 0; k < m; k++) s += layer_grad[(size_t)k * l + a] * x[(size_t)k * n + j];
                gw[(size_t)a * n + j] = s;
            }
#endif
        free(layer_grad);
        layer_grad = gx;
        /* neural_network.cu:224-226 */
        if (i > 0) orc_act_derivative(cache + coff[i], layer_grad, m * n, acts[i - 1]);
    }
    if (grad_x0) memcpy(grad_x0, layer_grad, (size_t)m * sizes[0] * sizeof(float));
    free(layer_grad);
}

/* ============================== GAE ============================================================ */

void orc_gae(const float* reward, const float* v, const float* v_next, const uint8_t* terminated,
             const uint8_t* truncated, int n, float gamma, float lambda, float* adv_raw,
             float* adv_target, float* adv_norm, float* mean_out, float* std_out) {
    float* adv = (float*)malloc(((size_t)n + 1) * sizeof(float));
    adv[n] = 0.0f;
    float sum = 0;
    /* ppo.cu:340-349 (delta folded into the reverse loop; same values) */
    for (int i = n - 1; i >= 0; i--) {
        float delta = reward[i] + gamma * v_next[i] * !terminated[i] - v[i];
        adv[i] = delta + gamma * lambda * !(truncated[i] || terminated[i]) * adv[i + 1];
        sum += adv[i];
    }
    /* ppo.cu:351-353 */
    for (int i = 0; i < n; i++) adv_target[i] = v[i] + adv[i];
    if (adv_raw) memcpy(adv_raw, adv, (size_t)n * sizeof(float));
    /* ppo.cu:355-368: float accumulators, each square formed in double */
    float mean = sum / n;
    float std = 0;
    for (int i = 0; i < n; i++) std += pow(adv[i] - mean, 2);
    std = sqrtf(std / n);
    if (adv_norm)
        for (int i = 0; i < n; i++) adv_norm[i] = (adv[i] - mean) / (std + 1e-8);
    if (mean_out) *mean_out = mean;
    if (std_out) *std_out = std;
    free(adv);
}

void orc_gae_f64(const float* reward, const float* v, const float* v_next, const uint8_t* terminated,
                 const uint8_t* truncated, int n, float gamma, float lambda, double* adv_raw,
                 double* adv_target, double* adv_norm, double* mean_out, double* std_out) {
    double* adv = (double*)malloc(((size_t)n + 1) * sizeof(double));
    adv[n] = 0.0;
    double g = gamma, gl = (double)gamma * (double)lambda, sum = 0.0;
    for (int i = n - 1; i >= 0; i--) {
        double delta = (double)reward[i] + g * v_next[i] * !terminated[i] - (double)v[i];
        adv[i] = delta + gl * !(truncated[i] || terminated[i]) * adv[i + 1];
        sum += adv[i];
    }
    double mean = sum / n, ss = 0.0;
    for (int i = 0; i < n; i++) ss += (adv[i] - mean) * (adv[i] - mean);
    double std = sqrt(ss / n);
    for (int i = 0; i < n; i++) {
        if (adv_target) adv_target[i] = (double)v[i] + adv[i];
        if (adv_raw) adv_raw[i] = adv[i];
        if (adv_norm) adv_norm[i] = (adv[i] - mean) / (std + 1e-8);
    }
    if (mean_out) *mean_out = mean;
    if (std_out) *std_out = std;
    free(adv);
}

void orc_welford_combine(const float* means, const float* m2s, const int* ns, int k, float* mean,
                         float* m2, int* n) {
    float smean = 0, sm2 = 0;
    int sn = 0;
    for (int i = 0; i < k; i++) { /* welford_var.h:58-66 */
        float delta = means[i] - smean;
        int nn = sn + ns[i];
        float nmean = smean + delta * ns[i] / nn;
        float nm2 = sm2 + m2s[i] + delta * delta * sn * ns[i] / nn;
        smean = nmean;
        sm2 = nm2;
        sn = nn;
    }
    *mean = smean;
    *m2 = sm2;
    *n = sn;
}

/* ============================== permutation / gather =========================================== */

void orc_shuffle(int* idx, int limit) {
    for (int i = 0; i < limit; i++) idx[i] = i;
    for (int i = 0; i < limit; i++) {
        int j = rand() % limit;
        int t = idx[i];
        idx[i] = idx[j];
        idx[j] = t;
    }
}

void orc_get_batch(const int* random_idx, int limit, int batch_idx, int batch_size, int S, int A,
                   const float* state, const float* action, const float* logprob,
                   const float* advantage, const float* adv_target, float* states, float* actions,
                   float* logprobs, float* advantages, float* adv_targets) {
    int offset = batch_idx * batch_size;
    for (int i = 0; i < batch_size; i++) {
        int idx = random_idx[(offset + i) % limit];
        for (int j = 0; j < S; j++) states[(size_t)i * S + j] = state[(size_t)idx * S + j];
        for (int j = 0; j < A; j++) actions[(size_t)i * A + j] = action[(size_t)idx * A + j];
        logprobs[i] = logprob[idx];
        advantages[i] = advantage[idx];
        adv_targets[i] = adv_target[idx];
    }
}

/* ============================== Gaussian policy / losses ======================================= */

void orc_gaussian_noise(float* out, int n) {
    if (n == 1) { /* policy.cu:48-51: first draw feeds the log, second the cos (probed) */
        float u1 = (float)rand() / RAND_MAX;
        float r = sqrtf(-2 * logf(u1));
        float c = cosf(2 * ORC_PI * (float)rand() / RAND_MAX);
        out[0] = r * c;
        return;
    }
    for (int i = 0; i <= n / 2; i += 2) { /* policy.cu:53-60, loop bound reproduced as written */
        float u1 = (float)rand() / RAND_MAX;
        float u2 = (float)rand() / RAND_MAX;
        float r = sqrtf(-2 * logf(u1));
        float theta = 2 * ORC_PI * u2;
        out[i] = r * cosf(theta);
        out[i + 1] = r * sinf(theta);
    }
    if (n % 2 == 1) {
        float u1 = (float)rand() / RAND_MAX;
        float r = sqrtf(-2 * logf(u1));
        out[n - 1] = r * cosf(2 * ORC_PI * (float)rand() / RAND_MAX);
    }
}

float orc_log_prob_one(const float* mu, const float* log_std, const float* action, int A) {
    float logprob = -0.5 * A * logf(2 * ORC_PI);
    for (int i = 0; i < A; i++)
        logprob -= log_std[i] + 0.5 * powf((action[i] - mu[i]) / expf(log_std[i]), 2);
    return logprob;
}

void orc_log_prob(const float* mu, const float* log_std, const float* action, int m, int A, float* out) {
    for (int i = 0; i < m; i++) out[i] = orc_log_prob_one(mu + (size_t)i * A, log_std, action + (size_t)i * A, A);
}

void orc_log_prob_backwards(const float* mu, const float* log_std, const float* action,
                            const float* grad_in, int m, int A, int ref_index, float* grad_mu,
                            float* grad_log_std) {
    memset(grad_log_std, 0, (size_t)A * sizeof(float));
    for (int i = 0; i < m; i++)
        for (int j = 0; j < A; j++) {
            size_t ij = (size_t)i * A + j;
            float g = ref_index ? grad_in[ij] : grad_in[i];
            grad_mu[ij] = (action[ij] - mu[ij]) * expf(-2 * log_std[j]) * g;
            grad_log_std[j] += (-1 + powf(action[ij] - mu[ij], 2) * expf(-2 * log_std[j])) * g;
        }
}

float orc_entropy(const float* log_std, int A) {
    float entropy = A * 0.5 * (1 + log(2 * ORC_PI));
    for (int j = 0; j < A; j++) entropy += log_std[j];
    return entropy;
}

float orc_policy_loss_and_grad(float* grad_logprob, float* grad_entropy, const float* adv,
                               const float* logprobs, const float* old_logprobs, float entropy,
                               float ent_coeff, float epsilon, int m) {
    float loss = 0;
    for (int i = 0; i < m; i++) {
        float ratio = expf(logprobs[i] - old_logprobs[i]); /* C++ exp(float) -> float overload */
        int adv_pos = adv[i] > 0;
        int ratio_pos = ratio > 1 + epsilon;
        int ratio_neg = ratio < 1 - epsilon;
        loss -= adv[i] * (adv_pos * (ratio_pos * (1 + epsilon) + !ratio_pos * ratio) +
                          !adv_pos * (ratio_neg * (1 - epsilon) + !ratio_neg * ratio));
        grad_logprob[i] = -(adv_pos * !ratio_pos + !adv_pos * !ratio_neg) * adv[i] * ratio / m;
    }
    loss /= m;
    loss -= ent_coeff * entropy;
    *grad_entropy = -ent_coeff;
    return loss;
}

float orc_mse(const float* y, const float* y_true, int m, int n) {
    float loss = 0.0;
    for (int i = 0; i < m * n; i++) loss += pow(y_true[i] - y[i], 2);
    return loss / (m * n);
}

void orc_mse_derivative(float* grad, const float* y, const float* y_true, int m, int n) {
    for (int i = 0; i < m * n; i++) grad[i] = 2 * (y[i] - y_true[i]) / (m * n);
}

/* ============================== Adam =========================================================== */

void orc_adam(float* w, const float* g, float* m, float* v, int n, float lr, float beta1,
              float beta2, int* time_step) {
    *time_step += 1;
    float bc1 = 1 - powf(beta1, *time_step);
    float bc2 = 1 - powf(beta2, *time_step);
    float step_size = lr / bc1;
    for (int i = 0; i < n; i++) {
        m[i] = beta1 * m[i] + (1 - beta1) * g[i];
        v[i] = beta2 * v[i] + (1 - beta2) * powf(g[i], 2);
        float denom = sqrtf(v[i] / bc2) + 1e-8;
        w[i] -= step_size * m[i] / denom;
    }
}

/* ============================== environments =================================================== */

static double orc_angle_normalize(double x) {
    /* ((x + pi) mod 2pi) - pi with Python's non-negative modulo */
    double y = fmod(x + ORC_PI, 2 * ORC_PI);
    if (y < 0) y += 2 * ORC_PI;
    return y - ORC_PI;
}

void orc_pendulum_obs(double theta, double theta_dot, float* obs) {
    obs[0] = (float)cos(theta);
    obs[1] = (float)sin(theta);
    obs[2] = (float)theta_dot;
}

void orc_pendulum_step(double* theta, double* theta_dot, float action, float* obs, float* reward) {
    const double g = 10.0, mass = 1.0, l = 1.0, dt = 0.05, max_speed = 8.0, max_torque = 2.0;
    double u = action;
    if (u > max_torque) u = max_torque;
    if (u < -max_torque) u = -max_torque;
    double th = *theta, thd = *theta_dot;
    double an = orc_angle_normalize(th);
    double cost = an * an + 0.1 * thd * thd + 0.001 * u * u;
    double nthd = thd + (3 * g / (2 * l) * sin(th) + 3.0 / (mass * l * l) * u) * dt;
    if (nthd > max_speed) nthd = max_speed;
    if (nthd < -max_speed) nthd = -max_speed;
    double nth = th + nthd * dt;
    *theta = nth;
    *theta_dot = nthd;
    orc_pendulum_obs(nth, nthd, obs);
    *reward = (float)(-cost);
}

/* ============================== update phase =================================================== */

float orc_update(const OrcConfig* cfg, OrcModel* model, OrcBuffer* buf, int* perm_log, float* loss_log) {
    const int n = buf->n, S = cfg->S, A = cfg->A, mb = cfg->batch_size, NL = cfg->num_layers;
    const int P_mu = orc_param_count(cfg->sizes_mu, NL), P_v = orc_param_count(cfg->sizes_v, NL);
    float last_loss = 0;

    /* ---- compute_gae, ppo.cu:326-369 ---- */
    {
        float* cache = (float*)malloc((size_t)orc_cache_floats(cfg->sizes_v, NL, n) * sizeof(float));
        float* vnext = (float*)malloc((size_t)n * sizeof(float));
        float* v = (float*)malloc((size_t)n * sizeof(float));
        orc_mlp_forward(model->v, cfg->sizes_v, cfg->acts, NL, buf->next_state, n, cache);
        memcpy(vnext, orc_mlp_output(cache, cfg->sizes_v, NL, n), (size_t)n * sizeof(float));
        orc_mlp_forward(model->v, cfg->sizes_v, cfg->acts, NL, buf->state, n, cache);
        memcpy(v, orc_mlp_output(cache, cfg->sizes_v, NL, n), (size_t)n * sizeof(float));
        orc_gae(buf->reward, v, vnext, buf->terminated, buf->truncated, n, cfg->gamma, cfg->lambda,
                NULL, buf->adv_target, buf->advantage, NULL, NULL);
        free(cache);
        free(vnext);
        free(v);
    }

    int num_batches = n / mb; /* ppo.cu:387-388: ceilf of an integer quotient */
    int* idx = (int*)malloc((size_t)n * sizeof(int));
    float* states = (float*)malloc((size_t)mb * S * sizeof(float));
    float* actions = (float*)malloc((size_t)mb * A * sizeof(float));
    float* lp_old = (float*)malloc((size_t)mb * sizeof(float));
    float* lp = (float*)malloc((size_t)mb * sizeof(float));
    float* adv = (float*)malloc((size_t)mb * sizeof(float));
    float* advt = (float*)malloc((size_t)mb * sizeof(float));
    float* gl = (float*)malloc((size_t)mb * sizeof(float));
    float* gmu = (float*)malloc((size_t)mb * A * sizeof(float));
    float* cache_v = (float*)malloc((size_t)orc_cache_floats(cfg->sizes_v, NL, mb) * sizeof(float));
    float* cache_mu = (float*)malloc((size_t)orc_cache_floats(cfg->sizes_mu, NL, mb) * sizeof(float));
    float* grads_v = (float*)malloc((size_t)P_v * sizeof(float));
    float* grads_mu = (float*)malloc((size_t)P_mu * sizeof(float));
    float* g_ls = (float*)malloc((size_t)A * sizeof(float));
    int perm_count = 0, loss_count = 0;

    /* ---- value epochs, ppo.cu:398-417 ---- */
    for (int j = 0; j < cfg->n_epochs_value; j++) {
        orc_shuffle(idx, n);
        if (perm_log) memcpy(perm_log + (size_t)(perm_count++) * n, idx, (size_t)n * sizeof(int));
        for (int k = 0; k < num_batches; k++) {
            orc_get_batch(idx, n, k, mb, S, A, buf->state, buf->action, buf->logprob, buf->advantage,
                          buf->adv_target, states, actions, lp_old, adv, advt);
            orc_mlp_forward(model->v, cfg->sizes_v, cfg->acts, NL, states, mb, cache_v);
            const float* out = orc_mlp_output(cache_v, cfg->sizes_v, NL, mb);
            float v_loss = orc_mse(out, advt, mb, 1);
            if (loss_log) loss_log[loss_count++] = v_loss;
            orc_mse_derivative(gl, out, advt, mb, 1);
            orc_mlp_backward(model->v, cfg->sizes_v, cfg->acts, NL, cache_v, gl, mb, grads_v, NULL);
            orc_adam(model->v, grads_v, model->m_v, model->v_v, P_v, cfg->lr_v, 0.9f, 0.999f, &model->t_v);
        }
    }
    /* ---- policy epochs, ppo.cu:419-444 ---- */
    for (int j = 0; j < cfg->n_epochs_policy; j++) {
        orc_shuffle(idx, n);
        if (perm_log) memcpy(perm_log + (size_t)(perm_count++) * n, idx, (size_t)n * sizeof(int));
        for (int k = 0; k < num_batches; k++) {
            orc_get_batch(idx, n, k, mb, S, A, buf->state, buf->action, buf->logprob, buf->advantage,
                          buf->adv_target, states, actions, lp_old, adv, advt);
            orc_mlp_forward(model->mu, cfg->sizes_mu, cfg->acts, NL, states, mb, cache_mu);
            const float* mu = orc_mlp_output(cache_mu, cfg->sizes_mu, NL, mb);
            orc_log_prob(mu, model->log_std, actions, mb, A, lp);
            float entropy = orc_entropy(model->log_std, A);
            float entropy_grad;
            last_loss = orc_policy_loss_and_grad(gl, &entropy_grad, adv, lp, lp_old, entropy,
                                                 cfg->ent_coeff, cfg->epsilon, mb);
            if (loss_log) loss_log[loss_count++] = last_loss;
            orc_log_prob_backwards(mu, model->log_std, actions, gl, mb, A, cfg->ref_index, gmu, g_ls);
            orc_mlp_backward(model->mu, cfg->sizes_mu, cfg->acts, NL, cache_mu, gmu, mb, grads_mu, NULL);
            for (int i = 0; i < A; i++) g_ls[i] += entropy_grad; /* ppo.cu:436-438 */
            orc_adam(model->log_std, g_ls, model->m_ls, model->v_ls, A, cfg->lr_policy, 0.9f, 0.999f, &model->t_ls);
            orc_adam(model->mu, grads_mu, model->m_mu, model->v_mu, P_mu, cfg->lr_policy, 0.9f, 0.999f, &model->t_mu);
        }
    }
    free(idx); free(states); free(actions); free(lp_old); free(lp); free(adv); free(advt);
    free(gl); free(gmu); free(cache_v); free(cache_mu); free(grads_v); free(grads_mu); free(g_ls);
    return last_loss;
}

/* ============================== rollout ======================================================== */

/* toy env, env.c:6-33 */
static float toy_state;
static int toy_step;
static double pend_th, pend_thd;
static int pend_step;

static void env_reset(int env_id, float* obs) {
    if (env_id == 0) {
        toy_state = 0;
        toy_step = 0;
        obs[0] = 0;
    } else { /* [EXT] theta ~ U(-pi,pi), theta_dot ~ U(-1,1), drawn from rand() in that order */
        double u1 = (double)rand() / RAND_MAX, u2 = (double)rand() / RAND_MAX;
        pend_th = (2 * u1 - 1) * ORC_PI;
        pend_thd = (2 * u2 - 1);
        pend_step = 0;
        orc_pendulum_obs(pend_th, pend_thd, obs);
    }
}

static void env_step(int env_id, const float* action, float* obs, float* reward, uint8_t* term, uint8_t* trunc) {
    if (env_id == 0) {
        toy_state += fmaxf(fminf(action[0], 1), -1);
        obs[0] = toy_state;
        toy_step += 1;
        if (toy_state >= 5) { *reward = 1; *term = 1; *trunc = 0; }
        else if (toy_step >= 15) { *reward = 0; *term = 0; *trunc = 1; }
        else { *reward = 0; *term = 0; *trunc = 0; }
    } else {
        orc_pendulum_step(&pend_th, &pend_thd, action[0], obs, reward);
        pend_step += 1;
        *term = 0;
        *trunc = pend_step >= 200;
    }
}

int orc_collect(const OrcConfig* cfg, const OrcModel* model, OrcBuffer* buf, int capacity,
                int start_idx, int steps, int env_id) {
    const int S = cfg->S, A = cfg->A, NL = cfg->num_layers;
    float* cache = (float*)malloc((size_t)orc_cache_floats(cfg->sizes_mu, NL, 1) * sizeof(float));
    float noise[64];
    int idx = start_idx;
    env_reset(env_id, buf->state + (size_t)idx * S); /* ppo.cu:55 */
    for (int i = 0; i < steps; i++) {
        /* sample_action, policy.cu:76-89 (m = 1) */
        orc_mlp_forward(model->mu, cfg->sizes_mu, cfg->acts, NL, buf->state + (size_t)idx * S, 1, cache);
        const float* mu = orc_mlp_output(cache, cfg->sizes_mu, NL, 1);
        orc_gaussian_noise(noise, A);
        float* act = buf->action + (size_t)idx * A;
        for (int j = 0; j < A; j++) act[j] = mu[j] + noise[j] * expf(model->log_std[j]);
        buf->logprob[idx] = orc_log_prob_one(mu, model->log_std, act, A);
        env_step(env_id, act, buf->next_state + (size_t)idx * S, buf->reward + idx,
                 buf->terminated + idx, buf->truncated + idx);
        int new_idx = (idx + 1) % capacity;
        if (i < steps - 1) { /* ppo.cu:64-69 */
            if (buf->truncated[idx] || buf->terminated[idx]) env_reset(env_id, buf->state + (size_t)new_idx * S);
            else memcpy(buf->state + (size_t)new_idx * S, buf->next_state + (size_t)idx * S, (size_t)S * sizeof(float));
        } else if (!buf->terminated[idx]) { /* ppo.cu:70-74 */
            buf->truncated[idx] = 1;
        }
        idx = new_idx;
    }
    free(cache);
    return idx;
}

/* eval_ppo, ppo.cu:560-583: reset_buffer, collect `steps` transitions from index 0, then the reverse walk that
 * accumulates the undiscounted reward sum and the per-episode discounted return J (float arithmetic, in the
 * reference's order; the episode that ends at steps-1 is counted in n_episodes but its J is only added when an
 * earlier done flag is met, exactly as the reference does). */
void orc_eval(const OrcConfig* cfg, const OrcModel* model, OrcBuffer* buf, int capacity, int steps, int env_id,
              float gamma, float* J_out, float* R_out, int* episodes_out) {
    orc_collect(cfg, model, buf, capacity, 0, steps, env_id);
    float rewards = buf->reward[steps - 1];
    float episode_J = buf->reward[steps - 1];
    int n_episodes = 1;
    float sum_J = 0;
    for (int i = steps - 2; i >= 0; i--) {
        rewards += buf->reward[i];
        episode_J = buf->reward[i] + gamma * episode_J;
        if (buf->terminated[i] || buf->truncated[i]) {
            n_episodes++;
            sum_J += episode_J;
            episode_J = 0;
        }
    }
    *J_out = sum_J / n_episodes;
    *R_out = rewards / n_episodes;
    *episodes_out = n_episodes;
}

/* [EXT] The Pendulum above behind the reference's Env hook signatures (include/env.h:7-15), so the UNMODIFIED reference
 * (oracle/_ref) can be driven through its own train_ppo_epoch on the benchmark's env: bench.py's C1 line and
 * tests/test_oracle_vs_ref.py.  Same state, same rand() draws at reset as env_id 1 of orc_collect. */
void orc_pendulum_hook_reset(float* obs) { env_reset(1, obs); }
void orc_pendulum_hook_step(float* action, float* obs, float* reward, _Bool* terminated, _Bool* truncated, int action_size) {
    (void)action_size;
    uint8_t te = 0, tr = 0;
    env_step(1, action, obs, reward, &te, &tr);
    *terminated = te != 0;
    *truncated = tr != 0;
}
void orc_pendulum_hook_free(void) {}
