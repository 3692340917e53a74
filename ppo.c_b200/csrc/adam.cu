// adam.cu — Adam as ONE streaming launch over a flat parameter vector.
//
// Replaces adam_update / adam_update_cuda + K15 (reference src/adam.cu:53-74, 138-169: one thread
// per element with a linear search over layers and double pointer indirection).  Networks created
// by this library keep all tensors of a net in one contiguous arena (W0,b0,W1,b1,..., the order of
// src/adam.cu:25-42), so the optimiser is a pure 128-bit streaming kernel:
//   read g,m,v,w + write m,v,w = 28 B/parameter.
// Arbitrary tensor lists passed to create_adam_cuda still work through a chunk table (no per-thread
// search).  The arithmetic is the reference's, operation by operation, with contraction disabled
// (__fmul_rn/__fadd_rn) so results match the x86 build bit for bit:
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; denom = sqrtf(v/bc2) + 1e-8 (double add) ;
//   w -= (lr/bc1)*m/denom                                               (src/adam.cu:56-69)
// Optionally the gradient is first formed as a fixed-order sum of split-K slabs (deterministic
// replacement for the float atomics of src/policy.cu:157), fusing the reduction into the update.
#include <unordered_map>

#include "common.cuh"
#include "internal.h"

namespace b200 {

struct AdamScalars {
    float beta1, beta2, omb1, omb2, bc2, step_size;
};

__device__ __forceinline__ void adam_one(float& w, float g, float& m, float& v, const AdamScalars& s) {
    m = __fadd_rn(__fmul_rn(s.beta1, m), __fmul_rn(s.omb1, g));
    v = __fadd_rn(__fmul_rn(s.beta2, v), __fmul_rn(s.omb2, __fmul_rn(g, g)));
    const float denom = (float)((double)__fsqrt_rn(__fdiv_rn(v, s.bc2)) + 1e-8);
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(s.step_size, m), denom));
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, int n, AdamScalars s, const float* __restrict__ partials,
                 int splits, size_t stride, int vec_ok) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n4 = vec_ok ? n / 4 : 0;
    for (long long i = tid; i < n4; i += nthreads) {
        float4 g4;
        if (partials) {
            g4 = *reinterpret_cast<const float4*>(partials + 4 * i);
            for (int k = 1; k < splits; k++) {
                const float4 p = *reinterpret_cast<const float4*>(partials + (size_t)k * stride + 4 * i);
                g4.x = __fadd_rn(g4.x, p.x); g4.y = __fadd_rn(g4.y, p.y);
                g4.z = __fadd_rn(g4.z, p.z); g4.w = __fadd_rn(g4.w, p.w);
            }
            *reinterpret_cast<float4*>(g + 4 * i) = g4;
        } else {
            g4 = *reinterpret_cast<const float4*>(g + 4 * i);
        }
        float4 w4 = *reinterpret_cast<float4*>(w + 4 * i);
        float4 m4 = *reinterpret_cast<float4*>(m + 4 * i);
        float4 v4 = *reinterpret_cast<float4*>(v + 4 * i);
        adam_one(w4.x, g4.x, m4.x, v4.x, s); adam_one(w4.y, g4.y, m4.y, v4.y, s);
        adam_one(w4.z, g4.z, m4.z, v4.z, s); adam_one(w4.w, g4.w, m4.w, v4.w, s);
        *reinterpret_cast<float4*>(w + 4 * i) = w4;
        *reinterpret_cast<float4*>(m + 4 * i) = m4;
        *reinterpret_cast<float4*>(v + 4 * i) = v4;
    }
    for (long long i = 4 * n4 + tid; i < n; i += nthreads) {
        float gi;
        if (partials) {
            gi = partials[i];
            for (int k = 1; k < splits; k++) gi = __fadd_rn(gi, partials[(size_t)k * stride + i]);
            g[i] = gi;
        } else {
            gi = g[i];
        }
        float wi = w[i], mi = m[i], vi = v[i];
        adam_one(wi, gi, mi, vi, s);
        w[i] = wi; m[i] = mi; v[i] = vi;
    }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(float* __restrict__ g, const float* __restrict__ partials, int splits,
                       size_t stride, int n) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        float gi = partials[i];
        for (int k = 1; k < splits; k++) gi = __fadd_rn(gi, partials[(size_t)k * stride + i]);
        g[i] = gi;
    }
}

struct AdamChunk { float* w; const float* g; int moff; int len; };

__global__ void __launch_bounds__(256)
adam_chunked_kernel(const AdamChunk* __restrict__ chunks, float* __restrict__ m, float* __restrict__ v, AdamScalars s) {
    const AdamChunk c = chunks[blockIdx.x];
    for (int i = threadIdx.x; i < c.len; i += blockDim.x) {
        float wi = c.w[i], mi = m[c.moff + i], vi = v[c.moff + i];
        adam_one(wi, c.g[i], mi, vi, s);
        c.w[i] = wi; m[c.moff + i] = mi; v[c.moff + i] = vi;
    }
}

static AdamScalars make_scalars(float lr, float beta1, float beta2, int time_step) {
    // host-side scalars exactly as src/adam.cu:56-59 / :158-161 (the reference computes them on the
    // host in both twins)
    AdamScalars s;
    const float bc1 = 1 - powf(beta1, time_step);
    s.bc2 = 1 - powf(beta2, time_step);
    s.step_size = lr / bc1;
    s.beta1 = beta1; s.beta2 = beta2;
    s.omb1 = 1 - beta1; s.omb2 = 1 - beta2;
    return s;
}

void adam_flat(float* w, float* g, float* m, float* v, int n, float lr, float beta1, float beta2,
               int time_step, const float* partials, int splits, size_t stride) {
    if (n <= 0) return;
    const AdamScalars s = make_scalars(lr, beta1, beta2, time_step);
    uintptr_t al = (uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)partials | (uintptr_t)(stride * 4);
    const int vec_ok = (al & 15) == 0;
    const int blocks = (int)std::min<long long>(div_up(div_up(n, 4), 256), (long long)num_sms() * 8);
    B200_LAUNCH(adam_flat_kernel, std::max(blocks, 1), 256, 0, w, g, m, v, n, s, partials, splits, stride, vec_ok);
}

void reduce_partials(float* g, const float* partials, int splits, size_t stride, int n) {
    if (n <= 0) return;
    const int blocks = (int)std::min<long long>(div_up(n, 256), (long long)num_sms() * 8);
    B200_LAUNCH(reduce_partials_kernel, blocks, 256, 0, g, partials, splits, stride, n);
}

// ---- reference API (include/adam.h) ---------------------------------------------------------------
struct AdamMeta {
    bool cuda = false;
    std::vector<float*> w, g;    // per-tensor pointers (host or device, per `cuda`)
    std::vector<int> len;
    bool contiguous = false;
    AdamChunk* d_chunks = nullptr;
    int n_chunks = 0;
};
static std::unordered_map<Adam*, AdamMeta> g_adam_meta;

static void build_meta(Adam* adam, float** weights, float** grads, int* length, int num_layers, bool cuda) {
    AdamMeta meta;
    meta.cuda = cuda;
    meta.w.assign(weights, weights + num_layers);
    meta.g.assign(grads, grads + num_layers);
    meta.len.assign(length, length + num_layers);
    meta.contiguous = true;
    for (int i = 0; i + 1 < num_layers; i++)
        if (meta.w[i] + meta.len[i] != meta.w[i + 1] || meta.g[i] + meta.len[i] != meta.g[i + 1]) meta.contiguous = false;
    if (cuda && !meta.contiguous) {
        std::vector<AdamChunk> chunks;
        int moff = 0;
        for (int i = 0; i < num_layers; i++) {
            for (int o = 0; o < meta.len[i]; o += 2048)
                chunks.push_back({meta.w[i] + o, meta.g[i] + o, moff + o, std::min(2048, meta.len[i] - o)});
            moff += meta.len[i];
        }
        meta.n_chunks = (int)chunks.size();
        meta.d_chunks = dmalloc<AdamChunk>(chunks.size());
        CUDA_CHECK(cudaMemcpy(meta.d_chunks, chunks.data(), chunks.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice));
    }
    g_adam_meta[adam] = std::move(meta);
}

}  // namespace b200

using namespace b200;

extern "C" {

void ppo_b200_adam_flat(float* w, const float* g, float* m, float* v, int n, float lr, float beta1,
                        float beta2, int time_step) {
    adam_flat(w, const_cast<float*>(g), m, v, n, lr, beta1, beta2, time_step, nullptr, 0, 0);
}

// ---- device twins: src/adam.cu:76-169 ---------------------------------------------------------------
Adam* create_adam_cuda(float** weights, float** grad_weights, int* length, int num_layers, int size,
                       float beta1, float beta2) {
    ensure_device();
    Adam* adam = (Adam*)malloc(sizeof(Adam));
    adam->m = dmalloc<float>(size);
    adam->v = dmalloc<float>(size);
    CUDA_CHECK(cudaMemsetAsync(adam->m, 0, (size_t)size * sizeof(float), stream()));
    CUDA_CHECK(cudaMemsetAsync(adam->v, 0, (size_t)size * sizeof(float), stream()));
    // device-side tables kept for ABI fidelity (src/adam.cu:84-97: lengths hold inclusive prefix sums)
    adam->weights = dmalloc<float*>(num_layers);
    adam->grad_weights = dmalloc<float*>(num_layers);
    adam->lengths = dmalloc<int>(num_layers);
    std::vector<int> presum(num_layers);
    int acc = 0;
    for (int i = 0; i < num_layers; i++) { acc += length[i]; presum[i] = acc; }
    CUDA_CHECK(cudaMemcpy(adam->weights, weights, num_layers * sizeof(float*), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(adam->grad_weights, grad_weights, num_layers * sizeof(float*), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(adam->lengths, presum.data(), num_layers * sizeof(int), cudaMemcpyHostToDevice));
    adam->size = size;
    adam->beta1 = beta1;
    adam->beta2 = beta2;
    adam->time_step = 0;
    adam->num_layers = num_layers;
    build_meta(adam, weights, grad_weights, length, num_layers, true);
    return adam;
}

Adam* create_adam_from_nn_cuda(NeuralNetwork* nn, float beta1, float beta2) {
    const int L = nn->num_layers - 1;
    std::vector<float*> w(2 * L), g(2 * L);
    std::vector<int> len(2 * L);
    int size = 0;
    for (int i = 0; i < L; i++) {
        w[2 * i] = nn->layers[i].d_weights;       w[2 * i + 1] = nn->layers[i].d_biases;
        g[2 * i] = nn->layers[i].d_grad_weights;  g[2 * i + 1] = nn->layers[i].d_grad_biases;
        len[2 * i] = nn->layers[i].input_size * nn->layers[i].output_size;
        len[2 * i + 1] = nn->layers[i].output_size;
        size += len[2 * i] + len[2 * i + 1];
    }
    return create_adam_cuda(w.data(), g.data(), len.data(), 2 * L, size, beta1, beta2);
}

void free_adam_cuda(Adam* adam) {
    if (!adam) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    auto it = g_adam_meta.find(adam);
    if (it != g_adam_meta.end()) {
        if (it->second.d_chunks) CUDA_CHECK(cudaFree(it->second.d_chunks));
        g_adam_meta.erase(it);
    }
    CUDA_CHECK(cudaFree(adam->m));
    CUDA_CHECK(cudaFree(adam->v));
    CUDA_CHECK(cudaFree(adam->weights));
    CUDA_CHECK(cudaFree(adam->grad_weights));
    CUDA_CHECK(cudaFree(adam->lengths));
    free(adam);
}

void adam_update_cuda(Adam* adam, float lr) {
    auto it = g_adam_meta.find(adam);
    if (it == g_adam_meta.end()) B200_FATAL("adam_update_cuda: Adam %p was not created by create_adam*_cuda", (void*)adam);
    AdamMeta& meta = it->second;
    adam->time_step += 1;
    for (float* w : meta.w) net_mark_params_written(w);
    if (meta.contiguous) {
        adam_flat(meta.w[0], meta.g[0], adam->m, adam->v, adam->size, lr, adam->beta1, adam->beta2,
                  adam->time_step, nullptr, 0, 0);
    } else {
        const AdamScalars s = make_scalars(lr, adam->beta1, adam->beta2, adam->time_step);
        B200_LAUNCH(adam_chunked_kernel, meta.n_chunks, 256, 0, meta.d_chunks, adam->m, adam->v, s);
    }
}

// ---- host-pointer twins (src/adam.cu:6-74): same kernel, staged through device scratch -----------
Adam* create_adam(float** weights, float** grad_weights, int* length, int num_layers, int size,
                  float beta1, float beta2) {
    Adam* adam = (Adam*)malloc(sizeof(Adam));
    adam->m = (float*)calloc(size, sizeof(float));
    adam->v = (float*)calloc(size, sizeof(float));
    adam->weights = (float**)malloc(num_layers * sizeof(float*));
    adam->grad_weights = (float**)malloc(num_layers * sizeof(float*));
    adam->lengths = (int*)malloc(num_layers * sizeof(int));
    memcpy(adam->weights, weights, num_layers * sizeof(float*));
    memcpy(adam->grad_weights, grad_weights, num_layers * sizeof(float*));
    memcpy(adam->lengths, length, num_layers * sizeof(int));
    adam->size = size;
    adam->beta1 = beta1;
    adam->beta2 = beta2;
    adam->time_step = 0;
    adam->num_layers = num_layers;
    build_meta(adam, weights, grad_weights, length, num_layers, false);
    return adam;
}

Adam* create_adam_from_nn(NeuralNetwork* nn, float beta1, float beta2) {
    const int L = nn->num_layers - 1;
    std::vector<float*> w(2 * L), g(2 * L);
    std::vector<int> len(2 * L);
    int size = 0;
    for (int i = 0; i < L; i++) {
        w[2 * i] = nn->layers[i].weights;       w[2 * i + 1] = nn->layers[i].biases;
        g[2 * i] = nn->layers[i].grad_weights;  g[2 * i + 1] = nn->layers[i].grad_biases;
        len[2 * i] = nn->layers[i].input_size * nn->layers[i].output_size;
        len[2 * i + 1] = nn->layers[i].output_size;
        size += len[2 * i] + len[2 * i + 1];
    }
    return create_adam(w.data(), g.data(), len.data(), 2 * L, size, beta1, beta2);
}

void free_adam(Adam* adam) {
    if (!adam) return;
    g_adam_meta.erase(adam);
    free(adam->weights);
    free(adam->grad_weights);
    free(adam->lengths);
    free(adam->m);
    free(adam->v);
    free(adam);
}

void adam_update(Adam* adam, float lr) {
    adam->time_step += 1;
    const size_t n = adam->size;
    float* d = static_cast<float*>(scratch(kScratchStage, 4 * n * sizeof(float)));
    float *dw = d, *dg = d + n, *dm = d + 2 * n, *dv = d + 3 * n;
    size_t off = 0;
    for (int i = 0; i < adam->num_layers; i++) {
        const size_t len = adam->lengths[i];
        CUDA_CHECK(cudaMemcpyAsync(dw + off, adam->weights[i], len * sizeof(float), cudaMemcpyHostToDevice, stream()));
        CUDA_CHECK(cudaMemcpyAsync(dg + off, adam->grad_weights[i], len * sizeof(float), cudaMemcpyHostToDevice, stream()));
        off += len;
    }
    CUDA_CHECK(cudaMemcpyAsync(dm, adam->m, n * sizeof(float), cudaMemcpyHostToDevice, stream()));
    CUDA_CHECK(cudaMemcpyAsync(dv, adam->v, n * sizeof(float), cudaMemcpyHostToDevice, stream()));
    adam_flat(dw, dg, dm, dv, (int)n, lr, adam->beta1, adam->beta2, adam->time_step, nullptr, 0, 0);
    off = 0;
    for (int i = 0; i < adam->num_layers; i++) {
        const size_t len = adam->lengths[i];
        CUDA_CHECK(cudaMemcpyAsync(adam->weights[i], dw + off, len * sizeof(float), cudaMemcpyDeviceToHost, stream()));
        off += len;
    }
    CUDA_CHECK(cudaMemcpyAsync(adam->m, dm, n * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaMemcpyAsync(adam->v, dv, n * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

// ---- checkpoint blocks: byte format of src/adam.cu:172-264 -------------------------------------------
void save_adam(Adam* adam, FILE* file, bool cuda) {
    fwrite(&adam->size, sizeof(int), 1, file);
    fwrite(&adam->time_step, sizeof(int), 1, file);
    fwrite(&adam->beta1, sizeof(float), 1, file);
    fwrite(&adam->beta2, sizeof(float), 1, file);
    fwrite(&adam->num_layers, sizeof(int), 1, file);
    if (cuda) {
        std::vector<float> m(adam->size), v(adam->size);   // heap, not the reference's stack VLAs
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        CUDA_CHECK(cudaMemcpy(m.data(), adam->m, adam->size * sizeof(float), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(v.data(), adam->v, adam->size * sizeof(float), cudaMemcpyDeviceToHost));
        fwrite(m.data(), sizeof(float), adam->size, file);
        fwrite(v.data(), sizeof(float), adam->size, file);
    } else {
        fwrite(adam->m, sizeof(float), adam->size, file);
        fwrite(adam->v, sizeof(float), adam->size, file);
    }
}

static void must_read(void* dst, size_t size, size_t count, FILE* file) {
    if (fread(dst, size, count, file) != count) B200_FATAL("checkpoint truncated");
}

Adam* load_adam(FILE* file, float** weights, float** grad_weights, int* length, bool cuda) {
    int size, time_step, num_layers;
    float beta1, beta2;
    must_read(&size, sizeof(int), 1, file);
    must_read(&time_step, sizeof(int), 1, file);
    must_read(&beta1, sizeof(float), 1, file);
    must_read(&beta2, sizeof(float), 1, file);
    must_read(&num_layers, sizeof(int), 1, file);
    std::vector<float> m(size), v(size);
    must_read(m.data(), sizeof(float), size, file);
    must_read(v.data(), sizeof(float), size, file);
    Adam* adam = cuda ? create_adam_cuda(weights, grad_weights, length, num_layers, size, beta1, beta2)
                      : create_adam(weights, grad_weights, length, num_layers, size, beta1, beta2);
    adam->time_step = time_step;
    if (cuda) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        CUDA_CHECK(cudaMemcpy(adam->m, m.data(), size * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(adam->v, v.data(), size * sizeof(float), cudaMemcpyHostToDevice));
    } else {
        memcpy(adam->m, m.data(), size * sizeof(float));
        memcpy(adam->v, v.data(), size * sizeof(float));
    }
    return adam;
}

Adam* load_adam_from_nn(FILE* file, NeuralNetwork* nn, bool cuda) {
    const int L = nn->num_layers - 1;
    std::vector<float*> w(2 * L), g(2 * L);
    std::vector<int> len(2 * L);
    for (int i = 0; i < L; i++) {
        w[2 * i] = cuda ? nn->layers[i].d_weights : nn->layers[i].weights;
        w[2 * i + 1] = cuda ? nn->layers[i].d_biases : nn->layers[i].biases;
        g[2 * i] = cuda ? nn->layers[i].d_grad_weights : nn->layers[i].grad_weights;
        g[2 * i + 1] = cuda ? nn->layers[i].d_grad_biases : nn->layers[i].grad_biases;
        len[2 * i] = nn->layers[i].input_size * nn->layers[i].output_size;
        len[2 * i + 1] = nn->layers[i].output_size;
    }
    return load_adam(file, w.data(), g.data(), len.data(), cuda);
}

}  // extern "C"
