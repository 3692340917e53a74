// nn.cu — the NeuralNetwork object behind include/neural_network.h.
//
// Mirrors create/forward/backward/save/load of reference src/neural_network.cu, with a different
// memory design: all parameters of a net live in ONE device arena (W0,b0,W1,b1,...), gradients in a
// second one, `layers[i].d_*` are slices of them.  Activation buffers grow monotonically (the
// reference frees and reallocates them whenever the batch size changes, src/neural_network.cu:76-89).
#include <unordered_map>

#include "common.cuh"
#include "internal.h"

namespace b200 {

static std::unordered_map<NeuralNetwork*, NetDev*> g_nets;

NetDev* net_dev(NeuralNetwork* nn) {
    auto it = g_nets.find(nn);
    if (it == g_nets.end()) B200_FATAL("NeuralNetwork %p was not created by this library", (void*)nn);
    return it->second;
}

// Any writer of a parameter arena other than the fused Adam kernel must invalidate the staged image.
void net_mark_params_written(const float* params_ptr) {
    for (auto& kv : g_nets)
        if (params_ptr >= kv.second->params && params_ptr < kv.second->params + kv.second->param_count) kv.second->image_dirty = true;
}

static void attach_device(NeuralNetwork* nn) {
    ensure_device();
    NetDev* nd = new NetDev();
    const int L = nn->num_layers - 1;
    nd->num_layers = nn->num_layers;
    size_t off = 0;
    for (int i = 0; i < L; i++) {
        nd->sizes.push_back(nn->layers[i].input_size);
        nd->acts.push_back(act_code(nn->activation_functions[i]));
        nd->w_off.push_back(off);
        off += (size_t)nn->layers[i].input_size * nn->layers[i].output_size;
        nd->b_off.push_back(off);
        off += nn->layers[i].output_size;
    }
    nd->sizes.push_back(nn->output_size);
    nd->param_count = off;
    nd->params = dmalloc<float>(off);
    nd->grads = dmalloc<float>(off);
    CUDA_CHECK(cudaMemsetAsync(nd->grads, 0, off * sizeof(float), stream()));
    nd->a.assign(nn->num_layers, nullptr);
    nd->gx.assign(nn->num_layers, nullptr);
    for (int i = 0; i < L; i++) {
        nn->layers[i].d_weights = nd->params + nd->w_off[i];
        nn->layers[i].d_biases = nd->params + nd->b_off[i];
        nn->layers[i].d_grad_weights = nd->grads + nd->w_off[i];
        nn->layers[i].d_grad_biases = nd->grads + nd->b_off[i];
        nn->layers[i].d_activation_function = build_activation_function_cuda(nn->activation_functions[i]);
        nn->layers[i].d_input = nullptr;
        nn->layers[i].d_grad_x = nullptr;
    }
    nn->layers[L].d_input = nullptr;
    nn->layers[L].d_grad_x = nullptr;
    nn->d_output = nullptr;
    nn->cache_m_forward = 0;
    nn->cache_m_backward = 0;
    nn->cublas_handle = nullptr;
    g_nets[nn] = nd;
}

static void ensure_fwd_capacity(NeuralNetwork* nn, NetDev* nd, int m) {
    if (m <= nd->cap_fwd) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    const int cap = m + m / 8;
    for (int i = 1; i < nn->num_layers; i++) {
        if (nd->a[i]) CUDA_CHECK(cudaFree(nd->a[i]));
        nd->a[i] = dmalloc<float>((size_t)cap * nd->sizes[i]);
        nn->layers[i].d_input = nd->a[i];
    }
    nd->cap_fwd = cap;
}

static void ensure_bwd_capacity(NeuralNetwork* nn, NetDev* nd, int m) {
    if (m <= nd->cap_bwd) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    const int cap = m + m / 8;
    for (int i = 1; i < nn->num_layers; i++) {   // gradient wrt layer-0 input is never formed
        if (nd->gx[i]) CUDA_CHECK(cudaFree(nd->gx[i]));
        nd->gx[i] = dmalloc<float>((size_t)cap * nd->sizes[i]);
        nn->layers[i].d_grad_x = nd->gx[i];
    }
    nd->cap_bwd = cap;
}

// ---- bf16 operand mode helpers ----------------------------------------------------------------------------
static bool bf16_layer(const NetDev* nd, int i, int m) { return tc_bf16_shape_ok(m, nd->sizes[i], nd->sizes[i + 1]); }
// first layer of a low-dimensional env (n < 64): its K is padded to one 64-element k-block so that forward and dW run on the
// tensor cores too (dX of layer 0 is never needed)
constexpr int kPadK = 64;
static bool bf16_layer0_padk(const NetDev* nd, int m) {
    return nd->num_layers >= 3 && nd->sizes[0] < 64 && m >= 128 && nd->sizes[1] >= 64 && (nd->sizes[1] % 8) == 0;
}
// element offset of layer i's W16 inside params_bf16 (16-byte aligned); Wt16 follows at + bf16_total(nd)
static size_t bf16_w_off(const NetDev* nd, int i) {
    size_t off = 0;
    for (int j = 0; j < i; j++) off = ((off + 7) & ~size_t(7)) + (size_t)nd->sizes[j] * nd->sizes[j + 1];
    return (off + 7) & ~size_t(7);
}
static size_t bf16_total(const NetDev* nd) { return (bf16_w_off(nd, nd->num_layers - 1) + 7) & ~size_t(7); }
static void ensure_shadow(std::vector<void*>& v, std::vector<int>& cap, int i, int m, int width, int elem_bytes = 2) {
    if ((int)v.size() <= i) { v.resize(i + 1, nullptr); cap.resize(i + 1, 0); }
    if (cap[i] >= m) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    if (v[i]) CUDA_CHECK(cudaFree(v[i]));
    cap[i] = m + m / 8;
    CUDA_CHECK(cudaMalloc(&v[i], (size_t)cap[i] * width * elem_bytes + 64));
}
// 3xTF32 split mode: layers whose three contractions run on the tensor cores (TMA needs 16-byte pitches and bases)
static bool x3_layer(const NetDev* nd, int i, int m) {
    const int n = nd->sizes[i], l = nd->sizes[i + 1];
    return matmul_precision() == 3 && m >= 128 && n >= 64 && l >= 64 && (n % 4) == 0 && (l % 4) == 0 && (nd->w_off[i] % 4) == 0;
}

void net_forward(NeuralNetwork* nn, const float* input, int m, bool borrow_input) {
    NetDev* nd = net_dev(nn);
    ensure_fwd_capacity(nn, nd, m);
    const int L = nn->num_layers - 1;
    if (borrow_input) {
        nd->a[0] = const_cast<float*>(input);
        nd->a0_borrowed = true;
    } else {   // the reference keeps a private copy of the input (src/neural_network.cu:81)
        if (m > nd->a0_cap) {
            CUDA_CHECK(cudaStreamSynchronize(stream()));
            if (nd->a0_owned) CUDA_CHECK(cudaFree(nd->a0_owned));
            nd->a0_cap = m + m / 8;
            nd->a0_owned = dmalloc<float>((size_t)nd->a0_cap * nd->sizes[0]);
        }
        CUDA_CHECK(cudaMemcpyAsync(nd->a0_owned, input, (size_t)m * nd->sizes[0] * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
        nd->a[0] = nd->a0_owned;
        nd->a0_borrowed = false;
    }
    nn->layers[0].d_input = nd->a[0];
    // TF32 mode: the tensor-core layers read an RNA-rounded shadow of the weights (refreshed per forward,
    // i.e. once per optimiser step; backward re-uses it).  Biases stay fp32 (added in the epilogue).
    const float* wsrc = nd->params;
    if (matmul_precision() == 1) {
        if (!nd->params_tf32) nd->params_tf32 = dmalloc<float>(nd->param_count);
        tc_round_copy(nd->params, nd->params_tf32, nd->param_count);
        wsrc = nd->params_tf32;
    }
    if (matmul_precision() == 3) {
        // 3xTF32 mode: lo companions of the weights (once per forward = once per optimiser step) and of the layer inputs
        if (!nd->params_lo) nd->params_lo = dmalloc<float>(nd->param_count + 64);
        bool split_w = false, lo_ready = false;     // lo_ready: alo[i] already holds the lo companion of a[i]
        for (int i = 0; i < L; i++) {
            const int n = nd->sizes[i], l = nd->sizes[i + 1];
            if (x3_layer(nd, i, m)) {
                if (!split_w) { tc_split_lo(nd->params, nd->params_lo, nd->param_count); split_w = true; }
                ensure_shadow(nd->alo, nd->alo_cap, i, m, n, 4);
                float* alo = static_cast<float*>(nd->alo[i]);
                if (!lo_ready) tc_split_lo(nd->a[i], alo, (size_t)m * n);
                lo_ready = false;
                tc_linear_forward_x3(nd->a[i + 1], nd->a[i], alo, nd->params + nd->w_off[i], nd->params_lo + nd->w_off[i], nd->params + nd->b_off[i],
                                     m, n, l, nd->acts[i]);
            } else {
                // a narrow first layer feeding a split layer writes the lo companion of its output itself
                if (i + 1 < L && n <= 32 && x3_layer(nd, i + 1, m)) {
                    ensure_shadow(nd->alo, nd->alo_cap, i + 1, m, l, 4);
                    if (narrow_first_forward(nd->a[i + 1], static_cast<float*>(nd->alo[i + 1]), nd->a[i], nd->params + nd->w_off[i],
                                             nd->params + nd->b_off[i], m, n, l, nd->acts[i])) { lo_ready = true; continue; }
                }
                linear_forward(nd->a[i + 1], nd->a[i], nd->params + nd->w_off[i], nd->params + nd->b_off[i], m, n, l, nd->acts[i]);
            }
        }
    } else if (matmul_precision() == 2) {
        // BF16 mode: weight copies W16 | Wt16 refreshed per forward (= once per optimiser step); a tensor-core layer reads the
        // bf16 shadow of its input (written by the previous tensor-core layer's epilogue, else converted here) and writes the
        // shadow of its output when the next layer wants it
        const size_t tot16 = bf16_total(nd);
        if (!nd->params_bf16) CUDA_CHECK(cudaMalloc(&nd->params_bf16, 2 * tot16 * 2 + 64));
        char* w16 = static_cast<char*>(nd->params_bf16);
        bool have16 = false;          // a16[i] holds the shadow of a[i]
        for (int i = 0; i < L; i++) {
            const int n = nd->sizes[i], l = nd->sizes[i + 1];
            if (i == 0 && bf16_layer0_padk(nd, m)) {
                if (!nd->w0pad_bf16) CUDA_CHECK(cudaMalloc(&nd->w0pad_bf16, (size_t)l * kPadK * 2 + 64));
                tc_pad_bf16_v(nd->params + nd->w_off[0], nd->w0pad_bf16, (size_t)l, n, kPadK);
                ensure_shadow(nd->a16, nd->a16_cap, 0, m, kPadK);
                tc_pad_bf16_v(nd->a[0], nd->a16[0], (size_t)m, n, kPadK);
                const bool next16 = L > 1 && bf16_layer(nd, 1, m);
                if (next16) ensure_shadow(nd->a16, nd->a16_cap, 1, m, l);
                tc_linear_forward_bf16_v(nd->a[1], next16 ? nd->a16[1] : nullptr, nd->a16[0], nd->w0pad_bf16, nd->params + nd->b_off[0], m, kPadK, l, nd->acts[0]);
                have16 = next16;
            } else if (bf16_layer(nd, i, m)) {
                const size_t wo = bf16_w_off(nd, i);
                tc_weights_bf16_v(nd->params + nd->w_off[i], w16 + 2 * wo, w16 + 2 * (tot16 + wo), l, n);
                if (!have16) { ensure_shadow(nd->a16, nd->a16_cap, i, m, n); tc_to_bf16_v(nd->a[i], nd->a16[i], (size_t)m * n); }
                const bool next16 = i + 1 < L && bf16_layer(nd, i + 1, m);
                if (next16) ensure_shadow(nd->a16, nd->a16_cap, i + 1, m, l);
                tc_linear_forward_bf16_v(nd->a[i + 1], next16 ? nd->a16[i + 1] : nullptr, nd->a16[i], w16 + 2 * wo,
                                         nd->params + nd->b_off[i], m, n, l, nd->acts[i]);
                have16 = next16;
            } else {
                linear_forward(nd->a[i + 1], nd->a[i], nd->params + nd->w_off[i], nd->params + nd->b_off[i], m, n, l, nd->acts[i]);
                have16 = false;
            }
        }
    } else {
        for (int i = 0; i < L; i++)
            linear_forward(nd->a[i + 1], nd->a[i], wsrc + nd->w_off[i], nd->params + nd->b_off[i], m,
                           nd->sizes[i], nd->sizes[i + 1], nd->acts[i]);
    }
    nn->cache_m_forward = m;
    nd->last_m = m;
    nn->d_output = nd->a[L];
}

// Narrow head layer (<= 8 outputs) above a wide layer: dW slabs and dX from ONE pass over the layer input (narrow.cu); db from
// the m x l gradient.  `W` = the weights the forward pass used; gx_lo (3xTF32 mode) receives the lo companion of dX.
// db_below: also emit the db slabs of layer i - 1 (column sums of dX), for callers that would otherwise run a column-sum pass.
static bool head_backward_fused(NetDev* nd, int i, int m, int splits, const float* g, const float* W, float* gx_lo, bool db_below = false) {
    const int n = nd->sizes[i], l = nd->sizes[i + 1];
    if (i == 0 || l > 8 || n < 64) return false;
    if (!narrow_head_backward(nd->partials + nd->w_off[i], nd->slab_stride(), splits, nd->gx[i], gx_lo,
                              db_below ? nd->partials + nd->b_off[i - 1] : nullptr, g, nd->a[i], W, m, n, l, nd->acts[i - 1]))
        return false;
    launch_colsum(nd->partials + nd->b_off[i], nd->slab_stride(), splits, g, m, l);
    return true;
}

void net_backward_partials(NeuralNetwork* nn, const float* grad_out, int m) {
    NetDev* nd = net_dev(nn);
    if (m != nd->last_m) B200_FATAL("backward with m=%d after forward with m=%d", m, nd->last_m);
    ensure_bwd_capacity(nn, nd, m);
    const int L = nn->num_layers - 1;
    int splits = choose_splits(m, nd->param_count);
    if (matmul_precision() >= 1 && m >= 128) {
        // tensor-core dW: one 128x256 tile per CTA, so split-K only until ~2 CTAs per SM exist for the widest
        // layer; more slabs would only add slab traffic (each slab is a full copy of the gradient).
        int tiles = 0;
        for (int i = 0; i < L; i++)
            if (nd->sizes[i] >= 64 && nd->sizes[i + 1] >= 64) tiles = std::max(tiles, div_up(nd->sizes[i], 256) * div_up(nd->sizes[i + 1], 128));
        if (tiles > 0) splits = std::min(splits, std::max(1, div_up(2 * num_sms(), tiles)));
    }
    const size_t need = (size_t)splits * nd->slab_stride();
    if (need > nd->partials_cap) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (nd->partials) CUDA_CHECK(cudaFree(nd->partials));
        nd->partials = dmalloc<float>(need);
        nd->partials_cap = need;
    }
    nd->last_splits = splits;
    // gradient wrt the output, with the output activation's derivative (src/neural_network.cu:199-201)
    const float* g = grad_out;
    if (nd->acts[L - 1] != kActNone) {
        CUDA_CHECK(cudaMemcpyAsync(nd->gx[L], grad_out, (size_t)m * nd->sizes[L] * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
        activation_grad_inplace(nd->a[L], nd->gx[L], (long long)m * nd->sizes[L], nd->acts[L - 1]);
        g = nd->gx[L];
    }
    if (matmul_precision() == 2) {
        char* w16 = static_cast<char*>(nd->params_bf16);
        const size_t tot16 = bf16_total(nd);
        bool have16 = false;          // gx16[i + 1] holds the shadow of g (the gradient wrt a[i + 1])
        bool db16_ready = false;      // the db slabs of layer i were written by the dX epilogue of layer i + 1
        for (int i = L - 1; i >= 0; i--) {
            const int n = nd->sizes[i], l = nd->sizes[i + 1];
            if (head_backward_fused(nd, i, m, splits, g, nd->params + nd->w_off[i], nullptr)) { g = nd->gx[i]; have16 = false; continue; }
            if (i == 0 && w16 && bf16_layer0_padk(nd, m) && nd->w0pad_bf16 && !nd->a16.empty() && nd->a16[0]) {
                // K-padded first layer: dW through the tensor cores into [splits][l][64] staging slabs, columns < n copied out
                if (!have16) { ensure_shadow(nd->gx16, nd->gx16_cap, 1, m, l); tc_to_bf16_v(g, nd->gx16[1], (size_t)m * l); }
                float* tmp = static_cast<float*>(scratch(kScratchSkinny, (size_t)splits * l * kPadK * sizeof(float)));
                tc_linear_backward_weights_bf16_v(tmp, (size_t)l * kPadK, splits, nd->gx16[1], nd->a16[0], m, kPadK, l);
                tc_unpad_slabs(tmp, nd->partials + nd->w_off[0], nd->slab_stride(), splits, l, n, kPadK);
                if (!db16_ready) tc_colsum_bf16_v(nd->partials + nd->b_off[0], nd->slab_stride(), splits, nd->gx16[1], m, l);
                db16_ready = false;
            } else if (w16 && bf16_layer(nd, i, m) && (int)nd->a16.size() > i && nd->a16[i]) {
                if (!have16) { ensure_shadow(nd->gx16, nd->gx16_cap, i + 1, m, l); tc_to_bf16_v(g, nd->gx16[i + 1], (size_t)m * l); }
                tc_linear_backward_weights_bf16_v(nd->partials + nd->w_off[i], nd->slab_stride(), splits, nd->gx16[i + 1], nd->a16[i], m, n, l);
                if (!db16_ready) tc_colsum_bf16_v(nd->partials + nd->b_off[i], nd->slab_stride(), splits, nd->gx16[i + 1], m, l);
                db16_ready = false;
                if (i > 0) {
                    const bool next16 = (bf16_layer(nd, i - 1, m) || (i == 1 && bf16_layer0_padk(nd, m))) && (int)nd->a16.size() > i - 1 && nd->a16[i - 1];
                    // the layer below runs on the tensor path too: its db (fp32 column sums) comes out of this dX epilogue
                    if (next16) { tc_request_colsum(nd->partials + nd->b_off[i - 1], nd->slab_stride(), splits); db16_ready = true; }
                    if (next16) ensure_shadow(nd->gx16, nd->gx16_cap, i, m, n);
                    tc_linear_backward_input_bf16_v(nd->gx[i], next16 ? nd->gx16[i] : nullptr, nd->gx16[i + 1],
                                                    w16 + 2 * (tot16 + bf16_w_off(nd, i)), nd->a[i], m, n, l, nd->acts[i - 1]);
                    g = nd->gx[i];
                    have16 = next16;
                }
            } else {
                linear_backward_params(nd->partials + nd->w_off[i], nd->partials + nd->b_off[i], nd->slab_stride(), splits, g, nd->a[i], m, n, l);
                if (i > 0) {
                    linear_backward_input(nd->gx[i], g, nd->params + nd->w_off[i], nd->a[i], m, n, l, nd->acts[i - 1]);
                    g = nd->gx[i];
                }
                have16 = false;
                db16_ready = false;
            }
        }
    } else if (matmul_precision() == 3) {
        bool lo_ready = false;        // glo[i + 1] already holds the lo companion of g (written by the fused head kernel)
        bool db_ready = false;        // ... and the db slabs of layer i are written too
        for (int i = L - 1; i >= 0; i--) {
            const int n = nd->sizes[i], l = nd->sizes[i + 1];
            float* gWp = nd->partials + nd->w_off[i];
            {
                float* next_lo = nullptr;
                if (i > 0 && x3_layer(nd, i - 1, m)) { ensure_shadow(nd->glo, nd->glo_cap, i, m, n, 4); next_lo = static_cast<float*>(nd->glo[i]); }
                if (head_backward_fused(nd, i, m, splits, g, nd->params + nd->w_off[i], next_lo, next_lo != nullptr)) {
                    g = nd->gx[i]; lo_ready = db_ready = next_lo != nullptr; continue;
                }
            }
            if (x3_layer(nd, i, m) && nd->params_lo && (int)nd->alo.size() > i && nd->alo[i] && ((nd->slab_stride() * 4) % 16) == 0 &&
                ((uintptr_t)gWp & 15) == 0) {
                ensure_shadow(nd->glo, nd->glo_cap, i + 1, m, l, 4);
                float* glo = static_cast<float*>(nd->glo[i + 1]);
                if (!lo_ready) tc_split_lo(g, glo, (size_t)m * l);
                lo_ready = false;
                tc_linear_backward_weights_x3(gWp, nd->slab_stride(), splits, g, glo, nd->a[i], static_cast<const float*>(nd->alo[i]), m, n, l);
                if (!db_ready) launch_colsum(nd->partials + nd->b_off[i], nd->slab_stride(), splits, g, m, l);
                db_ready = false;
                if (i > 0) {
                    // the layer below is a split layer too: its db comes out of this dX epilogue instead of a column-sum pass
                    if (x3_layer(nd, i - 1, m)) { tc_request_colsum(nd->partials + nd->b_off[i - 1], nd->slab_stride(), splits); db_ready = true; }
                    tc_linear_backward_input_x3(nd->gx[i], g, glo, nd->params + nd->w_off[i], nd->params_lo + nd->w_off[i], nd->a[i], m, n, l, nd->acts[i - 1]);
                    g = nd->gx[i];
                }
            } else {
                linear_backward_params(gWp, nd->partials + nd->b_off[i], nd->slab_stride(), splits, g, nd->a[i], m, n, l);
                if (i > 0) {
                    linear_backward_input(nd->gx[i], g, nd->params + nd->w_off[i], nd->a[i], m, n, l, nd->acts[i - 1]);
                    g = nd->gx[i];
                }
                lo_ready = db_ready = false;
            }
        }
    } else {
    bool db_done = false;
    for (int i = L - 1; i >= 0; i--) {
        const int n = nd->sizes[i], l = nd->sizes[i + 1];
        {
            const float* wfwd = (matmul_precision() == 1 && nd->params_tf32) ? nd->params_tf32 : nd->params;
            // TF32 mode: a tensor-core layer below would run a column-sum pass for its db; the head kernel emits it instead
            const bool tc_below = matmul_precision() == 1 && i > 0 && m >= 128 && nd->sizes[i - 1] >= 64 && n >= 64;
            if (head_backward_fused(nd, i, m, splits, g, wfwd + nd->w_off[i], nullptr, tc_below)) { g = nd->gx[i]; db_done = tc_below; continue; }
        }
        linear_backward_params(nd->partials + nd->w_off[i], nd->partials + nd->b_off[i], nd->slab_stride(), splits, g,
                               nd->a[i], m, n, l, db_done);
        db_done = false;
        if (i > 0) {  // dX of layer 0 is unused by every caller (the reference computes it anyway)
            const float* wsrc = (matmul_precision() == 1 && nd->params_tf32) ? nd->params_tf32 : nd->params;
            // TF32 mode, tensor-core layer above a tensor-core layer: the db of the layer below comes out of this dX epilogue
            if (matmul_precision() == 1 && m >= 128 && n >= 64 && l >= 64 && nd->sizes[i - 1] >= 64) {
                tc_request_colsum(nd->partials + nd->b_off[i - 1], nd->slab_stride(), splits);
                db_done = true;
            }
            linear_backward_input(nd->gx[i], g, wsrc + nd->w_off[i], nd->a[i], m, n, l, nd->acts[i - 1]);
            if (tc_colsum_pending()) { tc_request_colsum(nullptr, 0, 0); db_done = false; }      // this dX did not run on the tensor path
            g = nd->gx[i];
        }
    }
    }
    nn->cache_m_backward = m;
}

void net_reduce_grads(NeuralNetwork* nn) {
    NetDev* nd = net_dev(nn);
    reduce_partials(nd->grads, nd->partials, nd->last_splits, nd->slab_stride(), (int)nd->param_count);
}

static NeuralNetwork* alloc_host_net(int num_layers) {
    NeuralNetwork* nn = (NeuralNetwork*)malloc(sizeof(NeuralNetwork));
    nn->num_layers = num_layers;
    nn->layers = (Layer*)calloc(num_layers, sizeof(Layer));   // one extra slot, as the reference
    nn->activation_functions = (char**)malloc((num_layers - 1) * sizeof(char*));
    nn->output = nullptr;
    nn->d_output = nullptr;
    return nn;
}

static void alloc_host_layer(NeuralNetwork* nn, int i, int in, int out) {
    Layer& L = nn->layers[i];
    L.input_size = in;
    L.output_size = out;
    L.weights = (float*)malloc((size_t)in * out * sizeof(float));
    L.biases = (float*)malloc((size_t)out * sizeof(float));
    L.grad_weights = (float*)calloc((size_t)in * out, sizeof(float));
    L.grad_biases = (float*)calloc(out, sizeof(float));
    L.activation_function = build_activation_function(nn->activation_functions[i]);
    L.input = nullptr;
}

}  // namespace b200

using namespace b200;

extern "C" {

// src/neural_network.cu:6-72.  The initialiser consumes glibc rand() in the reference's order
// (per layer: all W then all b) and in its float arithmetic, so a seeded srand() reproduces the
// reference's initial weights bit for bit (host-side integer->float setup, not a compute path).
NeuralNetwork* create_neural_network(int* layer_sizes, char** activation_functions, int num_layers) {
    NeuralNetwork* nn = alloc_host_net(num_layers);
    for (int i = 0; i < num_layers - 1; i++) nn->activation_functions[i] = strdup(activation_functions[i]);
    for (int i = 0; i < num_layers - 1; i++) {
        alloc_host_layer(nn, i, layer_sizes[i], layer_sizes[i + 1]);
        const float gain = i == num_layers - 2 ? 1 : sqrtf(2.0);
        const float std = gain * sqrtf(2.0 / (layer_sizes[i] + layer_sizes[i + 1]));
        for (int j = 0; j < layer_sizes[i] * layer_sizes[i + 1]; j++)
            nn->layers[i].weights[j] = (2 * (float)rand() / RAND_MAX - 1) * sqrtf(3.0) * std;
        for (int j = 0; j < layer_sizes[i + 1]; j++)
            nn->layers[i].biases[j] = (2 * (float)rand() / RAND_MAX - 1) * (1. / sqrtf(layer_sizes[i]));
    }
    nn->layers[num_layers - 1].input_size = layer_sizes[num_layers - 1];
    nn->output_size = nn->layers[num_layers - 2].output_size;
    attach_device(nn);
    nn_write_weights_to_device(nn);
    return nn;
}

void free_neural_network(NeuralNetwork* nn) {
    if (!nn) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    auto it = g_nets.find(nn);
    if (it != g_nets.end()) {
        NetDev* nd = it->second;
        CUDA_CHECK(cudaFree(nd->params));
        CUDA_CHECK(cudaFree(nd->grads));
        for (int i = 1; i < nn->num_layers; i++) {
            if (nd->a[i]) CUDA_CHECK(cudaFree(nd->a[i]));
            if (nd->gx[i]) CUDA_CHECK(cudaFree(nd->gx[i]));
        }
        if (nd->a0_owned) CUDA_CHECK(cudaFree(nd->a0_owned));
        if (nd->partials) CUDA_CHECK(cudaFree(nd->partials));
        if (nd->image) CUDA_CHECK(cudaFree(nd->image));
        if (nd->params_tf32) CUDA_CHECK(cudaFree(nd->params_tf32));
        if (nd->params_bf16) CUDA_CHECK(cudaFree(nd->params_bf16));
        if (nd->w0pad_bf16) CUDA_CHECK(cudaFree(nd->w0pad_bf16));
        if (nd->params_lo) CUDA_CHECK(cudaFree(nd->params_lo));
        for (void* q : nd->alo) if (q) CUDA_CHECK(cudaFree(q));
        for (void* q : nd->glo) if (q) CUDA_CHECK(cudaFree(q));
        for (void* q : nd->a16) if (q) CUDA_CHECK(cudaFree(q));
        for (void* q : nd->gx16) if (q) CUDA_CHECK(cudaFree(q));
        delete nd;
        g_nets.erase(it);
    }
    for (int i = 0; i < nn->num_layers - 1; i++) {
        free(nn->layers[i].weights);
        free(nn->layers[i].biases);
        free(nn->layers[i].grad_weights);
        free(nn->layers[i].grad_biases);
        free(nn->layers[i].input);
        free(nn->layers[i].activation_function);
        free(nn->layers[i].d_activation_function);
        free(nn->activation_functions[i]);
    }
    free(nn->activation_functions);
    free(nn->layers);
    free(nn->output);
    free(nn);
}

void nn_write_weights_to_device(NeuralNetwork* nn) {   // src/neural_network.cu:233-239
    NetDev* nd = net_dev(nn);
    nd->image_dirty = true;
    for (int i = 0; i < nn->num_layers - 1; i++) {
        const Layer& L = nn->layers[i];
        CUDA_CHECK(cudaMemcpyAsync(nd->params + nd->w_off[i], L.weights, (size_t)L.input_size * L.output_size * sizeof(float), cudaMemcpyHostToDevice, stream()));
        CUDA_CHECK(cudaMemcpyAsync(nd->params + nd->b_off[i], L.biases, (size_t)L.output_size * sizeof(float), cudaMemcpyHostToDevice, stream()));
    }
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

void nn_write_weights_to_host(NeuralNetwork* nn) {     // src/neural_network.cu:241-247
    NetDev* nd = net_dev(nn);
    for (int i = 0; i < nn->num_layers - 1; i++) {
        const Layer& L = nn->layers[i];
        CUDA_CHECK(cudaMemcpyAsync(L.weights, nd->params + nd->w_off[i], (size_t)L.input_size * L.output_size * sizeof(float), cudaMemcpyDeviceToHost, stream()));
        CUDA_CHECK(cudaMemcpyAsync(L.biases, nd->params + nd->b_off[i], (size_t)L.output_size * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    }
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

void forward_propagation_cuda(NeuralNetwork* nn, float* input, int m) {   // src/neural_network.cu:74-105
    net_forward(nn, input, m, /*borrow_input=*/false);
}

void backward_propagation_cuda(NeuralNetwork* nn, float* grad_in, int m) { // src/neural_network.cu:121-161
    net_backward_partials(nn, grad_in, m);
    net_reduce_grads(nn);
}

// Host-pointer twins (src/neural_network.cu:163-231): device weights are refreshed from the host
// arrays (the host twins of the reference read layers[i].weights), input is staged, nn->output and
// the host gradient arrays are filled on return.
void forward_propagation(NeuralNetwork* nn, float* input, int m) {
    NetDev* nd = net_dev(nn);
    nn_write_weights_to_device(nn);
    float* d_in = static_cast<float*>(scratch(kScratchStage2, (size_t)m * nd->sizes[0] * sizeof(float)));
    CUDA_CHECK(cudaMemcpyAsync(d_in, input, (size_t)m * nd->sizes[0] * sizeof(float), cudaMemcpyHostToDevice, stream()));
    net_forward(nn, d_in, m, false);
    free(nn->output);
    nn->output = (float*)malloc((size_t)m * nn->output_size * sizeof(float));
    CUDA_CHECK(cudaMemcpyAsync(nn->output, nn->d_output, (size_t)m * nn->output_size * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

void backward_propagation(NeuralNetwork* nn, float* grad_in, int m) {
    NetDev* nd = net_dev(nn);
    float* d_g = static_cast<float*>(scratch(kScratchStage2, (size_t)m * nn->output_size * sizeof(float)));
    CUDA_CHECK(cudaMemcpyAsync(d_g, grad_in, (size_t)m * nn->output_size * sizeof(float), cudaMemcpyHostToDevice, stream()));
    net_backward_partials(nn, d_g, m);
    net_reduce_grads(nn);
    for (int i = 0; i < nn->num_layers - 1; i++) {
        const Layer& L = nn->layers[i];
        CUDA_CHECK(cudaMemcpyAsync(L.grad_weights, nd->grads + nd->w_off[i], (size_t)L.input_size * L.output_size * sizeof(float), cudaMemcpyDeviceToHost, stream()));
        CUDA_CHECK(cudaMemcpyAsync(L.grad_biases, nd->grads + nd->b_off[i], (size_t)L.output_size * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    }
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

// Checkpoint block, byte format of src/neural_network.cu:284-300 (host arrays are written, like the
// reference: callers sync device->host first, src/ppo.cu:536-538).
void save_neural_network(NeuralNetwork* nn, FILE* file) {
    fwrite(&nn->num_layers, sizeof(int), 1, file);
    fwrite(&nn->output_size, sizeof(int), 1, file);
    for (int i = 0; i < nn->num_layers - 1; i++) {
        int length = (int)strlen(nn->activation_functions[i]) + 1;
        fwrite(&length, sizeof(int), 1, file);
        fwrite(nn->activation_functions[i], sizeof(char), length, file);
    }
    for (int i = 0; i < nn->num_layers - 1; i++) {
        const Layer& L = nn->layers[i];
        fwrite(&L.input_size, sizeof(int), 1, file);
        fwrite(&L.output_size, sizeof(int), 1, file);
        fwrite(L.weights, sizeof(float), (size_t)L.input_size * L.output_size, file);
        fwrite(L.biases, sizeof(float), L.output_size, file);
    }
}

static void must_read(void* dst, size_t size, size_t count, FILE* file) {
    if (fread(dst, size, count, file) != count) B200_FATAL("checkpoint truncated");
}

NeuralNetwork* load_neural_network(FILE* file) {       // src/neural_network.cu:303-358
    int num_layers, output_size;
    must_read(&num_layers, sizeof(int), 1, file);
    must_read(&output_size, sizeof(int), 1, file);
    NeuralNetwork* nn = alloc_host_net(num_layers);
    nn->output_size = output_size;
    for (int i = 0; i < num_layers - 1; i++) {
        int length;
        must_read(&length, sizeof(int), 1, file);
        nn->activation_functions[i] = (char*)malloc(length);
        must_read(nn->activation_functions[i], 1, length, file);
    }
    for (int i = 0; i < num_layers - 1; i++) {
        int in, out;
        must_read(&in, sizeof(int), 1, file);
        must_read(&out, sizeof(int), 1, file);
        alloc_host_layer(nn, i, in, out);
        must_read(nn->layers[i].weights, sizeof(float), (size_t)in * out, file);
        must_read(nn->layers[i].biases, sizeof(float), out, file);
    }
    nn->layers[num_layers - 1].input_size = output_size;
    attach_device(nn);
    nn_write_weights_to_device(nn);
    return nn;
}

}  // extern "C"
