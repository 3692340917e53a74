// tc_gemm.cu — tcgen05 (5th-gen tensor core) GEMMs for the wide-MLP configs (SURVEY.md §8 rows a9/a10,
// BASELINE config 4: 3x1024 nets, minibatch 65536).
//
// The reference runs these layers as cublasSgemm FP32 (src/mat_mul.cu:149-208).  Here one warp-specialised
// kernel template covers the three contractions of a dense layer directly on the row-major fp32 arrays
// the rest of the library uses — no transposed copies, no dtype conversion passes:
//
//   forward   y  = act(x . W^T + b)      A = x  [m][in]   K-major     B = W [out][in]  K-major
//   dX        gx = (g . W) * act'(h)     A = g  [m][out]  K-major     B = W [out][in]  MN-major (k = out)
//   dW        gW = g^T . x  (split-K)    A = g  [m][out]  MN-major    B = x [m][in]    MN-major (k = batch)
//
// Operands are fed as TF32 (tcgen05.mma kind::tf32 reads the fp32 words and uses the top 19 bits),
// accumulation is fp32 in TMEM.  Tolerance for this path is stated separately from the fp32 kernels
// (north-star): ~1e-3 norm-wise on layer outputs / gradients, measured in tests/test_gpu_tc.py.
//
// BF16 operand mode (kind::f16, bf16 x bf16 -> fp32 in TMEM; PPO_B200_TF32=2 / ppo_b200_set_matmul_precision(2)): the same
// kernel template with 2-byte operands.  Activations, gradients and weights keep their fp32 arrays (the rest of the library
// and the optimiser read those); every tensor-core layer additionally writes a bf16 shadow of its output from the epilogue
// and reads bf16 shadows of its inputs, so operand traffic halves and the MMA rate doubles.  dX reads a pre-transposed bf16
// copy of the weights (K-major B), dW reads the bf16 shadows MN-major (SWIZZLE_128B, 64-element chunks, 8-row atoms).
//
// Kernel anatomy (Blackwell guide "canonical GEMM"): persistent CTAs (one per SM, clusters of 2), 64 + 32 * 8 threads =
//   warp 0  TMA producer   cp.async.bulk.tensor.2d (128B swizzle) -> 4-stage shared-memory ring,
//                          mbarrier expect_tx / complete_tx; the B tile of a cluster is loaded half-and-half and multicast
//   warp 1  MMA issuer     one elected lane issues tcgen05.mma (M=128, N=256, K=8 tf32 / 16 bf16) from shared-memory
//                          descriptors; tcgen05.commit releases ring slots (in both CTAs) and publishes the accumulator
//   warps 2-9 epilogue     tcgen05.ld 32x32b.x32 TMEM -> registers, fused bias+activation / activation-derivative mask,
//                          128-byte row stores (+ the packed bf16 shadow)
// Two TMEM accumulator buffers (2 x 256 columns = all of TMEM): the epilogue of tile i overlaps the MMAs of tile i + 1.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kTcBM = 128;
constexpr int kTcBKBytes = 128;            // one 128B swizzle row per k-block
constexpr int kTcStagesMax = 4;            // ring depth: 4 for the split-K dW kernels, 3 where the epilogue stages its tiles for TMA stores
constexpr int kTcStgBytes = 8192;          // per epilogue warp: [32 rows][32 fp32] + [32 rows][64 bf16], both 128-byte rows, swizzled
// element-type constants: BF = false -> tf32 operands read from fp32 words, BF = true -> bf16 operands
template <bool BF> struct TcElem {
    static constexpr int kSize = BF ? 2 : 4;
    static constexpr int kBK = kTcBKBytes / kSize;             // elements per k-block: 32 tf32 / 64 bf16
    static constexpr int kUmmaK = 32 / kSize;                  // one MMA consumes 32 bytes of K: 8 tf32 / 16 bf16
    static constexpr int kMnChunk = kTcBKBytes / kSize;        // MN-major: elements per 128-byte chunk
    static constexpr uint32_t kMnSbo = BF ? 1024u : 512u;      // MN-major atom: 8 k-rows (SWIZZLE_128B) / 4 k-rows (.._BASE32B)
    static constexpr uint32_t kMnLayout = BF ? 2u : 1u;
    static constexpr uint32_t kFormat = BF ? 1u : 2u;          // instruction descriptor a/b format: BF16 = 1, TF32 = 2
};
constexpr int kTcEpiWarps = 8;               // two warps per TMEM lane quadrant, each takes half of the columns
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;

enum TcEpilogue { kTcFwd = 0, kTcDx = 1, kTcDw = 2 };

struct TcArgs {
    float* C;
    __nv_bfloat16* C16;     // optional bf16 shadow of C (kTcFwd / kTcDx), same leading dimension
    int M, N, K;            // GEMM extents (output M x N, reduction K)
    int ldc;
    const float* bias;      // kTcFwd
    const float* xin;       // kTcDx: post-activation input of the layer [M][N]
    int act;
    int k_per_split;        // kTcDw: reduction rows per split
    int splits;             // kTcDw: number of split-K slabs
    size_t c_split_stride;
    int tma_store;          // kTcFwd / kTcDx: the epilogue writes C (and C16) with TMA bulk tensor stores from swizzled staging tiles
    float* colsum_part;     // kTcDx, optional: [ceil(M / 128) * 4][N] column sums of C per 32-row quadrant of every tile (the bias gradient
                            // of the layer below, folded in fixed order by colsum_fold_kernel)
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(s_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(s_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool BF>
__device__ __forceinline__ void tc_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (BF)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version 1 <<46 | layout SWIZZLE_128B (2) <<61
// K-major operands use SWIZZLE_128B (16-byte chunks, 8-row atoms).  MN-major TF32 operands must use
// SWIZZLE_128B_BASE32B (layout type 1: 32-byte chunks, 4-row atoms; cutlass sm100_common.inl: "for mn-major
// tf32 operands, SW128_32B is the only available smem layout"), fed by TMA's SWIZZLE_128B_ATOM_32B mode.
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}

// Round-to-nearest fp32 -> tf32 (kept in an fp32 word).  kind::tf32 TRUNCATES its operands; truncation is
// biased (every product shrinks), and the bias compounds through the layers.  Outputs that feed the next
// tensor-core GEMM are therefore stored already RNA-rounded, so the next MMA reads exact TF32 values.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = round_tf32(src[i]);
}

// CL = CTAs per cluster along the M-tile axis (1 or 2).  With CL = 2 the two CTAs of a cluster work on vertically
// adjacent output tiles, i.e. they need the SAME B tile: each loads half of it and TMA-multicasts it into both
// shared memories, which halves the B traffic out of L2 (a 128x256 TF32 tile pulls 1.5 MB of operands through L2
// per 67 MFLOP - the kernel is L2-bandwidth bound, not tensor bound).  A ring slot may be refilled only after BOTH
// CTAs' MMAs released it, so tcgen05.commit arrives on the empty barrier of both CTAs (count = CL).
//
// X3 = "3xTF32" split mode (fp32-accurate contractions on the tensor cores, precision mode 3): every fp32 operand x is the exact
// sum hi + lo with hi = its top 19 bits (what kind::tf32 reads from the raw word) and lo = x - hi, kept in a companion array
// (tc_split_lo).  A stage then holds four tiles [A | B | A_lo | B_lo] and every k-step issues THREE MMAs into the same TMEM
// accumulator: A.B (= hi.hi), A.B_lo (= hi.lo) and A_lo.B (= lo.hi); the dropped lo.lo term is <= 2^-20 relative per product.
// Measured against float64 the result is as close as an fp32 FFMA GEMM (tests/test_gpu_tc.py), so the 1e-5 parity of the
// fp32 path holds while the contraction runs at a third of the TF32 rate instead of the FFMA rate.
template <int BN, bool A_MN, bool B_MN, int EPI, int CL, bool BF, bool X3>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmC16, const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
               const TcArgs p) {
    static_assert(!(X3 && BF), "the split mode is a TF32 mode");
    using E = TcElem<BF>;
    constexpr int kTcStages = X3 ? 2 : (EPI == kTcDw) ? kTcStagesMax : kTcStagesMax - 1;
    constexpr bool kStaged = EPI != kTcDw;                     // epilogue staging tiles for the TMA-store path
    constexpr int kStgPerWarp = X3 ? kTcStgBytes / 2 : kTcStgBytes;      // the split mode has no bf16 shadow: fp32 staging only
    constexpr int kTcBK = E::kBK, kTcUmmaK = E::kUmmaK;
    constexpr uint32_t A_BYTES = kTcBM * kTcBKBytes;           // 16 KB per stage
    constexpr uint32_t B_BYTES = BN * kTcBKBytes;
    constexpr uint32_t PAIR_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t STAGE_BYTES = (X3 ? 2 : 1) * PAIR_BYTES;
    // Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32 or BF16, majors, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (E::kFormat << 7) | (E::kFormat << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);

    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stg_base = smem + kTcStages * STAGE_BYTES;                        // staging tiles of the epilogue warps (1024-byte aligned)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + (kStaged ? kTcEpiWarps * kStgPerWarp : 0));
    uint64_t* empty_bar = full_bar + kTcStages;
    uint64_t* tmem_full_bar = empty_bar + kTcStages;      // [2] accumulator buffer b holds a finished tile
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2] the epilogue warps have drained buffer b
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL > 1 ? tc_cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

    // ---- persistent tile schedule.  Work item w = (split z, M-tile group mg, N tile nt), nt fastest so that the N tiles of
    // one A row block run back to back (L2 reuse of A).  A cluster walks items c, c + #clusters, ...; CTA `crank` of the
    // cluster owns M tile mg * CL + crank of the item.
    const int tiles_n = (p.N + BN - 1) / BN;
    const int groups_m = ((p.M + kTcBM - 1) / kTcBM + CL - 1) / CL;
    const int splits = (EPI == kTcDw) ? p.splits : 1;
    const int n_items = tiles_n * groups_m * splits;
    const int n_clusters = gridDim.x / CL, cluster_id = blockIdx.x / CL;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB) : "memory");
        if (X3) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmAlo) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&tmBlo) : "memory");
        }
        for (int s = 0; s < kTcStages; s++) { tc_mbar_init(&full_bar[s], 1); tc_mbar_init(&empty_bar[s], CL); }
        for (int b = 0; b < 2; b++) { tc_mbar_init(&tmem_full_bar[b], 1); tc_mbar_init(&tmem_empty_bar[b], kTcEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-wide: two accumulator buffers of BN fp32 columns each (2 * 256 = all of TMEM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (CL > 1) tc_cluster_sync();          // the peer's barriers are initialised before anything is multicast into them

    auto item_coords = [&](int w, int& m0, int& n0, int& kbeg, int& kend, int& z) {
        const int nt = w % tiles_n, rest = w / tiles_n;
        const int mg = rest % groups_m;
        z = rest / groups_m;
        m0 = (mg * CL + (int)crank) * kTcBM;
        n0 = nt * BN;
        kbeg = 0; kend = p.K;
        if (EPI == kTcDw) { kbeg = z * p.k_per_split; kend = min(p.K, kbeg + p.k_per_split); }
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;                                     // ring position, runs across tiles
            for (int w = cluster_id; w < n_items; w += n_clusters) {
                int m0, n0, kbeg, kend, z;
                item_coords(w, m0, n0, kbeg, kend, z);
                const int num_kb = max(0, (kend - kbeg + kTcBK - 1) / kTcBK);
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % kTcStages;
                    const uint32_t phase = (it / kTcStages) & 1;
                    tc_mbar_wait(&empty_bar[s], phase ^ 1);
                    tc_mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    const int k0 = kbeg + kb * kTcBK;
                    auto load_pair = [&](const CUtensorMap* mA, const CUtensorMap* mB, uint8_t* sa) {
                        uint8_t* sb = sa + A_BYTES;
                        if (!A_MN) {
                            tc_tma_load_2d(sa, mA, k0, m0, &full_bar[s]);                         // box {32 k, 128 rows}
                        } else {
#pragma unroll
                            for (int i = 0; i < kTcBM / E::kMnChunk; i++)                               // box {one 128-byte mn chunk, BK k-rows}
                                tc_tma_load_2d(sa + i * (kTcBK * kTcBKBytes), mA, m0 + E::kMnChunk * i, k0, &full_bar[s]);
                        }
                        if (CL == 1) {
                            if (!B_MN) {
                                tc_tma_load_2d(sb, mB, k0, n0, &full_bar[s]);
                            } else {
#pragma unroll
                                for (int i = 0; i < BN / E::kMnChunk; i++)
                                    tc_tma_load_2d(sb + i * (kTcBK * kTcBKBytes), mB, n0 + E::kMnChunk * i, k0, &full_bar[s]);
                            }
                        } else {                     // my half of the shared B tile, multicast to both CTAs of the cluster
                            if (!B_MN) {             // box {32 k, BN/CL rows}
                                constexpr int HR = BN / CL;
                                tc_tma_load_2d_mc(sb + crank * (HR * kTcBKBytes), mB, k0, n0 + (int)crank * HR, &full_bar[s], kMask);
                            } else {
                                constexpr int HB = BN / E::kMnChunk / CL;
#pragma unroll
                                for (int i = 0; i < HB; i++) {
                                    const int bi = (int)crank * HB + i;
                                    tc_tma_load_2d_mc(sb + bi * (kTcBK * kTcBKBytes), mB, n0 + E::kMnChunk * bi, k0, &full_bar[s], kMask);
                                }
                            }
                        }
                    };
                    load_pair(&tmA, &tmB, smem + s * STAGE_BYTES);
                    if (X3) load_pair(&tmAlo, &tmBlo, smem + s * STAGE_BYTES + PAIR_BYTES);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread): tile i accumulates into TMEM buffer i & 1 while the epilogue drains the other =====
        if (lane == 0) {
            uint32_t it = 0, tile = 0;
            for (int w = cluster_id; w < n_items; w += n_clusters, tile++) {
                int m0, n0, kbeg, kend, z;
                item_coords(w, m0, n0, kbeg, kend, z);
                const int num_kb = max(0, (kend - kbeg + kTcBK - 1) / kTcBK);
                const uint32_t buf = tile & 1;
                tc_mbar_wait(&tmem_empty_bar[buf], ((tile >> 1) & 1) ^ 1);      // first use of each buffer passes immediately
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + buf * BN;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int s = it % kTcStages;
                    const uint32_t phase = (it / kTcStages) & 1;
                    tc_mbar_wait(&full_bar[s], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = s_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < kTcBK / kTcUmmaK; k++) {
                        // K-major: 8 rows x 128 B swizzle atoms, 1024 B apart; a k-step is 32 B inside the row.
                        // MN-major: 128-byte MN chunks kTcBK*128 B apart (LBO); tf32: 4 k-rows = 512 B atoms (SBO), bf16: 8 k-rows =
                        //           1024 B atoms; a k-step (8 tf32 / 16 bf16 k-rows of 128 B) spans two atoms.
                        constexpr uint32_t kMnStep = (uint32_t)kTcUmmaK * kTcBKBytes;
                        const uint64_t da = A_MN ? tc_make_desc(sa + k * kMnStep, kTcBK * kTcBKBytes, E::kMnSbo, E::kMnLayout) : tc_make_desc(sa + k * 32, 16, 1024, 2);
                        const uint64_t db = B_MN ? tc_make_desc(sb + k * kMnStep, kTcBK * kTcBKBytes, E::kMnSbo, E::kMnLayout) : tc_make_desc(sb + k * 32, 16, 1024, 2);
                        tc_umma<BF>(tmem_acc, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
                        if (X3) {          // the lo tiles sit PAIR_BYTES behind their hi tiles: same descriptors, start address advanced
                            constexpr uint64_t kLoAdv = (uint64_t)(PAIR_BYTES >> 4);
                            tc_umma<BF>(tmem_acc, da, db + kLoAdv, IDESC, 1u);
                            tc_umma<BF>(tmem_acc, da + kLoAdv, db, IDESC, 1u);
                        }
                    }
                    if (CL == 1) tc_umma_commit(&empty_bar[s]);          // frees the ring slot when these MMAs retire
                    else tc_umma_commit_mc(&empty_bar[s], kMask);        // ... in both CTAs (each writes into the other's slot)
                }
                tc_umma_commit(&tmem_full_bar[buf]);      // accumulator of this tile complete
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quadrant = warp % 4; warps 2..5 take columns [0, BN/2), warps 6..9 the rest =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        uint32_t tile = 0;
        for (int w = cluster_id; w < n_items; w += n_clusters, tile++) {
            int m0, n0, kbeg, kend, z;
            item_coords(w, m0, n0, kbeg, kend, z);
            const int num_kb = max(0, (kend - kbeg + kTcBK - 1) / kTcBK);
            const uint32_t buf = tile & 1;
            const int row = m0 + q * 32 + lane;
            tc_mbar_wait(&tmem_full_bar[buf], (tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + buf * BN;
            float* Crow = p.C + (EPI == kTcDw ? (size_t)z * p.c_split_stride : 0) + (size_t)row * p.ldc;
            const bool rows_ok = row < p.M;
            constexpr int CPW = BN / (kTcEpiWarps / 4);          // columns per warp
            const bool use_tma = kStaged && p.tma_store;
            float* stg32 = reinterpret_cast<float*>(stg_base + (warp - 2) * kStgPerWarp);
            uint4* stg16 = reinterpret_cast<uint4*>(stg_base + (warp - 2) * kStgPerWarp + 4096);      // bf16 mode only
#pragma unroll 1
            for (int c0 = half * CPW; c0 < (half + 1) * CPW; c0 += 32) {
                const int nb = n0 + c0;
                // the activation-derivative operand is fetched while the TMEM load is in flight
                float4 hv[8];
                const bool full32 = nb + 32 <= p.N;
                const bool need_h = EPI == kTcDx && p.act != kActNone && rows_ok;
                if (need_h && full32 && (p.N & 3) == 0) {
                    const float4* hrow4 = reinterpret_cast<const float4*>(p.xin + (size_t)row * p.N + nb);
#pragma unroll
                    for (int j = 0; j < 8; j++) hv[j] = __ldg(hrow4 + j);
                }
                uint32_t r[32];
                if (num_kb > 0) {
                    tc_tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = 0u;
                }
                if (c0 + 32 >= (half + 1) * CPW) {            // last TMEM read of this tile: hand the buffer back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s_u32(&tmem_empty_bar[buf])) : "memory");
                }
                float v[32];
                if (rows_ok || use_tma) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
                    if (EPI == kTcFwd) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (nb + j < p.N) {
                                const float y = act_apply(v[j] + __ldg(p.bias + nb + j), p.act);
                                v[j] = (BF || X3) ? y : round_tf32(y);
                            }
                    } else if (EPI == kTcDx) {
                        if (need_h) {
                            if (full32 && (p.N & 3) == 0) {
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    v[4 * j + 0] = act_grad(hv[j].x, v[4 * j + 0], p.act);
                                    v[4 * j + 1] = act_grad(hv[j].y, v[4 * j + 1], p.act);
                                    v[4 * j + 2] = act_grad(hv[j].z, v[4 * j + 2], p.act);
                                    v[4 * j + 3] = act_grad(hv[j].w, v[4 * j + 3], p.act);
                                }
                            } else {
                                const float* hrow = p.xin + (size_t)row * p.N + nb;
#pragma unroll
                                for (int j = 0; j < 32; j++)
                                    if (nb + j < p.N) v[j] = act_grad(hrow[j], v[j], p.act);
                            }
                        }
                        if (!BF && !X3) {
#pragma unroll
                            for (int j = 0; j < 32; j++) v[j] = round_tf32(v[j]);
                        }
                    }
                }
                if (EPI == kTcDx && p.colsum_part != nullptr) {
                    // Column sums of this 32 x 32 chunk over the warp's 32 rows: recursive halving, 31 shuffles; afterwards lane L
                    // holds the sum of column L (fixed order -> deterministic).  Rows past M contribute zero.
                    float w[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) w[j] = rows_ok ? v[j] : 0.f;
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; i++) {
                            const float send = upper ? w[i] : w[i + off];
                            const float keep = upper ? w[i + off] : w[i];
                            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    if (nb + lane < p.N) p.colsum_part[(size_t)((m0 / kTcBM) * 4 + q) * p.N + nb + lane] = w[0];
                }
                if (rows_ok || use_tma) {
                    if (use_tma) {
                        // Stage the 32 x 32 chunk (thread = row) in the 128B-swizzled layout the store tensor map expects (16-byte piece
                        // j of row r sits at piece j ^ (r & 7): conflict-free per quarter-warp), then one lane issues the bulk tensor
                        // store: full 128-byte lines leave through the TMA unit instead of 32 partial lines per st.v4 (the st.v4
                        // epilogue kept the load/store unit as busy as the tensor pipe).  Rows / columns past M / N are clipped by TMA.
                        const int cidx = (c0 - half * CPW) >> 5;
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // previous stores have read the staging tiles
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            *reinterpret_cast<float4*>(stg32 + lane * 32 + 4 * (j ^ (lane & 7))) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        const bool shadow = BF && p.C16 != nullptr;
                        if (shadow) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                uint4 pk;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), t1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), t3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                stg16[lane * 8 + (((cidx & 1) * 4 + j) ^ (lane & 7))] = pk;
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            tc_tma_store_2d(&tmC, stg32, nb, m0 + q * 32);
                            if (shadow && (cidx & 1)) tc_tma_store_2d(&tmC16, stg16, nb - 32, m0 + q * 32);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        continue;
                    }
                    if (BF && EPI != kTcDw && p.C16) {          // bf16 shadow for the next tensor-core GEMM (round-to-nearest-even)
                        __nv_bfloat16* Hrow = p.C16 + (size_t)row * p.ldc + nb;
                        if (full32 && (p.ldc & 7) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 pk;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                *reinterpret_cast<uint4*>(Hrow + j) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (nb + j < p.N) Hrow[j] = __float2bfloat16_rn(v[j]);
                        }
                    }
                    if (full32 && (p.ldc & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(Crow + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (nb + j < p.N) Crow[nb + j] = v[j];
                    }
                }
            }
        }
    }
    if (kStaged && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // my bulk stores have completed
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
    }
    if (CL > 1) tc_cluster_sync();          // nobody leaves while the peer may still multicast / arrive into this CTA
}

// ---- host: tensor maps --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) B200_FATAL("cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Row-major matrix [rows][cols] of fp32 (tf32 operands) or bf16; box = {box_cols (innermost, 128 B), box_rows}, 128B swizzle.
// MN-major tf32 operands need the 32-byte-atom flavour of the swizzle (SWIZZLE_128B_ATOM_32B <-> UMMA SWIZZLE_128B_BASE32B).
static CUtensorMap make_map(const void* base, int rows, int cols, int box_cols, int box_rows, bool mn_major = false, bool bf16 = false) {
    CUtensorMap m;
    const size_t esz = bf16 ? 2 : 4;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * esz};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    const CUresult r = encode_fn()(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) B200_FATAL("cuTensorMapEncodeTiled failed (%d) for %dx%d box %dx%d", (int)r, rows, cols, box_rows, box_cols);
    return m;
}

static bool tc_tma_store_enabled() {      // PPO_B200_TC_TMA_STORE=0 keeps the per-lane st.v4 epilogue (A/B runs)
    static int on = -1;
    if (on < 0) { const char* e = getenv("PPO_B200_TC_TMA_STORE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}
static int tc_cluster() {      // PPO_B200_TC_CLUSTER=1 disables the 2-CTA multicast clusters (A/B runs)
    static int cl = -1;
    if (cl < 0) { const char* e = getenv("PPO_B200_TC_CLUSTER"); cl = (e && e[0] == '1') ? 1 : 2; }
    return cl;
}

// ---- bias gradient of the layer below, from the dX epilogue ------------------------------------------------------------
// db of layer i - 1 = column sums of the array a dX GEMM of layer i writes.  A separate column-sum pass re-reads that whole
// [minibatch][width] array (268 MB at width 1024: as long as a GEMM); instead the caller announces where the db slabs go
// (tc_request_colsum) and the NEXT dX launch reduces every 32 x 32 chunk in its epilogue, then colsum_fold_kernel adds the
// per-quadrant rows in fixed order, `splits` contiguous groups of them into the `splits` slabs.
static struct { float* gb = nullptr; size_t stride = 0; int splits = 0; } g_colsum_req;
void tc_request_colsum(float* gb_part, size_t stride, int splits) { g_colsum_req.gb = gb_part; g_colsum_req.stride = stride; g_colsum_req.splits = splits; }
bool tc_colsum_pending() { return g_colsum_req.gb != nullptr; }      // true when no dX launch has consumed the request

__global__ void __launch_bounds__(256) colsum_fold_kernel(float* __restrict__ gb_part, size_t stride, const float* __restrict__ part, int n_parts, int N,
                                                          int per) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= N) return;
    const int p0 = blockIdx.y * per, p1 = min(n_parts, p0 + per);
    float t = 0.f;
    int q = p0;
    for (; q + 8 <= p1; q += 8) {
        float u[8];
#pragma unroll
        for (int i = 0; i < 8; i++) u[i] = __ldcg(part + (size_t)(q + i) * N + col);
#pragma unroll
        for (int i = 0; i < 8; i++) t += u[i];
    }
    for (; q < p1; q++) t += __ldcg(part + (size_t)q * N + col);
    gb_part[(size_t)blockIdx.y * stride + col] = t;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CL, bool BF, bool X3>
static void launch_tc_cl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& talo, const CUtensorMap& tblo, const TcArgs& a_in,
                         dim3 tiles /* (N tiles, M tiles, splits) */) {
    constexpr int kStages = X3 ? 2 : (EPI == kTcDw) ? kTcStagesMax : kTcStagesMax - 1;
    constexpr bool kStaged = EPI != kTcDw;
    const size_t smem = (size_t)kStages * (X3 ? 2 : 1) * (kTcBM + BN) * kTcBKBytes + (kStaged ? kTcEpiWarps * (X3 ? kTcStgBytes / 2 : kTcStgBytes) : 0) +
                        1024 /*align*/ + 256 /*barriers*/;
    // output tensor maps of the TMA-store epilogue (forward / dX): fp32 boxes {32 columns, 32 rows}, bf16 shadow boxes {64, 32}
    TcArgs a = a_in;
    CUtensorMap tc = ta, tc16 = ta;        // placeholders when the store path is off
    a.tma_store = 0;
    a.colsum_part = nullptr;
    float* cs_gb = nullptr;
    size_t cs_stride = 0;
    int cs_splits = 0, cs_parts = 0;
    if (EPI == kTcDx && g_colsum_req.gb) {
        cs_gb = g_colsum_req.gb; cs_stride = g_colsum_req.stride; cs_splits = g_colsum_req.splits;
        cs_parts = div_up(a.M, kTcBM) * 4;
        a.colsum_part = static_cast<float*>(scratch(kScratchColsum, (size_t)cs_parts * a.N * sizeof(float)));
    }
    if (EPI == kTcDx) g_colsum_req.gb = nullptr;
    if (kStaged && tc_tma_store_enabled() && (a.ldc % 4) == 0 && a.ldc == a.N && ((uintptr_t)a.C & 15) == 0 &&
        (!a.C16 || ((a.ldc % 8) == 0 && ((uintptr_t)a.C16 & 15) == 0))) {
        a.tma_store = 1;
        tc = make_map(a.C, a.M, a.N, 32, 32, false, false);
        if (a.C16) tc16 = make_map(a.C16, a.M, a.N, 64, 32, false, true);
    }
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<BN, A_MN, B_MN, EPI, CL, BF, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // persistent: one CTA per SM (190 KB of shared memory each), clusters of CL CTAs walk the work items
    const long long items = (long long)tiles.x * div_up(tiles.y, CL) * tiles.z;
    const int clusters = (int)std::min<long long>(items, num_sms() / CL);
    const dim3 grid(clusters * CL, 1, 1);
    if (CL == 1) {
        B200_LAUNCH((tc_gemm_kernel<BN, A_MN, B_MN, EPI, CL, BF, X3>), grid, kTcThreads, smem, ta, tb, tc, tc16, talo, tblo, a);
        if (cs_gb) B200_LAUNCH(colsum_fold_kernel, dim3(div_up(a.N, 256), cs_splits, 1), 256, 0, cs_gb, cs_stride, a.colsum_part, cs_parts, a.N,
                               div_up(cs_parts, cs_splits));
        return;
    }
    const char* label = BF ? "(tc_gemm_kernel<bf16>)" : X3 ? "(tc_gemm_kernel<3xtf32>)" : "(tc_gemm_kernel<BN, A_MN, B_MN, EPI, CL>)";
    if (g_profiling) profile_mark(label, true);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, A_MN, B_MN, EPI, CL, BF, X3>, ta, tb, tc, tc16, talo, tblo, a));
    ++g_launches;
    if (g_profiling) profile_mark(label, false);
    if (cs_gb) B200_LAUNCH(colsum_fold_kernel, dim3(div_up(a.N, 256), cs_splits, 1), 256, 0, cs_gb, cs_stride, a.colsum_part, cs_parts, a.N,
                           div_up(cs_parts, cs_splits));
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool BF>
static void launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& a, dim3 grid) {
    if (tc_cluster() == 2) launch_tc_cl<BN, A_MN, B_MN, EPI, 2, BF, false>(ta, tb, ta, tb, a, grid);
    else launch_tc_cl<BN, A_MN, B_MN, EPI, 1, BF, false>(ta, tb, ta, tb, a, grid);
}
template <int BN, bool A_MN, bool B_MN, int EPI>
static void launch_tc_x3(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& talo, const CUtensorMap& tblo, const TcArgs& a, dim3 grid) {
    if (tc_cluster() == 2) launch_tc_cl<BN, A_MN, B_MN, EPI, 2, false, true>(ta, tb, talo, tblo, a, grid);
    else launch_tc_cl<BN, A_MN, B_MN, EPI, 1, false, true>(ta, tb, talo, tblo, a, grid);
}

// Shapes the tensor path accepts: TMA needs 16-byte row pitches and aligned bases.
bool tc_shape_ok(const void* a, const void* b, int lda_cols, int ldb_cols) {
    return ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && (lda_cols % 4) == 0 && (ldb_cols % 4) == 0;
}

void tc_round_copy(const float* src, float* dst, size_t n) {
    if (n == 0) return;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)num_sms() * 8);
    B200_LAUNCH(round_tf32_kernel, blocks, 256, 0, src, dst, n);
}

constexpr int kTcBN = 256;
constexpr int kTcBK32 = TcElem<false>::kBK, kTcBK16 = TcElem<true>::kBK;

void tc_linear_forward(float* y, const float* x, const float* W, const float* b, int m, int n, int l, int act) {
    TcArgs a{};
    a.C = y; a.M = m; a.N = l; a.K = n; a.ldc = l; a.bias = b; a.act = act;
    const CUtensorMap ta = make_map(x, m, n, kTcBK32, kTcBM);        // A K-major
    const CUtensorMap tb = make_map(W, l, n, kTcBK32, kTcBN / tc_cluster());   // B K-major (half tile per CTA under multicast)
    launch_tc<kTcBN, false, false, kTcFwd, false>(ta, tb, a, dim3(div_up(l, kTcBN), div_up(m, kTcBM), 1));
}

void tc_linear_backward_input(float* gx, const float* g, const float* W, const float* xin, int m, int n, int l, int act_prev) {
    TcArgs a{};
    a.C = gx; a.M = m; a.N = n; a.K = l; a.ldc = n; a.xin = xin; a.act = act_prev;
    const CUtensorMap ta = make_map(g, m, l, kTcBK32, kTcBM);        // A K-major (k = out)
    const CUtensorMap tb = make_map(W, l, n, 32, kTcBK32, true);     // B MN-major: W[k=out][n=in], box {32 n, 32 k}
    launch_tc<kTcBN, false, true, kTcDx, false>(ta, tb, a, dim3(div_up(n, kTcBN), div_up(m, kTcBM), 1));
}

void tc_linear_backward_weights(float* gW_part, size_t stride, int splits, const float* g, const float* x, int m, int n, int l) {
    TcArgs a{};
    int rows = div_up(m, splits);
    rows = div_up(rows, kTcBK32) * kTcBK32;
    a.C = gW_part; a.M = l; a.N = n; a.K = m; a.ldc = n; a.k_per_split = rows; a.c_split_stride = stride; a.splits = splits;
    const CUtensorMap ta = make_map(g, m, l, 32, kTcBK32, true);     // A MN-major: g[k=batch][m'=out]
    const CUtensorMap tb = make_map(x, m, n, 32, kTcBK32, true);     // B MN-major: x[k=batch][n=in]
    launch_tc<kTcBN, true, true, kTcDw, false>(ta, tb, a, dim3(div_up(n, kTcBN), div_up(l, kTcBM), splits));
}

// ---- 3xTF32 split mode (precision 3): fp32-accurate, operands come with their lo = x - top19bits(x) companions --------
__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ src, float* __restrict__ lo, size_t n4, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = ld_stream4(src + 4 * i);
        float4 r;
        r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        *reinterpret_cast<float4*>(lo + 4 * i) = r;
    }
    if (blockIdx.x == 0)
        for (size_t i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) lo[i] = src[i] - __uint_as_float(__float_as_uint(src[i]) & 0xFFFFE000u);
}
// lo[i] = src[i] - (src[i] with the 13 low mantissa bits cleared); exact in fp32.  Both pointers 16-byte aligned.
void tc_split_lo(const float* src, float* lo, size_t n) {
    if (n == 0) return;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((n / 4 + 255) / 256, (size_t)num_sms() * 8));
    B200_LAUNCH(split_lo_kernel, blocks, 256, 0, src, lo, n / 4, n);
}

void tc_linear_forward_x3(float* y, const float* x, const float* xlo, const float* W, const float* Wlo, const float* b, int m, int n, int l, int act) {
    TcArgs a{};
    a.C = y; a.M = m; a.N = l; a.K = n; a.ldc = l; a.bias = b; a.act = act;
    const int hb = kTcBN / tc_cluster();
    launch_tc_x3<kTcBN, false, false, kTcFwd>(make_map(x, m, n, kTcBK32, kTcBM), make_map(W, l, n, kTcBK32, hb), make_map(xlo, m, n, kTcBK32, kTcBM),
                                              make_map(Wlo, l, n, kTcBK32, hb), a, dim3(div_up(l, kTcBN), div_up(m, kTcBM), 1));
}

void tc_linear_backward_input_x3(float* gx, const float* g, const float* glo, const float* W, const float* Wlo, const float* xin, int m, int n,
                                 int l, int act_prev) {
    TcArgs a{};
    a.C = gx; a.M = m; a.N = n; a.K = l; a.ldc = n; a.xin = xin; a.act = act_prev;
    launch_tc_x3<kTcBN, false, true, kTcDx>(make_map(g, m, l, kTcBK32, kTcBM), make_map(W, l, n, 32, kTcBK32, true), make_map(glo, m, l, kTcBK32, kTcBM),
                                            make_map(Wlo, l, n, 32, kTcBK32, true), a, dim3(div_up(n, kTcBN), div_up(m, kTcBM), 1));
}

void tc_linear_backward_weights_x3(float* gW_part, size_t stride, int splits, const float* g, const float* glo, const float* x, const float* xlo,
                                   int m, int n, int l) {
    TcArgs a{};
    int rows = div_up(m, splits);
    rows = div_up(rows, kTcBK32) * kTcBK32;
    a.C = gW_part; a.M = l; a.N = n; a.K = m; a.ldc = n; a.k_per_split = rows; a.c_split_stride = stride; a.splits = splits;
    launch_tc_x3<kTcBN, true, true, kTcDw>(make_map(g, m, l, 32, kTcBK32, true), make_map(x, m, n, 32, kTcBK32, true), make_map(glo, m, l, 32, kTcBK32, true),
                                           make_map(xlo, m, n, 32, kTcBK32, true), a, dim3(div_up(n, kTcBN), div_up(l, kTcBM), splits));
}

// ---- bf16 operand mode ---------------------------------------------------------------------------------
// bf16 arrays need 16-byte row pitches: column counts that are multiples of 8
bool tc_bf16_shape_ok(int m, int n, int l) { return m >= 128 && n >= 64 && l >= 64 && (n % 8) == 0 && (l % 8) == 0; }

// y = act(x W^T + b): x16 [m][n], W16 [l][n] (both K-major); writes y (fp32) and, when y16 != null, its bf16 shadow
void tc_linear_forward_bf16(float* y, __nv_bfloat16* y16, const __nv_bfloat16* x16, const __nv_bfloat16* W16, const float* b, int m, int n,
                            int l, int act) {
    TcArgs a{};
    a.C = y; a.C16 = y16; a.M = m; a.N = l; a.K = n; a.ldc = l; a.bias = b; a.act = act;
    const CUtensorMap ta = make_map(x16, m, n, kTcBK16, kTcBM, false, true);
    const CUtensorMap tb = make_map(W16, l, n, kTcBK16, kTcBN / tc_cluster(), false, true);
    launch_tc<kTcBN, false, false, kTcFwd, true>(ta, tb, a, dim3(div_up(l, kTcBN), div_up(m, kTcBM), 1));
}

// gx = (g W) act'(xin): g16 [m][l] K-major, Wt16 [n][l] = the TRANSPOSED weights (K-major B); xin fp32
void tc_linear_backward_input_bf16(float* gx, __nv_bfloat16* gx16, const __nv_bfloat16* g16, const __nv_bfloat16* Wt16, const float* xin,
                                   int m, int n, int l, int act_prev) {
    TcArgs a{};
    a.C = gx; a.C16 = gx16; a.M = m; a.N = n; a.K = l; a.ldc = n; a.xin = xin; a.act = act_prev;
    const CUtensorMap ta = make_map(g16, m, l, kTcBK16, kTcBM, false, true);
    const CUtensorMap tb = make_map(Wt16, n, l, kTcBK16, kTcBN / tc_cluster(), false, true);
    launch_tc<kTcBN, false, false, kTcDx, true>(ta, tb, a, dim3(div_up(n, kTcBN), div_up(m, kTcBM), 1));
}

// gW = g^T x (split-K slabs, fp32): g16 [m][l] and x16 [m][n] read MN-major (k = batch)
void tc_linear_backward_weights_bf16(float* gW_part, size_t stride, int splits, const __nv_bfloat16* g16, const __nv_bfloat16* x16, int m,
                                     int n, int l) {
    TcArgs a{};
    int rows = div_up(m, splits);
    rows = div_up(rows, kTcBK16) * kTcBK16;
    a.C = gW_part; a.M = l; a.N = n; a.K = m; a.ldc = n; a.k_per_split = rows; a.c_split_stride = stride; a.splits = splits;
    const CUtensorMap ta = make_map(g16, m, l, 64, kTcBK16, true, true);     // box {64 m' (128 B), 64 k-rows}
    const CUtensorMap tb = make_map(x16, m, n, 64, kTcBK16, true, true);
    launch_tc<kTcBN, true, true, kTcDw, true>(ta, tb, a, dim3(div_up(n, kTcBN), div_up(l, kTcBM), splits));
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread; n must be a multiple of 8 and both pointers 16-byte aligned
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n8) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = ld_stream4(src + 8 * i), b = ld_stream4(src + 8 * i + 4);
        __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x, a.y), t1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(b.x, b.y), t3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(dst + 8 * i) = pk;
    }
}
void tc_to_bf16(const float* src, __nv_bfloat16* dst, size_t n) {
    if (n == 0) return;
    if (n % 8) B200_FATAL("tc_to_bf16: %zu elements (must be a multiple of 8)", n);
    const int blocks = (int)std::min<size_t>((n / 8 + 255) / 256, (size_t)num_sms() * 8);
    B200_LAUNCH(to_bf16_kernel, blocks, 256, 0, src, dst, n / 8);
}

// W [l][n] fp32 -> W16 [l][n] and Wt16 [n][l] (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) weights_bf16_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ W16,
                                                           __nv_bfloat16* __restrict__ Wt16, int l, int n) {
    __shared__ float tile[32][33];
    const int j0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, k = k0 + tx;
        float v = 0.f;
        if (j < l && k < n) { v = W[(size_t)j * n + k]; W16[(size_t)j * n + k] = __float2bfloat16_rn(v); }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, j = j0 + tx;
        if (k < n && j < l) Wt16[(size_t)k * l + j] = __float2bfloat16_rn(tile[tx][r]);
    }
}
void tc_weights_bf16(const float* W, __nv_bfloat16* W16, __nv_bfloat16* Wt16, int l, int n) {
    dim3 grid(div_up(n, 32), div_up(l, 32), 1);
    B200_LAUNCH(weights_bf16_kernel, grid, 256, 0, W, W16, Wt16, l, n);
}

// db slabs from the bf16 shadow of g: gb_part[s][col] = sum over the rows of split s of g16[row][col].  Thread = (8-column group,
// row lane): 4 groups x 64 row lanes per CTA (32 columns), eight independent 128-bit loads in flight per thread (the grid is only
// (l / 32) x splits CTAs, so bytes in flight per CTA set the bandwidth); fixed-order combine of the row lanes.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(float* __restrict__ gb_part, size_t stride, const __nv_bfloat16* __restrict__ g, int m, int l, int rows_per_split) {
    __shared__ float red[64][4][9];
    const int cx = threadIdx.x & 3, ry = threadIdx.x >> 2;
    const int col = blockIdx.x * 32 + 8 * cx;
    const int r0 = blockIdx.y * rows_per_split, r1 = min(m, r0 + rows_per_split);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = 0.f;
    auto add = [&](const uint4& t) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; j++) { s[2 * j] += __uint_as_float(w[j] << 16); s[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u); }
    };
    if (col < l) {            // l is a multiple of 8 (bf16 layers)
        int r = r0 + ry;
        for (; r + 7 * 64 < r1; r += 8 * 64) {
            uint4 t[8];
#pragma unroll
            for (int u = 0; u < 8; u++) t[u] = __ldg(reinterpret_cast<const uint4*>(g + (size_t)(r + 64 * u) * l + col));
#pragma unroll
            for (int u = 0; u < 8; u++) add(t[u]);
        }
        for (; r < r1; r += 64) add(__ldg(reinterpret_cast<const uint4*>(g + (size_t)r * l + col)));
    }
#pragma unroll
    for (int j = 0; j < 8; j++) red[ry][cx][j] = s[j];
    __syncthreads();
    if (threadIdx.x < 32 && blockIdx.x * 32 + (int)threadIdx.x < l) {      // thread = column: folds the 64 row lanes in order
        const int c = threadIdx.x >> 3, j = threadIdx.x & 7;
        float t = red[0][c][j];
#pragma unroll 8
        for (int k = 1; k < 64; k++) t += red[k][c][j];
        gb_part[(size_t)blockIdx.y * stride + blockIdx.x * 32 + threadIdx.x] = t;
    }
}
void tc_colsum_bf16_v(float* gb_part, size_t stride, int splits, const void* g16, int m, int l) {
    int rows = div_up(m, splits);
    rows = div_up(rows, 32) * 32;
    dim3 grid(div_up(l, 32), splits, 1);
    B200_LAUNCH(colsum_bf16_kernel, grid, 256, 0, gb_part, stride, static_cast<const __nv_bfloat16*>(g16), m, l, rows);
}

// [rows][n] fp32 -> [rows][npad] bf16, zero padded columns (first layer of a low-dimensional env: K padded to one 64-wide k-block)
__global__ void __launch_bounds__(256) pad_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t rows, int n, int npad) {
    const size_t total = rows * (size_t)npad, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const size_t r = e / npad;
        const int k = (int)(e - r * npad);
        dst[e] = __float2bfloat16_rn(k < n ? src[r * n + k] : 0.f);
    }
}
void tc_pad_bf16_v(const float* src, void* dst, size_t rows, int n, int npad) {
    const int blocks = (int)std::min<size_t>((rows * npad + 255) / 256, (size_t)num_sms() * 8);
    B200_LAUNCH(pad_bf16_kernel, blocks, 256, 0, src, static_cast<__nv_bfloat16*>(dst), rows, n, npad);
}
// slabs [splits][l][npad] (tensor-core dW of a K-padded layer) -> gW slabs [l][n] at gW_part + s * stride
__global__ void __launch_bounds__(256) unpad_slabs_kernel(const float* __restrict__ src, float* __restrict__ gW_part, size_t stride, int l, int n, int npad) {
    const int total = l * n;
    const float* s = src + (size_t)blockIdx.y * l * npad;
    float* d = gW_part + (size_t)blockIdx.y * stride;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int j = e / n, k = e - j * n;
        d[e] = s[(size_t)j * npad + k];
    }
}
void tc_unpad_slabs(const float* src, float* gW_part, size_t stride, int splits, int l, int n, int npad) {
    dim3 grid(std::max(1, std::min(64, div_up(l * n, 256))), splits, 1);
    B200_LAUNCH(unpad_slabs_kernel, grid, 256, 0, src, gW_part, stride, l, n, npad);
}

void tc_linear_forward_bf16_v(float* y, void* y16, const void* x16, const void* W16, const float* b, int m, int n, int l, int act) {
    tc_linear_forward_bf16(y, static_cast<__nv_bfloat16*>(y16), static_cast<const __nv_bfloat16*>(x16), static_cast<const __nv_bfloat16*>(W16), b, m, n, l, act);
}
void tc_linear_backward_input_bf16_v(float* gx, void* gx16, const void* g16, const void* Wt16, const float* xin, int m, int n, int l, int act_prev) {
    tc_linear_backward_input_bf16(gx, static_cast<__nv_bfloat16*>(gx16), static_cast<const __nv_bfloat16*>(g16), static_cast<const __nv_bfloat16*>(Wt16), xin, m, n, l, act_prev);
}
void tc_linear_backward_weights_bf16_v(float* gW_part, size_t stride, int splits, const void* g16, const void* x16, int m, int n, int l) {
    tc_linear_backward_weights_bf16(gW_part, stride, splits, static_cast<const __nv_bfloat16*>(g16), static_cast<const __nv_bfloat16*>(x16), m, n, l);
}
void tc_to_bf16_v(const float* src, void* dst, size_t n) { tc_to_bf16(src, static_cast<__nv_bfloat16*>(dst), n); }
void tc_weights_bf16_v(const float* W, void* W16, void* Wt16, int l, int n) {
    tc_weights_bf16(W, static_cast<__nv_bfloat16*>(W16), static_cast<__nv_bfloat16*>(Wt16), l, n);
}

}  // namespace b200

using namespace b200;

// Test / bench entry point: mode 0 forward, 1 backward-input, 2 backward-weights (one slab per split).
extern "C" void ppo_b200_tc_linear(int mode, float* out, const float* a, const float* b, const float* aux, int m, int n, int l,
                                   int act, int splits) {
    if (mode == 0) tc_linear_forward(out, a, b, aux, m, n, l, act);
    else if (mode == 1) tc_linear_backward_input(out, a, b, aux, m, n, l, act);
    else tc_linear_backward_weights(out, (size_t)n * l, splits, a, b, m, n, l);
}

// Same contractions in the 3xTF32 split mode: the lo companions are formed in scratch first.
extern "C" void ppo_b200_tc_linear_x3(int mode, float* out, const float* a, const float* b, const float* aux, int m, int n, int l, int act,
                                      int splits) {
    const size_t na = (size_t)m * (mode == 0 ? n : l), nb = mode == 2 ? (size_t)m * n : (size_t)l * n;
    const size_t na_al = (na + 63) & ~size_t(63);
    float* lo = static_cast<float*>(scratch(kScratchStage3, (na_al + nb) * sizeof(float) + 256));
    tc_split_lo(a, lo, na);
    tc_split_lo(b, lo + na_al, nb);
    if (mode == 0) tc_linear_forward_x3(out, a, lo, b, lo + na_al, aux, m, n, l, act);
    else if (mode == 1) tc_linear_backward_input_x3(out, a, lo, b, lo + na_al, aux, m, n, l, act);
    else tc_linear_backward_weights_x3(out, (size_t)n * l, splits, a, lo, b, lo + na_al, m, n, l);
}

// Same contractions with bf16 operands: the fp32 inputs are converted into scratch bf16 arrays first (the conversion is
// outside what a caller should time: the training path keeps bf16 shadows resident), `out16` (may be null) receives the
// bf16 shadow of the output.  mode 0: a = x [m][n], b = W [l][n], aux = bias; 1: a = g [m][l], b = W [l][n], aux = xin [m][n];
// 2: a = g [m][l], b = x [m][n].
extern "C" void ppo_b200_tc_linear_bf16(int mode, float* out, void* out16, const float* a, const float* b, const float* aux, int m,
                                        int n, int l, int act, int splits, int convert_only, int skip_convert) {
    const size_t na = (size_t)m * (mode == 0 ? n : l), nw = (size_t)l * n, nb = mode == 2 ? (size_t)m * n : nw;
    char* base = static_cast<char*>(scratch(kScratchStage3, (na + 2 * nb) * 2 + 1024));
    __nv_bfloat16* a16 = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* b16 = reinterpret_cast<__nv_bfloat16*>(base + ((na * 2 + 255) & ~size_t(255)));
    __nv_bfloat16* bt16 = reinterpret_cast<__nv_bfloat16*>(base + ((na * 2 + 255) & ~size_t(255)) + ((nb * 2 + 255) & ~size_t(255)));
    if (!skip_convert) {
        tc_to_bf16(a, a16, na);
        if (mode == 2) tc_to_bf16(b, b16, nb);
        else tc_weights_bf16(b, b16, bt16, l, n);
    }
    if (convert_only) return;
    if (mode == 0) tc_linear_forward_bf16(out, static_cast<__nv_bfloat16*>(out16), a16, b16, aux, m, n, l, act);
    else if (mode == 1) tc_linear_backward_input_bf16(out, static_cast<__nv_bfloat16*>(out16), a16, bt16, aux, m, n, l, act);
    else tc_linear_backward_weights_bf16(out, (size_t)n * l, splits, a16, b16, m, n, l);
}
