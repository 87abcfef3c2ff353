// C = act(beta * C + A.B + bias) on the 5th-generation tensor cores (tcgen05, sm_100a) for the UPDATE path:
// every operand layout the explicit backward passes need, few-tile / long-K shapes split over K.
//
// Reference contractions (paths under simulation-mode/): nn.Linear / nn.GRU forward at update batch sizes
// (problem-04-sac-gru/src/networks.py:82-110, 209-237: M = 256, K ~ 3000, N = 384), their input gradients
// dx = dy W and weight gradients dW = dy^T x (autograd of sac_agent.py:197-231, qmix_agent.py:275-285).
// In terms of C[m][n] = sum_k A(m, k) B(k, n):
//   forward   A = x  [M][K] (k contiguous)          B(k, n) = W[n][k]  (k contiguous)
//   dx        A = dy [M][N'] (k contiguous)         B(k, n) = W[k][n]  (n contiguous: "transposed")
//   dW        A(m, k) = dy[k][m] (m contiguous)     B(k, n) = x[k][n]  (n contiguous)
// tcgen05.mma.kind::tf32 wants both operands K-major in shared memory.  TMA brings every tile in whatever direction
// is contiguous in global memory; the eight worker warps -- which touch every element anyway to split it into a TF32
// "hi" part and a TF32 "lo" remainder (3xTF32: x.w ~= xl.wh + xh.wl + xh.wh, 1e-5 parity with the fp32 reference) --
// write hi and lo K-major into the 128-byte-swizzled operand ring, transposing on the way when needed.
//
// Like mlb_linear_tc.cu the accumulator in tensor memory holds two k-blocks at a time (the tensor core rounds its fp32
// accumulator toward zero on every instruction) and is drained into per-thread registers with round-to-nearest adds.
//
// Split-K (reductions of 4 k-blocks and more on few tiles): grid.z CTAs share one output tile's k-blocks; raw partial
// tiles go to a workspace ([split][M padded to 128][N]) and gemm_tc_reduce_kernel adds them in split order
// (deterministic) with beta / bias / activation, spread over as many SMs as the output has 256-element blocks.  (Tried:
// the last CTA of a tile, found by a ticket counter, adding the partial tiles itself -- one launch less, but its 256
// threads need 16 rounds of L2 latency for a 128 x 128 tile: 1.82 ms per SAC update instead of 1.14.)  Unsplit tiles
// apply beta * C (dW accumulation), bias and the activation in their own epilogue.
//
// One CTA = one 128 x NT tile (NT = 64 or 128) over a range of k-blocks, 10 warps:
//   warp 0      TMA producer (raw ring, 2 stages)
//   warp 1      tensor memory allocation, tcgen05.mma issue (one elected lane), tcgen05.commit
//   warps 2..9  raw tile -> hi / lo operand tiles (2 stages), TMEM drain, epilogue (shared memory -> coalesced stores)
#include <algorithm>

#include "../../include/marllb_b200.h"
#include "../../include/marllb_b200_policy.h"
#include "mlb_tc_common.cuh"

extern "C" float* mlb_workspace_get(int kind, size_t need, void* stream);   // mlb_policy.cu

namespace {
using namespace mlb_tc;

constexpr int RAW_STAGES = 2;
constexpr int OP_STAGES = 2;
constexpr int DRAIN_KB = 2;
constexpr int WORKER_WARPS = 8;
constexpr int WORK_THREADS = 32 * WORKER_WARPS;    // 256
constexpr int THREADS = 64 + WORK_THREADS;

struct GtParams {
    float* C;              // final output (splits == 1) ...
    float* ws;             // ... or the workspace of raw partial tiles [split][m_pad][N] (splits > 1)
    const float* bias;     // only read when the tile is final
    int64_t ldc;
    float beta;
    int M, N, K, act;
    int kb_per_split;      // k-blocks per CTA along grid.z
    int m_pad;             // rows per split in the workspace (splits > 1), else 0
};

__device__ __forceinline__ float gt_act(float v, int act) {
    if (act == MLB_ACT_RELU) return fmaxf(v, 0.f);
    if (act == MLB_ACT_ABS) return fabsf(v);
    return v;
}

__device__ __forceinline__ uint32_t tf32_hi(uint32_t v) { return v & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(uint32_t v) {
    return __float_as_uint(__uint_as_float(v) - __uint_as_float(v & 0xffffe000u)) & 0xffffe000u;
}

// raw tile -> (hi, lo) K-major 128-byte-swizzled tiles of ROWS rows x 32 contraction elements.
//   TRANSPOSED = false: raw is already K-major swizzled (TMA SWIZZLE_128B): element-wise at the same address
//   TRANSPOSED = true : raw is [32 contraction][ROWS] floats, unswizzled; unit (row r, 16-byte chunk q) gathers the
//                       four contraction elements 4q..4q+3 of row r (lanes = consecutive rows: conflict-free reads)
template <int ROWS, bool TRANSPOSED>
__device__ __forceinline__ void split_tile(const unsigned char* raw, unsigned char* hi, unsigned char* lo, int t) {
#pragma unroll
    for (int i = 0; i < (ROWS * TILE_K / 4) / WORK_THREADS; i++) {
        const int u = t + i * WORK_THREADS;
        uint4 v;
        uint32_t dst;
        if (TRANSPOSED) {
            const int r = u % ROWS, q = u / ROWS;
            const float* src = reinterpret_cast<const float*>(raw) + (4 * q) * ROWS + r;
            v.x = __float_as_uint(src[0]);
            v.y = __float_as_uint(src[ROWS]);
            v.z = __float_as_uint(src[2 * ROWS]);
            v.w = __float_as_uint(src[3 * ROWS]);
            dst = (uint32_t)(r * 128 + ((q ^ (r & 7)) << 4));
        } else {
            v = reinterpret_cast<const uint4*>(raw)[u];
            dst = (uint32_t)u * 16u;
        }
        uint4 h, l;
        h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
        l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
        *reinterpret_cast<uint4*>(hi + dst) = h;
        *reinterpret_cast<uint4*>(lo + dst) = l;
    }
}

template <int NT, bool AT, bool BT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const GtParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t a_bytes = TILE_M * TILE_K * 4;          // 16 KB
    constexpr uint32_t b_bytes = (uint32_t)NT * TILE_K * 4;
    constexpr uint32_t raw_stage = a_bytes + b_bytes;           // A | B
    constexpr uint32_t op_stage = 2 * (a_bytes + b_bytes);      // A hi | A lo | B hi | B lo
    constexpr uint32_t TMEM_COLS = 2 * NT <= 128 ? 128 : 256;   // double-buffered partial accumulator
    unsigned char* op_base = smem + (size_t)RAW_STAGES * raw_stage;
    unsigned char* bars = op_base + (size_t)OP_STAGES * op_stage;
    const uint32_t bar0 = smem_u32(bars);
    auto raw_full = [&](int s) { return bar0 + 8u * s; };
    auto raw_empty = [&](int s) { return bar0 + 8u * (RAW_STAGES + s); };
    auto op_ready = [&](int s) { return bar0 + 8u * (2 * RAW_STAGES + s); };
    auto op_empty = [&](int s) { return bar0 + 8u * (2 * RAW_STAGES + OP_STAGES + s); };
    auto pfull = [&](int b) { return bar0 + 8u * (2 * RAW_STAGES + 2 * OP_STAGES + b); };
    auto drained = [&](int b) { return bar0 + 8u * (2 * RAW_STAGES + 2 * OP_STAGES + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 8 * (2 * RAW_STAGES + 2 * OP_STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TILE_M;
    const int n0 = blockIdx.x * NT;
    const int nkb_all = (p.K + TILE_K - 1) / TILE_K;
    const int kb_lo = blockIdx.z * p.kb_per_split;
    const int nkb = min(nkb_all, kb_lo + p.kb_per_split) - kb_lo;   // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < RAW_STAGES; s++) {
            mbar_init(raw_full(s), 1);
            mbar_init(raw_empty(s), WORKER_WARPS);
        }
        for (int s = 0; s < OP_STAGES; s++) {
            mbar_init(op_ready(s), WORKER_WARPS);
            mbar_init(op_empty(s), 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(pfull(b), 1);
            mbar_init(drained(b), WORKER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (elect_one()) {
            for (int i = 0; i < nkb; i++) {
                const int s = i % RAW_STAGES;
                const int k0 = (kb_lo + i) * TILE_K;
                mbar_wait(raw_empty(s), ((i / RAW_STAGES) & 1) ^ 1);
                const uint32_t a_raw = smem_u32(smem + (size_t)s * raw_stage);
                const uint32_t b_raw = a_raw + a_bytes;
                mbar_arrive_expect_tx(raw_full(s), a_bytes + b_bytes);
                if (AT) tma_load_2d(a_raw, &map_a, m0, k0, raw_full(s)); else tma_load_2d(a_raw, &map_a, k0, m0, raw_full(s));
                if (BT) tma_load_2d(b_raw, &map_b, n0, k0, raw_full(s)); else tma_load_2d(b_raw, &map_b, k0, n0, raw_full(s));
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_tf32(TILE_M, NT);
        for (int i = 0; i < nkb; i++) {
            const int so = i % OP_STAGES;
            const int ch = i / DRAIN_KB, pb = ch & 1;
            const bool first = i % DRAIN_KB == 0, last = i % DRAIN_KB == DRAIN_KB - 1 || i == nkb - 1;
            mbar_wait(op_ready(so), (i / OP_STAGES) & 1);
            if (first) mbar_wait(drained(pb), ((ch >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t a_hi = smem_u32(op_base + (size_t)so * op_stage);
                const uint32_t a_lo = a_hi + a_bytes;
                const uint32_t b_hi = a_lo + a_bytes;
                const uint32_t b_lo = b_hi + b_bytes;
                const uint32_t d = tmem_base + (uint32_t)(pb * NT);
#pragma unroll
                for (int k = 0; k < TILE_K / UMMA_K; k++) {   // the two correction products first (tiny partial sums)
                    const uint32_t koff = (uint32_t)k * UMMA_K * 4;
                    umma_tf32(d, umma_desc_k_sw128(a_lo + koff), umma_desc_k_sw128(b_hi + koff), idesc, (k > 0 || !first) ? 1u : 0u);
                    umma_tf32(d, umma_desc_k_sw128(a_hi + koff), umma_desc_k_sw128(b_lo + koff), idesc, 1u);
                }
#pragma unroll
                for (int k = 0; k < TILE_K / UMMA_K; k++) {
                    const uint32_t koff = (uint32_t)k * UMMA_K * 4;
                    umma_tf32(d, umma_desc_k_sw128(a_hi + koff), umma_desc_k_sw128(b_hi + koff), idesc, 1u);
                }
                umma_commit(op_empty(so));
                if (last) umma_commit(pfull(pb));
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------ workers: split / transpose, drain, epilogue
        const int t = threadIdx.x - 64;
        const int q = warp & 3;                 // TMEM lanes [32q, 32q + 32)
        const int half = (warp - 2) >> 2;       // half of the columns
        constexpr int NC = NT / 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * NC);
        float tot[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) tot[j] = 0.f;
        auto drain = [&](int ch) {
            const int pb = ch & 1;
            mbar_wait(pfull(pb), (ch >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[NC];
#pragma unroll
            for (int c = 0; c < NC; c += 16) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[c + 0]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]),
                      "=r"(r[c + 6]), "=r"(r[c + 7]), "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]),
                      "=r"(r[c + 12]), "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
                    : "r"(taddr + (uint32_t)(pb * NT + c)));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < NC; j++) tot[j] += __uint_as_float(r[j]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(drained(pb));
        };
        for (int i = 0; i < nkb; i++) {
            const int s = i % RAW_STAGES, so = i % OP_STAGES;
            mbar_wait(raw_full(s), (i / RAW_STAGES) & 1);
            mbar_wait(op_empty(so), ((i / OP_STAGES) & 1) ^ 1);
            const unsigned char* a_raw = smem + (size_t)s * raw_stage;
            const unsigned char* b_raw = a_raw + a_bytes;
            unsigned char* a_hi = op_base + (size_t)so * op_stage;
            unsigned char* b_hi = a_hi + 2 * a_bytes;
            split_tile<TILE_M, AT>(a_raw, a_hi, a_hi + a_bytes, t);
            split_tile<NT, BT>(b_raw, b_hi, b_hi + b_bytes, t);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(op_ready(so));
                mbar_arrive(raw_empty(s));
            }
            if (i > 0 && i % DRAIN_KB == 0) drain(i / DRAIN_KB - 1);
        }
        drain((nkb - 1) / DRAIN_KB);
        // ---- epilogue: a thread owns an accumulator ROW, so direct stores would put one 32-byte sector per lane and
        // instruction on the wire.  The tile is staged in shared memory instead (the operand ring is free: the last
        // drain saw every MMA complete) as NT/32 sub-tiles of [128 rows x 128 bytes] with the 128-byte XOR swizzle
        // (conflict-free float4 writes), then all worker threads write it out row-contiguously (8 lanes = one 128-byte
        // line).  A final tile gets beta * C, bias and the activation on the way; partial tiles go out raw.
        const int r = q * 32 + lane;
        const int c0 = half * NC;
#pragma unroll
        for (int c = 0; c < NC; c += 4) {
            const int col = c0 + c, sub = col >> 5, chunk = (col & 31) >> 2;
            *reinterpret_cast<float4*>(op_base + (size_t)sub * (TILE_M * 128) + r * 128 + ((chunk ^ (r & 7)) << 4)) =
                make_float4(tot[c + 0], tot[c + 1], tot[c + 2], tot[c + 3]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WORK_THREADS) : "memory");
        const bool fin = p.m_pad == 0;
        float* out = fin ? p.C : p.ws + (size_t)blockIdx.z * p.m_pad * p.N;
        const int64_t ld_out = fin ? p.ldc : (int64_t)p.N;
        auto finish = [&](float4 v, int gm, int gn) {      // beta * C + bias, activation, store to C
            float* dst = p.C + (int64_t)gm * p.ldc + gn;
            if (p.bias) {
                v.x += __ldg(p.bias + gn); v.y += __ldg(p.bias + gn + 1);
                v.z += __ldg(p.bias + gn + 2); v.w += __ldg(p.bias + gn + 3);
            }
            if (p.beta != 0.f) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                v.x = fmaf(p.beta, o.x, v.x); v.y = fmaf(p.beta, o.y, v.y);
                v.z = fmaf(p.beta, o.z, v.z); v.w = fmaf(p.beta, o.w, v.w);
            }
            v.x = gt_act(v.x, p.act); v.y = gt_act(v.y, p.act); v.z = gt_act(v.z, p.act); v.w = gt_act(v.w, p.act);
            *reinterpret_cast<float4*>(dst) = v;
        };
#pragma unroll
        for (int sub = 0; sub < NT / 32; sub++) {
#pragma unroll
            for (int it = 0; it < (TILE_M * 8) / WORK_THREADS; it++) {
                const int idx = t + it * WORK_THREADS;
                const int row = idx >> 3, chunk = idx & 7;
                const int gm = m0 + row, gn = n0 + sub * 32 + chunk * 4;
                if (gm < p.M && gn < p.N) {            // N % 4 == 0: a float4 never straddles the edge
                    const float4 v = *reinterpret_cast<const float4*>(op_base + (size_t)sub * (TILE_M * 128) + row * 128 + ((chunk ^ (row & 7)) << 4));
                    if (fin) finish(v, gm, gn);
                    else __stcg(reinterpret_cast<float4*>(out + (int64_t)gm * ld_out + gn), v);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// C = act(beta * C + sum_s ws[s] + bias); partial sums in split order (deterministic); float4 along n when aligned
__global__ void gemm_tc_reduce_kernel(const float* __restrict__ ws, int splits, int64_t split_stride, float* __restrict__ C,
                                      int64_t ldc, const float* __restrict__ bias, int M, int N, float beta, int act) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (int64_t)m * N);
    float v = 0.f;
    for (int s = 0; s < splits; s++) v += ws[(int64_t)s * split_stride + i];
    if (bias) v += __ldg(bias + n);
    float* c = C + (int64_t)m * ldc + n;
    if (beta != 0.f) v += beta * *c;
    *c = gt_act(v, act);
}

template <int NT>
const void* pick_kernel(bool at, bool bt) {
    if (at) return bt ? (const void*)gemm_tc_kernel<NT, true, true> : (const void*)gemm_tc_kernel<NT, true, false>;
    return bt ? (const void*)gemm_tc_kernel<NT, false, true> : (const void*)gemm_tc_kernel<NT, false, false>;
}

size_t smem_bytes(int NT) {
    const size_t a = TILE_M * TILE_K * 4, b = (size_t)NT * TILE_K * 4;
    return RAW_STAGES * (a + b) + OP_STAGES * 2 * (a + b) + 8 * (2 * RAW_STAGES + 2 * OP_STAGES + 6) + 1024;
}

// per-device, per-slot workspace for partial tiles (mlb_policy.cu: mlb_workspace_get), grown outside stream capture only
float* workspace(size_t need, cudaStream_t stream) { return mlb_workspace_get(1, need, stream); }

}  // namespace

extern "C" {

// Which (layout, size) combinations the tensor-core GEMM takes; everything else stays on mlb_gemm's FFMA kernel.
int mlb_gemm_tc_supported(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                          const float* C, int64_t ldc, int32_t M, int32_t N, int32_t K) {
    if (!A || !B || !C || M < 1 || N < 32 || K < 32) return 0;
    const bool at = a_cs != 1;          // A(m, k): contraction contiguous (a_cs == 1) or rows contiguous (a_rs == 1)
    const bool bt = b_rs != 1;          // B(k, n): contraction contiguous (b_rs == 1) or columns contiguous (b_cs == 1)
    if (at && a_rs != 1) return 0;
    if (bt && b_cs != 1) return 0;
    const int64_t lda = at ? a_cs : a_rs, ldb = bt ? b_rs : b_cs;
    if ((lda % 4) || (ldb % 4) || (ldc % 4) || (N % 4)) return 0;
    if (lda < (at ? M : K) || ldb < (bt ? N : K) || ldc < N) return 0;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) return 0;
    return 1;
}

int mlb_gemm_tc(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs, float* C,
                int64_t ldc, const float* bias, int32_t M, int32_t N, int32_t K, float beta, int32_t act, void* stream) {
    if (!mlb_gemm_tc_supported(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K)) return MLB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const bool at = a_cs != 1, bt = b_rs != 1;
    const int64_t lda = at ? a_cs : a_rs, ldb = bt ? b_rs : b_cs;
    const int NT = N <= 64 ? 64 : 128;
    const int tiles_m = (M + TILE_M - 1) / TILE_M, tiles_n = (N + NT - 1) / NT;
    const int nkb = (K + TILE_K - 1) / TILE_K;
    // split K until the grid covers the SMs about once, keeping at least two k-blocks per CTA
    int splits = 1;
    const int tiles = tiles_m * tiles_n;
    // (measured on the SAC update at C4 sizes: a CTA spends ~1.2 us per k-block, so even 4-8 k-blocks are worth a
    // second launch that adds the partial tiles)
    if (tiles < 120 && nkb >= 4) splits = std::max(1, std::min({148 / tiles, nkb / 2, 32}));
    int kb_per = (nkb + splits - 1) / splits;
    splits = (nkb + kb_per - 1) / kb_per;                       // no empty CTA
    const bool partial = splits > 1;
    const int m_pad = tiles_m * TILE_M;
    float* ws = nullptr;
    if (partial) {
        ws = workspace((size_t)splits * m_pad * N * sizeof(float), st);
        if (!ws) return MLB_ENOMEM;                             // caller falls back to the FFMA kernel
    }
    CUtensorMap ma, mb;
    const bool ok_a = at ? make_map_2d(&ma, A, M, K, lda, TILE_M, TILE_K, false) : make_map_2d(&ma, A, K, M, lda, TILE_K, TILE_M, true);
    const bool ok_b = bt ? make_map_2d(&mb, B, N, K, ldb, NT, TILE_K, false) : make_map_2d(&mb, B, K, N, ldb, TILE_K, NT, true);
    if (!ok_a || !ok_b) return MLB_ESTATE;            // nothing launched: the caller may use the FFMA kernel
    GtParams p{C, ws, partial ? nullptr : bias, ldc, beta, M, N, K, act, kb_per, partial ? m_pad : 0};
    const void* fn = NT == 64 ? pick_kernel<64>(at, bt) : pick_kernel<128>(at, bt);
    const size_t smem = smem_bytes(NT);
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MLB_ECUDA;
    dim3 grid(tiles_n, tiles_m, splits);
    void* args[] = {&ma, &mb, &p};
    if (cudaLaunchKernel(fn, grid, dim3(THREADS), args, smem, st) != cudaSuccess) return MLB_ECUDA;
    if (partial) {
        const int64_t total = (int64_t)M * N;
        gemm_tc_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, splits, (int64_t)m_pad * N, C, ldc, bias,
                                                                              M, N, beta, act);
    }
    return cudaGetLastError() == cudaSuccess ? MLB_OK : MLB_ECUDA;
}

}  // extern "C"
