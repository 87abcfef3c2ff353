// Warp-cooperative reservoir statistics (one warp per reservoir).
//
// Computes the five features of ReservoirSampler.get_features
// (reference: simulation-mode/problem-01-reservoir-sampling/src/reservoir.py:105-196)
//   mean, p90 (numpy 'linear' percentile), std (ddof 0),
//   mean_decay = sum(v*w)/sum(w), p90_decay = weighted percentile,
//   with w = decay^(now - timestamp).
//
// Layout: the K<=128 slots of one reservoir are spread over the 32 lanes,
// EPL = 1, 2 or 4 consecutive slots per lane (one 32/64/128-bit load each).
//
// Order statistics come from a per-slot RANK (position of the slot's value in
// ascending order).  Ranks are either computed from scratch -- an in-register
// bitonic network over (value, slot) pairs, intra-lane stages as compare-
// exchanges and inter-lane stages as shuffles -- or, when the caller keeps the
// ranks of the previous step, updated incrementally for the few slots that
// Algorithm R replaced (remove the old rank, count-and-insert the new value:
// ~50 instructions per replaced slot instead of a ~600-instruction sort).
// Values and weights are then scattered by rank into a 1 KB per-warp shared-
// memory scratch, which makes "value at rank r" a broadcast read and the
// weighted percentile a blocked read + warp scan.
//
// The decay weights only matter up to a common factor (both weighted outputs
// are ratios / order decisions), so the fast path evaluates them in float32
// relative to the newest timestamp, w = 2^(log2(decay) * (t_max - t)), and
// takes the weighted percentile index from a float32 warp scan.  If any
// cumulative weight lies within a relative margin of the 0.9*W cutoff the
// warp (uniformly) re-does the decision the way the reference does it:
// float64 pow(decay, now - t) and a strictly sequential float64 cumsum in
// value order (np.cumsum + np.searchsorted, reservoir.py:186-190).
#pragma once
#include "mlb_common.cuh"

namespace mlb {

// relative distance to the cutoff below which the float32 index decision is
// not trusted (error bound of the fast path is < 1e-5 * W)
#define MLB_WP_MARGIN 5e-5f

// One compare-exchange stage of the bitonic network; SIZE / STRIDE are template
// constants so every register index is static (no local-memory arrays).
template <int EPL, int SIZE, int STRIDE>
__device__ __forceinline__ void bitonic_stage(float (&k)[EPL], float (&p)[EPL], int lane) {
    // Branch-free compare-exchange: the new key is min or max (FMNMX + FSEL); the payload
    // follows iff the key changed, so equal keys keep their own payload on both sides.
    if constexpr (STRIDE >= EPL) {
        constexpr int LSTRIDE = STRIDE / EPL;
        const bool keep_min = (((lane * EPL) & SIZE) == 0) == ((lane & LSTRIDE) == 0);
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const float ok = __shfl_xor_sync(MLB_FULL, k[r], LSTRIDE);
            const float op = __shfl_xor_sync(MLB_FULL, p[r], LSTRIDE);
            const float nk = keep_min ? fminf(k[r], ok) : fmaxf(k[r], ok);
            p[r] = (nk != k[r]) ? op : p[r];
            k[r] = nk;
        }
    } else {
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            if ((r & STRIDE) == 0) {
                const int r2 = r | STRIDE;
                const bool up = (((lane * EPL + r) & SIZE) == 0);
                const float a = k[r], b = k[r2], pa = p[r], pb = p[r2];
                const float lo = fminf(a, b), hi = fmaxf(a, b);
                const float na = up ? lo : hi;
                const bool sw = na != a;
                k[r] = na;
                k[r2] = up ? hi : lo;
                p[r] = sw ? pb : pa;
                p[r2] = sw ? pa : pb;
            }
        }
    }
    if constexpr (STRIDE > 1) bitonic_stage<EPL, SIZE, STRIDE / 2>(k, p, lane);
}

template <int EPL, int SIZE>
__device__ __forceinline__ void bitonic_merge_levels(float (&k)[EPL], float (&p)[EPL], int lane) {
    bitonic_stage<EPL, SIZE, SIZE / 2>(k, p, lane);
    if constexpr (SIZE < 32 * EPL) bitonic_merge_levels<EPL, SIZE * 2>(k, p, lane);
}

// ascending sort of the 32*EPL (key, payload) pairs, element index = lane*EPL + r
template <int EPL>
__device__ __forceinline__ void bitonic_sort_kv(float (&k)[EPL], float (&p)[EPL], int lane) {
    bitonic_merge_levels<EPL, 2>(k, p, lane);
}

// blocked read of EPL consecutive floats per lane from shared memory
template <int EPL>
__device__ __forceinline__ void load_smem_block(const float* base, int lane, float (&x)[EPL]) {
    if constexpr (EPL == 4) {
        const float4 q = *reinterpret_cast<const float4*>(base + lane * 4);
        x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else if constexpr (EPL == 2) {
        const float2 q = *reinterpret_cast<const float2*>(base + lane * 2);
        x[0] = q.x; x[1] = q.y;
    } else {
        x[0] = base[lane];
    }
}

// element `slot` (warp-uniform) of an array spread EPL-per-lane.  Written as a select tree
// on the sub-index bits: a sequential `if (sub == r) c = x[r]` chain makes nvcc materialise
// x[] as a dynamically indexed local-memory array.
template <int EPL, typename T>
__device__ __forceinline__ T slot_fetch(const T (&x)[EPL], int slot) {
    T c;
    if constexpr (EPL == 4) {
        const int sub = slot & 3;
        const T lo = (sub & 1) ? x[1] : x[0];
        const T hi = (sub & 1) ? x[3] : x[2];
        c = (sub & 2) ? hi : lo;
    } else if constexpr (EPL == 2) {
        c = (slot & 1) ? x[1] : x[0];
    } else {
        c = x[0];
    }
    return __shfl_sync(MLB_FULL, c, slot / EPL);
}

template <int EPL>
__device__ __forceinline__ void load_slots(const float* __restrict__ base, int lane, float (&x)[EPL]) {
    if constexpr (EPL == 4) {
        const float4 q = ldg_stream4(base + lane * 4);
        x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else if constexpr (EPL == 2) {
        const float2 q = ldg_stream2(base + lane * 2);
        x[0] = q.x; x[1] = q.y;
    } else {
        x[0] = ldg_stream1(base + lane);
    }
}

// ranks are one byte per slot in global memory
template <int EPL>
__device__ __forceinline__ void load_ranks(const uint8_t* __restrict__ base, int lane, int (&rk)[EPL]) {
    if constexpr (EPL == 4) {
        const uint32_t q = *reinterpret_cast<const uint32_t*>(base + lane * 4);
        rk[0] = q & 255; rk[1] = (q >> 8) & 255; rk[2] = (q >> 16) & 255; rk[3] = q >> 24;
    } else if constexpr (EPL == 2) {
        const uint32_t q = *reinterpret_cast<const uint16_t*>(base + lane * 2);
        rk[0] = q & 255; rk[1] = q >> 8;
    } else {
        rk[0] = base[lane];
    }
}
template <int EPL>
__device__ __forceinline__ void store_ranks(uint8_t* __restrict__ base, int lane, const int (&rk)[EPL]) {
    if constexpr (EPL == 4) {
        *reinterpret_cast<uint32_t*>(base + lane * 4) =
            (uint32_t)(rk[0] & 255) | ((uint32_t)(rk[1] & 255) << 8) | ((uint32_t)(rk[2] & 255) << 16) | ((uint32_t)rk[3] << 24);
    } else if constexpr (EPL == 2) {
        *reinterpret_cast<uint16_t*>(base + lane * 2) = (uint16_t)((rk[0] & 255) | ((rk[1] & 255) << 8));
    } else {
        base[lane] = (uint8_t)rk[0];
    }
}

// per-warp shared-memory scratch: sv[128] (values by rank), sw[128] (weights by rank)
struct WarpScratch {
    float* sv;
    float* sw;
};
#define MLB_SCRATCH_BYTES 1024

// Ranks from scratch: bitonic sort of (value, slot) pairs, then transpose position->slot
// into slot->position through shared-memory bytes.  Slots >= n get rank 255.
// noinline: cold once reservoirs are full and ranks are maintained incrementally; keeping
// it out of line keeps the steady-state loop inside the instruction cache.  Values go in by
// value and the EPL ranks come back packed in one word so that nothing lives in local memory.
template <int EPL>
struct SlotVals {
    float x[EPL];
};

template <int EPL>
__device__ __noinline__ uint32_t ranks_full_sort_packed(SlotVals<EPL> v, int n, float* sw) {
    const int lane = lane_id();
    float k[EPL], p[EPL];
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const int slot = lane * EPL + r;
        k[r] = slot < n ? v.x[r] : MLB_INF;  // padding sorts to the end
        p[r] = __int_as_float(slot);
    }
    bitonic_sort_kv<EPL>(k, p, lane);
    uint8_t* sb = reinterpret_cast<uint8_t*>(sw);
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const int pos = lane * EPL + r;
        sb[__float_as_int(p[r])] = (uint8_t)(pos < n ? pos : 255);
    }
    __syncwarp();
    uint32_t packed = 0;
#pragma unroll
    for (int r = 0; r < EPL; r++) packed |= (uint32_t)sb[lane * EPL + r] << (8 * r);
    __syncwarp();
    return packed;
}

template <int EPL>
__device__ __forceinline__ void ranks_full_sort(const float (&v)[EPL], int n, int (&rk)[EPL],
                                                const WarpScratch& sc, int lane) {
    SlotVals<EPL> sv;
#pragma unroll
    for (int r = 0; r < EPL; r++) sv.x[r] = v[r];
    const uint32_t packed = ranks_full_sort_packed<EPL>(sv, n, sc.sw);
#pragma unroll
    for (int r = 0; r < EPL; r++) rk[r] = (packed >> (8 * r)) & 255u;
    (void)lane;
}

// Incremental rank maintenance.  `mw` = bit mask (warp-uniform) of the slots Algorithm R
// wrote this step; slots < n_old existed before (their stored rank is removed), every
// changed slot < n_new is (re-)inserted by counting the present elements below it.
// Ties are ordered by slot index; any order of equal values is a valid sorted order.
template <int EPL>
__device__ __forceinline__ void ranks_update(const float (&v)[EPL], int (&rk)[EPL], int n_new, int n_old,
                                             const uint32_t* mws, int stride, int nchg,
                                             const WarpScratch& sc, int lane) {
    // mws[k * stride], k = 0..3: mask words in shared memory (warp-uniform addresses).
    // Every changed slot is published as (new value, slot, old rank) in a small shared list;
    // then each lane fixes the ranks of its own slots in ONE pass over that list:
    //   unchanged slot i : rank - #{removed with lower old rank} + #{inserted (x,c) < (v_i,i)}
    //   changed slot c   : #{unchanged (v_i,i) < (x_c,c)} + #{inserted (x,c') < (x_c,c)}
    constexpr int NW = EPL == 4 ? 4 : EPL;  // words that can hold slots < 32*EPL
    const int s0 = lane * EPL;
    const int wi = s0 >> 5;
    uint32_t word = mws[0];
    int below = 0;
    if constexpr (NW >= 2) {
        const uint32_t m1 = mws[stride];
        below = wi >= 1 ? __popc(word) : 0;
        word = wi == 1 ? m1 : word;
        if constexpr (NW == 4) {
            const uint32_t m2 = mws[2 * stride], m3 = mws[3 * stride];
            below += wi >= 2 ? __popc(m1) : 0;
            below += wi >= 3 ? __popc(m2) : 0;
            word = wi == 2 ? m2 : (wi == 3 ? m3 : word);
        }
    }
    const uint32_t chg = (word >> (s0 & 31)) & ((1u << EPL) - 1u);
    below += __popc(word & ((1u << (s0 & 31)) - 1u));
    float* Lx = sc.sv;
    int* Lc = reinterpret_cast<int*>(sc.sw);
    int* Lr = Lc + 32;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        if ((chg >> r) & 1u) {
            Lx[below] = v[r];
            Lc[below] = s0 + r;
            Lr[below] = s0 + r < n_old ? rk[r] : 1000;  // 1000: did not exist, removes nothing
            below++;
        }
    }
    __syncwarp();
    int adj[EPL], self[EPL];
#pragma unroll
    for (int r = 0; r < EPL; r++) { adj[r] = 0; self[r] = 0; }
#pragma unroll 1
    for (int k = 0; k < nchg; k++) {
        const float x = Lx[k];
        const int c = Lc[k], rold = Lr[k];
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const int slot = s0 + r;
            const bool lt = x < v[r] || (x == v[r] && c < slot);  // (x_k, c_k) < (v_i, i)
            const bool unchanged = slot < n_new && !((chg >> r) & 1u);
            adj[r] += lt ? 1 : 0;
            adj[r] -= (unchanged && rk[r] > rold) ? 1 : 0;
            cnt += (unchanged && !lt) ? 1 : 0;  // strict total order: !lt <=> (v_i, i) < (x_k, c_k)
        }
        const int tot = __reduce_add_sync(MLB_FULL, cnt);
#pragma unroll
        for (int r = 0; r < EPL; r++) self[r] = (s0 + r == c) ? tot : self[r];
    }
#pragma unroll
    for (int r = 0; r < EPL; r++) rk[r] = (((chg >> r) & 1u) ? self[r] : rk[r]) + adj[r];
    __syncwarp();  // the list lives in the scratch that features_ranked overwrites next
}

// Cold paths of the weighted percentile, taken when the float32 cumulative weights come
// within MLB_WP_MARGIN of the cutoff.  st = timestamps in rank order (shared memory), n <= 128.
//
// Tier 2: float64 weights 2^(log2(decay) * (tmax - t)) and a float64 warp scan; decides unless
// a cumulative weight is within 1e-9 (relative) of the cutoff, i.e. a genuine tie.
static __device__ __noinline__ int weighted_index_scan_f64(const float* st, int n, float tmax, double decay,
                                                           bool* ambiguous) {
    const int lane = lane_id();
    const double l2d = log2(decay);
    double w[4], c[4];
    double carry = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {  // position p = k*32 + lane
        const int pos = k * 32 + lane;
        w[k] = pos < n ? exp2(l2d * ((double)tmax - (double)st[pos])) : 0.0;
        double sc = w[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(MLB_FULL, sc, o);
            if (lane >= o) sc += y;
        }
        c[k] = carry + sc;
        carry = __shfl_sync(MLB_FULL, c[k], 31);
    }
    const double cut = 0.9 * carry;
    int below = 0;
    double dmin = 1.0e308;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool valid = k * 32 + lane < n;
        below += (valid && c[k] < cut) ? 1 : 0;
        dmin = valid ? fmin(dmin, fabs(c[k] - cut)) : dmin;
    }
    dmin = warp_min(dmin);
    *ambiguous = dmin < 1e-9 * carry;
    return __reduce_add_sync(MLB_FULL, below);
}

// Tier 3: the reference's own arithmetic (reservoir.py:148-149,181-196): w = pow(decay, now - t)
// and a strictly sequential cumsum in value order.  Returns the searchsorted-left index of
// 0.9 * cumsum[-1] (n if none).
static __device__ __noinline__ int weighted_index_f64(const float* st, int n, float now, double decay) {
    const int lane = lane_id();
    double w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int pos = k * 32 + lane;
        w[k] = pos < n ? pow(decay, (double)now - (double)st[pos]) : 0.0;
    }
    double c = 0.0;
#pragma unroll 1
    for (int p = 0; p < n; p++) {
        const int k = p >> 5;
        const double wk = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
        c += __shfl_sync(MLB_FULL, wk, p & 31);
    }
    const double cut = 0.9 * c;  // percentile * cumsum[-1]
    c = 0.0;
    int idx = n;
#pragma unroll 1
    for (int p = 0; p < n; p++) {
        const int k = p >> 5;
        const double wk = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
        c += __shfl_sync(MLB_FULL, wk, p & 31);
        if (c >= cut) {
            idx = p;
            break;
        }
    }
    return idx;
}

// The five features given slot-ordered values/timestamps and valid ranks.
template <int EPL>
__device__ __forceinline__ void features_ranked(const float (&v)[EPL], const float (&t)[EPL],
                                                const int (&rk)[EPL], int n, float now, double decay,
                                                float log2_decay, const WarpScratch& sc, float (&out)[5]) {
    const int lane = lane_id();
    const int s0 = lane * EPL;
    // ---- mean / std (np.mean, np.std on float32: reservoir.py:143,145)
    float s = 0.f, tmax = -MLB_INF;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = s0 + r < n;
        s += valid ? v[r] : 0.f;
        tmax = valid ? fmaxf(tmax, t[r]) : tmax;
    }
    const float nf = (float)n;
    const float mean = warp_sum(s) / nf;
    tmax = warp_max(tmax);
    // ---- decay weights relative to the newest sample (float32 fast path), scatter by rank
    float s2 = 0.f, svw = 0.f;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = s0 + r < n;
        const float d = v[r] - mean;
        s2 += valid ? d * d : 0.f;
        const float w = valid ? fast_exp2(log2_decay * (tmax - t[r])) : 0.f;
        svw += valid ? v[r] * w : 0.f;
        const int pos = valid ? rk[r] : s0 + r;  // ranks cover [0,n); slots >= n pad positions >= n
        sc.sv[pos] = v[r];
        sc.sw[pos] = w;
    }
    const float sd = sqrtf(warp_sum(s2) / nf);
    svw = warp_sum(svw);
    __syncwarp();
    // ---- weighted percentile: blocked read of weights in rank order + warp scan
    float wq[EPL], cum[EPL];
    load_smem_block<EPL>(sc.sw, lane, wq);
    float sw = 0.f;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        sw += wq[r];
        cum[r] = sw;
    }
    const float incl = warp_scan_incl(sw, lane);
    const float W = __shfl_sync(MLB_FULL, incl, 31);
    const float excl = incl - sw;
    const float mean_decay = svw / W;
    const float cutoff = 0.9f * W;
    int below = 0;
    float dmin = MLB_INF;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = s0 + r < n;  // here s0+r is a POSITION in rank order
        const float c = excl + cum[r];
        below += (valid && c < cutoff) ? 1 : 0;
        dmin = valid ? fminf(dmin, fabsf(c - cutoff)) : dmin;
    }
    int idx = __reduce_add_sync(MLB_FULL, below);
    const uint32_t dmin_bits = __reduce_min_sync(MLB_FULL, __float_as_uint(dmin));
    if (__uint_as_float(dmin_bits) < MLB_WP_MARGIN * W) {
        // ---- not trusted: redo the decision in the reference's float64 arithmetic
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EPL; r++)
            if (s0 + r < n) sc.sw[rk[r]] = t[r];
        __syncwarp();
        bool ambiguous;
        idx = weighted_index_scan_f64(sc.sw, n, tmax, decay, &ambiguous);
        if (ambiguous) idx = weighted_index_f64(sc.sw, n, now, decay);
    }
    idx = idx > n - 1 ? n - 1 : idx;  // reservoir.py:193-194
    // ---- p90 = np.percentile(values, 90): float32 'linear' rule of numpy >= 2
    const float vidx = (float)(n - 1) * (90.0f / 100.0f);
    const float fl = floorf(vidx);
    int lo = (int)fl, hi = lo + 1;
    if (vidx >= (float)(n - 1)) { lo = n - 1; hi = n - 1; }
    hi = hi > n - 1 ? n - 1 : hi;
    const float gamma = vidx - fl;
    const float a = sc.sv[lo], b = sc.sv[hi];
    const float p90_decay = sc.sv[idx];
    const float diff = __fsub_rn(b, a);
    float p90 = __fadd_rn(a, __fmul_rn(diff, gamma));
    if (gamma >= 0.5f) p90 = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
    __syncwarp();  // scratch is reused by the next reservoir
    out[0] = mean;
    out[1] = p90;
    out[2] = sd;
    out[3] = mean_decay;
    out[4] = p90_decay;
}

// Stateless form: ranks from scratch every time.  n = min(count, K), may be 0.
__device__ __forceinline__ void warp_features(const float* __restrict__ vals,
                                              const float* __restrict__ tss, int n, float now,
                                              double decay, float log2_decay, const WarpScratch& sc,
                                              float (&out)[5]) {
    const int lane = lane_id();
    if (n <= 0) {  // reservoir.py:127-134
#pragma unroll
        for (int q = 0; q < 5; q++) out[q] = 0.f;
    } else if (n <= 32) {
        float v[1], t[1];
        int rk[1];
        load_slots<1>(vals, lane, v);
        load_slots<1>(tss, lane, t);
        ranks_full_sort<1>(v, n, rk, sc, lane);
        features_ranked<1>(v, t, rk, n, now, decay, log2_decay, sc, out);
    } else if (n <= 64) {
        float v[2], t[2];
        int rk[2];
        load_slots<2>(vals, lane, v);
        load_slots<2>(tss, lane, t);
        ranks_full_sort<2>(v, n, rk, sc, lane);
        features_ranked<2>(v, t, rk, n, now, decay, log2_decay, sc, out);
    } else {
        float v[4], t[4];
        int rk[4];
        load_slots<4>(vals, lane, v);
        load_slots<4>(tss, lane, t);
        ranks_full_sort<4>(v, n, rk, sc, lane);
        features_ranked<4>(v, t, rk, n, now, decay, log2_decay, sc, out);
    }
}

// Stateful form used by the env step: ranks live in global memory next to the reservoir
// and are updated incrementally when only a few slots changed.
//   n_old  valid slots when the stored ranks were computed (0: no ranks yet)
//   mws    mask (4 words, shared memory, word k at mws[k*stride]) of slots written since then
//   force_full: ignore stored ranks (feature_cache modes 0 / 2)
template <int EPL>
__device__ __forceinline__ void features_cached_epl(const float* __restrict__ vals, const float* __restrict__ tss,
                                                    uint8_t* __restrict__ ranks, int n, int n_old,
                                                    const uint32_t* mws, int stride, int nchg, bool force_full,
                                                    float now, double decay, float log2_decay,
                                                    const WarpScratch& sc, float (&out)[5]) {
    const int lane = lane_id();
    float v[EPL], t[EPL];
    int rk[EPL];
    load_slots<EPL>(vals, lane, v);
    load_slots<EPL>(tss, lane, t);
    // stored ranks are only meaningful if they were laid out for a population that this
    // EPL variant also covers (n_old <= 32*EPL always holds since n_old <= n)
    constexpr int kMaxIncremental = EPL == 1 ? 2 : (EPL == 2 ? 5 : 10);  // ~50 instr per slot vs the sort
    if (force_full || n_old == 0 || nchg > kMaxIncremental) {
        ranks_full_sort<EPL>(v, n, rk, sc, lane);
    } else {
        load_ranks<EPL>(ranks, lane, rk);
        ranks_update<EPL>(v, rk, n, n_old, mws, stride, nchg, sc, lane);
    }
    store_ranks<EPL>(ranks, lane, rk);
    features_ranked<EPL>(v, t, rk, n, now, decay, log2_decay, sc, out);
}

__device__ __forceinline__ void warp_features_cached(const float* __restrict__ vals, const float* __restrict__ tss,
                                                     uint8_t* __restrict__ ranks, int n, int n_old,
                                                     const uint32_t* mws, int stride, int nchg, bool force_full,
                                                     float now, double decay, float log2_decay,
                                                     const WarpScratch& sc, float (&out)[5]) {
    if (n <= 0) {
#pragma unroll
        for (int q = 0; q < 5; q++) out[q] = 0.f;
    } else if (n <= 32) {
        features_cached_epl<1>(vals, tss, ranks, n, n_old, mws, stride, nchg, force_full, now, decay, log2_decay, sc, out);
    } else if (n <= 64) {
        features_cached_epl<2>(vals, tss, ranks, n, n_old, mws, stride, nchg, force_full, now, decay, log2_decay, sc, out);
    } else {
        features_cached_epl<4>(vals, tss, ranks, n, n_old, mws, stride, nchg, force_full, now, decay, log2_decay, sc, out);
    }
}

}  // namespace mlb
