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
// Sorting is an in-register bitonic network over (value, timestamp) pairs:
// intra-lane stages are compare-exchanges, inter-lane stages use shuffles.
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

// value at sorted position `pos` (warp-uniform).  Written as a select tree on
// the sub-index bits: a sequential `if (sub == r) c = k[r]` chain makes nvcc
// materialise k[] as a dynamically indexed local-memory array.
template <int EPL>
__device__ __forceinline__ float sorted_at(const float (&k)[EPL], int pos) {
    float c;
    if constexpr (EPL == 4) {
        const int sub = pos & 3;
        const float lo = (sub & 1) ? k[1] : k[0];
        const float hi = (sub & 1) ? k[3] : k[2];
        c = (sub & 2) ? hi : lo;
    } else if constexpr (EPL == 2) {
        c = (pos & 1) ? k[1] : k[0];
    } else {
        c = k[0];
    }
    return __shfl_sync(MLB_FULL, c, pos / EPL);
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

// Features of one reservoir held in registers: v/t are this lane's EPL slots
// (slot index lane*EPL + r), n = number of valid slots (>= 1).
template <int EPL>
__device__ __forceinline__ void warp_features_regs(float (&v)[EPL], float (&t)[EPL], int n, float now,
                                                   double decay, float log2_decay, float (&out)[5]) {
    const int lane = lane_id();
    // ---- mean / std (np.mean, np.std on float32: reservoir.py:143,145)
    float s = 0.f, tmax = -MLB_INF;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = lane * EPL + r < n;
        s += valid ? v[r] : 0.f;
        tmax = valid ? fmaxf(tmax, t[r]) : tmax;
    }
    const float nf = (float)n;
    const float mean = warp_sum(s) / nf;
    tmax = warp_max(tmax);
    float s2 = 0.f;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = lane * EPL + r < n;
        const float d = v[r] - mean;
        s2 += valid ? d * d : 0.f;
        if (!valid) v[r] = MLB_INF;  // padding sorts to the end
    }
    const float sd = sqrtf(warp_sum(s2) / nf);

    // ---- sort by value, timestamp rides along
    bitonic_sort_kv<EPL>(v, t, lane);

    // ---- p90 = np.percentile(values, 90): float32 'linear' rule of numpy >= 2
    const float vidx = (float)(n - 1) * (90.0f / 100.0f);
    const float fl = floorf(vidx);
    int lo = (int)fl, hi = lo + 1;
    if (vidx >= (float)(n - 1)) { lo = n - 1; hi = n - 1; }
    hi = hi > n - 1 ? n - 1 : hi;
    const float gamma = vidx - fl;
    const float a = sorted_at<EPL>(v, lo), b = sorted_at<EPL>(v, hi);
    const float diff = __fsub_rn(b, a);
    float p90 = __fadd_rn(a, __fmul_rn(diff, gamma));
    if (gamma >= 0.5f) p90 = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));

    // ---- decay weights relative to the newest sample (float32 fast path)
    float w[EPL], cum[EPL];
    float sw = 0.f, svw = 0.f;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = lane * EPL + r < n;
        const float x = log2_decay * (tmax - t[r]);
        w[r] = valid ? exp2f(x) : 0.f;
        sw += w[r];
        svw += valid ? v[r] * w[r] : 0.f;
        cum[r] = sw;
    }
    const float incl = warp_scan_incl(sw, lane);
    const float W = __shfl_sync(MLB_FULL, incl, 31);
    const float excl = incl - sw;
    const float mean_decay = warp_sum(svw) / W;
    const float cutoff = 0.9f * W;
    int below = 0;
    float dmin = MLB_INF;
#pragma unroll
    for (int r = 0; r < EPL; r++) {
        const bool valid = lane * EPL + r < n;
        const float c = excl + cum[r];
        below += (valid && c < cutoff) ? 1 : 0;
        dmin = valid ? fminf(dmin, fabsf(c - cutoff)) : dmin;
    }
    int idx = __reduce_add_sync(MLB_FULL, below);
    const uint32_t dmin_bits = __reduce_min_sync(MLB_FULL, __float_as_uint(dmin));
    if (__uint_as_float(dmin_bits) < MLB_WP_MARGIN * W) {
        // ---- faithful float64 decision (reservoir.py:148-149,181-196)
        double wd[EPL], cd[EPL];
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const bool valid = lane * EPL + r < n;
            wd[r] = valid ? pow(decay, (double)now - (double)t[r]) : 0.0;
            cd[r] = 0.0;
        }
        double c = 0.0;
        for (int L = 0; L * EPL < n; L++) {
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                c += __shfl_sync(MLB_FULL, wd[r], L);  // padding adds 0.0
                if (lane == L) cd[r] = c;
            }
        }
        const double cut = 0.9 * c;  // percentile * cumsum[-1]
        below = 0;
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const bool valid = lane * EPL + r < n;
            below += (valid && cd[r] < cut) ? 1 : 0;
        }
        idx = __reduce_add_sync(MLB_FULL, below);
    }
    idx = idx > n - 1 ? n - 1 : idx;  // reservoir.py:193-194
    const float p90_decay = sorted_at<EPL>(v, idx);
    out[0] = mean;
    out[1] = p90;
    out[2] = sd;
    out[3] = mean_decay;
    out[4] = p90_decay;
}

// Load + compute for one reservoir in global memory.  n = min(count, K), may be 0.
__device__ __forceinline__ void warp_features(const float* __restrict__ vals,
                                              const float* __restrict__ tss, int n, float now,
                                              double decay, float log2_decay, float (&out)[5]) {
    const int lane = lane_id();
    if (n <= 0) {  // reservoir.py:127-134
#pragma unroll
        for (int q = 0; q < 5; q++) out[q] = 0.f;
    } else if (n <= 32) {
        float v[1], t[1];
        load_slots<1>(vals, lane, v);
        load_slots<1>(tss, lane, t);
        warp_features_regs<1>(v, t, n, now, decay, log2_decay, out);
    } else if (n <= 64) {
        float v[2], t[2];
        load_slots<2>(vals, lane, v);
        load_slots<2>(tss, lane, t);
        warp_features_regs<2>(v, t, n, now, decay, log2_decay, out);
    } else {
        float v[4], t[4];
        load_slots<4>(vals, lane, v);
        load_slots<4>(tss, lane, t);
        warp_features_regs<4>(v, t, n, now, decay, log2_decay, out);
    }
}

}  // namespace mlb
