// Stand-alone batched pieces of the hot path behind the C ABI:
//   ReservoirSampler.add / get_features (problem-01-reservoir-sampling/src/reservoir.py)
//   fairness metrics (problem-03-rl-environment/src/rewards.py)
//   legacy random observation (problem-03-rl-environment/src/env.py:425-448)
#include <cuda_runtime.h>

#include "mlb_step_kernel.cuh"

using namespace mlb;

// One thread per reservoir: adds are sequential within a reservoir by definition.
__global__ void reservoir_add_kernel(float* __restrict__ values, float* __restrict__ ts,
                                     uint32_t* __restrict__ count, uint32_t* __restrict__ cursor,
                                     const uint32_t* __restrict__ mt_table,
                                     const int32_t* __restrict__ seed_row, int L, int R, int K, int KP,
                                     const float* __restrict__ add_v, const float* __restrict__ add_t,
                                     const int32_t* __restrict__ n_add, int max_add,
                                     uint8_t* __restrict__ accepted, int* status) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    uint32_t cnt = count[r], cur = cursor[r];
    const uint32_t* row = mt_table + (size_t)seed_row[r] * L;
    const int n = n_add[r];
    for (int i = 0; i < n; i++) {
        const int slot = res_draw_slot(cnt, cur, row, L, K, status);
        cnt++;
        if (slot >= 0) {
            values[(size_t)r * KP + slot] = add_v[(size_t)r * max_add + i];
            ts[(size_t)r * KP + slot] = add_t[(size_t)r * max_add + i];
        }
        if (accepted) accepted[(size_t)r * max_add + i] = slot >= 0;
    }
    count[r] = cnt;
    cursor[r] = cur;
}

// One warp per reservoir.
__global__ void reservoir_features_kernel(const float* __restrict__ values, const float* __restrict__ ts,
                                          const uint32_t* __restrict__ count, int R, int K, int KP,
                                          double decay, float log2_decay, const float* __restrict__ now,
                                          float* __restrict__ out) {
    __shared__ __align__(16) float scratch[4][256];
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const uint32_t c = count[r];
    const int n = c < (uint32_t)K ? (int)c : K;
    const float mine = warp_features_sorted(values + (size_t)r * KP, ts + (size_t)r * KP, nullptr, n, now[r], decay,
                                            log2_decay, reinterpret_cast<float2*>(scratch[threadIdx.x >> 5]));
    if (lane < 5) out[(size_t)r * 5 + lane] = mine;
}

// One warp per row of values.
__global__ void reward_metric_kernel(int metric, const double* __restrict__ values,
                                     const int32_t* __restrict__ n, int B, int stride,
                                     double* __restrict__ out) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    double r;
    if (n[b] == 0) {
        // empty-list conventions of the bare metric functions (rewards.py:49-50,91-92,...)
        r = (metric == MLB_REWARD_JAIN || metric == MLB_REWARD_FAIR_JAIN) ? 1.0 : 0.0;
    } else {
        r = reward_staged<double>(metric, values + (size_t)b * stride, nullptr, n[b]);
    }
    if ((threadIdx.x & 31) == 0) out[b] = r;
}

// env.py:461-468, elementwise in float64.  No fused multiply-adds: numpy rounds every operation.
__global__ void normalize_obs_kernel(const float* __restrict__ obs, double* __restrict__ mean,
                                     double* __restrict__ sd, double count, double* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = (double)obs[i];
        const double m0 = mean[i], s0 = sd[i];
        const double delta = __dsub_rn(x, m0);                                   // :462
        const double m1 = __dadd_rn(m0, __ddiv_rn(delta, count));                // :463
        const double delta2 = __dsub_rn(x, m1);                                  // :464
        const double num = __dadd_rn(__dmul_rn(__dmul_rn(s0, s0), count - 1.0), __dmul_rn(delta, delta2));
        const double s1 = __dsqrt_rn(fmax(__ddiv_rn(num, count), 1e-8));         // :465
        mean[i] = m1;
        sd[i] = s1;
        out[i] = __ddiv_rn(__dsub_rn(x, m1), __dadd_rn(s1, 1e-8));               // :468
    }
}

// ---- device MT19937, state [E][625] (624 words + position) -------------------
__device__ __forceinline__ uint32_t mt_next(uint32_t* st) {
    uint32_t pos = st[624];
    if (pos >= 624) {
        for (int k = 0; k < 624; k++) {
            const uint32_t y = (st[k] & 0x80000000u) | (st[(k + 1) % 624] & 0x7fffffffu);
            st[k] = st[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        pos = 0;
    }
    uint32_t y = st[pos];
    st[624] = pos + 1;
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__device__ __forceinline__ double mt_double(uint32_t* st) {  // random_sample
    const uint32_t a = mt_next(st) >> 5, b = mt_next(st) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

__global__ void legacy_seed_kernel(uint32_t* __restrict__ state, const uint32_t* __restrict__ seeds, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    uint32_t* st = state + (size_t)e * 625;
    uint32_t x = seeds[e];
    st[0] = x;
    for (int i = 1; i < 624; i++) {
        x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
        st[i] = x;
    }
    st[624] = 624;
}

// env.py:425-448: per server randint(5,20) then six uniform() draws; the derived
// columns are float32 products (numpy >= 2 semantics, see oracle/flow_oracle.c).
__global__ void legacy_obs_kernel(uint32_t* __restrict__ state, int E, int S, float* __restrict__ obs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    uint32_t* st = state + (size_t)e * 625;
    for (int s = 0; s < S; s++) {
        float* o = obs + ((size_t)e * S + s) * MLB_OBS_COLS;
        uint32_t v;
        do { v = mt_next(st) & 15u; } while (v > 14u);          // randint(5, 20)
        const float c0 = (float)(5 + (int)v);
        // uniform(lo, hi) = lo + (hi - lo) * random_sample(), no FMA contraction
        const float c1 = (float)__dadd_rn(5.0, __dmul_rn(15.0 - 5.0, mt_double(st)));
        const float c2 = (float)__dadd_rn(10.0, __dmul_rn(25.0 - 10.0, mt_double(st)));
        const float c3 = (float)__dadd_rn(1.0, __dmul_rn(5.0 - 1.0, mt_double(st)));
        const float c6 = (float)__dadd_rn(8.0, __dmul_rn(18.0 - 8.0, mt_double(st)));
        const float c7 = (float)__dadd_rn(15.0, __dmul_rn(30.0 - 15.0, mt_double(st)));
        const float c8 = (float)__dadd_rn(2.0, __dmul_rn(8.0 - 2.0, mt_double(st)));
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
        o[4] = __fmul_rn(c1, 0.9f);
        o[5] = __fmul_rn(c2, 0.9f);
        o[6] = c6; o[7] = c7; o[8] = c8;
        o[9] = __fmul_rn(c6, 0.85f);
        o[10] = __fmul_rn(c6, 0.9f);
    }
}

// ---- warp-per-env legacy step -----------------------------------------------------------------------------
// The reference's whole simulation-mode step is `_simulate_observation` + the reward (env.py:254-262, 425-448):
// per server one randint(5, 20) (masked rejection: 1 + #rejected words) and six uniform() draws (2 words each),
// in server order, from ONE MT19937 stream per env.  One warp per env:
//   * the state row [624 words | position] of the env is contiguous, so every access is a coalesced 128-byte line
//     (keeping the row in shared memory for the step was measured slower: 9 KB per warp leaves 24 warps per SM);
//     the twist runs in three phases of up to 224 words (word k needs words k, k+1 and k+397 of the previous round or k-227 of this one:
//     inside 224 consecutive words nothing depends on a word written in the same phase once all reads precede
//     the writes), at the moment numpy would run it (position 624), so state and position stay numpy's own
//     representation;
//   * tempered words are staged in shared memory in consumption order, on demand and clipped at the end of the
//     round, so a step never twists further than numpy would have;
//   * where a server's words start depends on every rejection before it: a fix-point over the lanes (guess, count
//     the rejected words at the guess, exclusive scan) settles all 32 starts in two or three rounds -- rejections
//     are rare, one trial in 16 --, then lane s evaluates server s in the reference's float64 /
//     float32 arithmetic; rows leave through shared memory as coalesced stores; the reward is reward_staged.
#define MLB_LG_STAGE 1024           // staging ring (words), power of two, >= 13 * 32 + rejections + one chunk
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// in-place twist of a whole state row by one warp (numpy's mt19937_gen), 224 words per phase
__device__ __forceinline__ void mt_twist_warp(uint32_t* st, int lane) {
#pragma unroll 1
    for (int base = 0; base < 624; base += 224) {
        uint32_t nw[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const int k = base + lane + 32 * j;
            nw[j] = 0;
            if (k < 624) {
                const uint32_t a = st[k], b = st[k + 1 == 624 ? 0 : k + 1], c = st[k + 397 >= 624 ? k - 227 : k + 397];
                const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
                nw[j] = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
        }
        __syncwarp();                               // every read of this phase precedes its writes
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const int k = base + lane + 32 * j;
            if (k < 624) st[k] = nw[j];
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128)
legacy_step_kernel(uint32_t* __restrict__ state, int E, int S, int metric, int field, float* __restrict__ obs,
                   double* __restrict__ reward, int* __restrict__ status) {
    __shared__ uint32_t s_stage[4][MLB_LG_STAGE];
    __shared__ float s_rows[4][32 * MLB_OBS_COLS];
    __shared__ float s_rv[4][256];                 // reward-field value of every server
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * 4 + warp;
    if (e >= E) return;
    uint32_t* const st = state + (size_t)e * 625;  // stays in global memory: L2-resident for the step
    if (lane < 20) prefetch_l2(st + lane * 32);    // the whole row is needed within this step (staging, then the twist)
    uint32_t* stage = s_stage[warp];
    float* rows = s_rows[warp];
    int pos = (int)st[624];                         // numpy's `pos`: next unread word of the current round
    uint32_t staged = 0, used = 0;                  // words staged / consumed so far in this step (ring counters)
    float* rv = s_rv[warp];
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int nb = S - s0 < 32 ? S - s0 : 32;
        // ---- where does each server's draw start?  Server s starts 13 words after server s-1 plus every word
        // randint(5, 20) rejected on the way (masked rejection, env.py:436).  Fix-point over the lanes:
        // lane s guesses its start from the rejections R_s counted so far before it, counts the rejected words k_s at
        // that guess, and an exclusive scan of k gives the next R.  After i rounds the first i servers are exact, and
        // with one rejection in 16 trials two or three rounds settle all 32.
        int done = nb;
        uint32_t cur = used;
        uint32_t my_o = used;                       // accepted randint word of this lane's server
        // the batch consumes at least 13 words per server: stage those in one go (never more than numpy would draw)
        for (int needw = 13 * nb - (int)(staged - used); needw > 0;) {
            if (pos == 624) {
                mt_twist_warp(st, lane);
                pos = 0;
            }
            const int take = 624 - pos < needw ? 624 - pos : needw;
            for (int i = lane; i < take; i += 32) stage[(staged + i) & (MLB_LG_STAGE - 1)] = mt_temper(st[pos + i]);
            staged += take;
            pos += take;
            needw -= take;
        }
        __syncwarp();
        for (;;) {
            uint32_t R = 0;
            bool isshort = false;
            for (int round = 0; round <= nb; round++) {
                uint32_t k = 0;
                isshort = false;
                if (lane < nb) {
                    const uint32_t o = used + 13u * (uint32_t)lane + R;
                    while (o + k < staged && (stage[(o + k) & (MLB_LG_STAGE - 1)] & 15u) == 15u) k++;
                    isshort = o + k + 13u > staged;          // words not staged yet count as "not rejected" for now
                    my_o = o + k;
                }
                uint32_t incl = k;
#pragma unroll
                for (int d2 = 1; d2 < 32; d2 <<= 1) {
                    const uint32_t t = __shfl_up_sync(MLB_FULL, incl, d2);
                    if (lane >= d2) incl += t;
                }
                const uint32_t newR = incl - k;
                const bool same = newR == R;
                R = newR;
                cur = __shfl_sync(MLB_FULL, used + 13u * (uint32_t)nb + incl, nb - 1);
                if (__all_sync(MLB_FULL, same)) break;
            }
            // at the fix-point the lowest short lane has an exact start, so it really needs more words
            if (!__any_sync(MLB_FULL, isshort)) break;
            if (staged - used > MLB_LG_STAGE - 64) {          // > 500 rejections in one batch: not a real stream
                if (lane == 0) atomicOr(status, ST_ERR_RNG);
                done = 0;
                break;
            }
            if (pos == 624) {                       // numpy twists exactly here
                mt_twist_warp(st, lane);
                pos = 0;
            }
            const int take = 624 - pos < 32 ? 624 - pos : 32;
            if (lane < take) stage[(staged + lane) & (MLB_LG_STAGE - 1)] = mt_temper(st[pos + lane]);
            staged += take;
            pos += take;
            __syncwarp();
        }
        __syncwarp();
        // ---- lane s: server s0 + s (env.py:436-446); uniform(lo, hi) = lo + (hi - lo) * random_sample()
        if (lane < done) {
            const uint32_t o = my_o;
            auto word = [&](int i) { return stage[(o + i) & (MLB_LG_STAGE - 1)]; };
            auto rs = [&](int i) {                  // random_sample from words i, i + 1
                const uint32_t a = word(i) >> 5, b = word(i + 1) >> 6;
                return (a * 67108864.0 + b) / 9007199254740992.0;
            };
            const float c0 = (float)(5 + (int)(word(0) & 15u));
            const float c1 = (float)__dadd_rn(5.0, __dmul_rn(15.0 - 5.0, rs(1)));
            const float c2 = (float)__dadd_rn(10.0, __dmul_rn(25.0 - 10.0, rs(3)));
            const float c3 = (float)__dadd_rn(1.0, __dmul_rn(5.0 - 1.0, rs(5)));
            const float c6 = (float)__dadd_rn(8.0, __dmul_rn(18.0 - 8.0, rs(7)));
            const float c7 = (float)__dadd_rn(15.0, __dmul_rn(30.0 - 15.0, rs(9)));
            const float c8 = (float)__dadd_rn(2.0, __dmul_rn(8.0 - 2.0, rs(11)));
            float* o11 = rows + lane * MLB_OBS_COLS;
            o11[0] = c0; o11[1] = c1; o11[2] = c2; o11[3] = c3;
            o11[4] = __fmul_rn(c1, 0.9f);           // derived columns: float32 products (numpy >= 2), env.py:440-446
            o11[5] = __fmul_rn(c2, 0.9f);
            o11[6] = c6; o11[7] = c7; o11[8] = c8;
            o11[9] = __fmul_rn(c6, 0.85f);
            o11[10] = __fmul_rn(c6, 0.9f);
            rv[s0 + lane] = o11[field];
        }
        __syncwarp();
        float* dst = obs + ((size_t)e * S + s0) * MLB_OBS_COLS;
        for (int i = lane; i < done * MLB_OBS_COLS; i += 32) dst[i] = rows[i];
        used = cur;
        __syncwarp();
    }
    // unread staged words all belong to the current round (staging is on demand and clipped at the round end)
    if (lane == 0) st[624] = (uint32_t)(pos - (int)(staged - used));
    if (reward) {
        // every server is active (n_flow_on >= 5 > 0, env.py:410-417): the metric runs over all S values
        __syncwarp();
        const double r = reward_staged<float>(metric, rv, nullptr, S);
        if (lane == 0) reward[e] = r;
    }
}

#define CKL()                                          \
    do {                                               \
        if (cudaGetLastError() != cudaSuccess) return MLB_ECUDA; \
    } while (0)

extern "C" {

int mlb_reservoir_add(float* values, float* ts, uint32_t* count, uint32_t* cursor,
                      const uint32_t* mt_table, const int32_t* seed_row, int32_t table_len,
                      int32_t R, int32_t K, const float* add_v, const float* add_t,
                      const int32_t* n_add, int32_t max_add, uint8_t* accepted, int32_t* status,
                      void* stream) {
    if (!values || !ts || !count || !cursor || !mt_table || !seed_row || !add_v || !add_t || !n_add || !status)
        return MLB_EINVAL;
    if (K < 1 || K > 128 || R < 0) return MLB_EINVAL;
    if (R == 0) return MLB_OK;
    const int KP = (K + 31) & ~31;
    reservoir_add_kernel<<<(R + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        values, ts, count, cursor, mt_table, seed_row, table_len, R, K, KP, add_v, add_t, n_add,
        max_add, accepted, status);
    CKL();
    return MLB_OK;
}

int mlb_reservoir_features(const float* values, const float* ts, const uint32_t* count, int32_t R,
                           int32_t K, double decay, const float* now, float* out, void* stream) {
    if (!values || !ts || !count || !now || !out) return MLB_EINVAL;
    if (K < 1 || K > 128 || R < 0 || !(decay > 0)) return MLB_EINVAL;
    if (R == 0) return MLB_OK;
    const int KP = (K + 31) & ~31;
    const int wpb = 4;
    reservoir_features_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        values, ts, count, R, K, KP, decay, (float)log2(decay), now, out);
    CKL();
    return MLB_OK;
}

int mlb_reward_metric(int metric, const double* values, const int32_t* n, int32_t B, int32_t stride,
                      double* out, void* stream) {
    if (!values || !n || !out) return MLB_EINVAL;
    if (metric < 0 || metric >= MLB_REWARD_COUNT_) return MLB_EINVAL;
    if (B == 0) return MLB_OK;
    const int wpb = 4;
    reward_metric_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(metric, values, n, B, stride, out);
    CKL();
    return MLB_OK;
}

int mlb_normalize_obs(const float* obs, double* mean, double* std, int64_t count, double* out, int64_t n,
                      void* stream) {
    if (!obs || !mean || !std || !out || count < 1 || n < 0) return MLB_EINVAL;
    if (n == 0) return MLB_OK;
    const int threads = 256;
    const int64_t want = (n + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);                 // grid-stride, 16 blocks per SM
    normalize_obs_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(obs, mean, std, (double)count, out, n);
    CKL();
    return MLB_OK;
}

int mlb_legacy_seed(uint32_t* mt_state, const uint32_t* seeds, int32_t E, void* stream) {
    if (!mt_state || !seeds || E < 0) return MLB_EINVAL;
    if (E == 0) return MLB_OK;
    legacy_seed_kernel<<<(E + 63) / 64, 64, 0, (cudaStream_t)stream>>>(mt_state, seeds, E);
    CKL();
    return MLB_OK;
}

int mlb_legacy_obs(uint32_t* mt_state, int32_t E, int32_t S, float* obs, void* stream) {
    if (!mt_state || !obs || E < 0 || S < 1) return MLB_EINVAL;
    if (E == 0) return MLB_OK;
    legacy_obs_kernel<<<(E + 63) / 64, 64, 0, (cudaStream_t)stream>>>(mt_state, E, S, obs);
    CKL();
    return MLB_OK;
}

int mlb_legacy_step(uint32_t* mt_state, int32_t E, int32_t S, int32_t metric, int32_t field, float* obs,
                    double* reward, int32_t* status, void* stream) {
    if (!mt_state || !obs || !status || E < 0 || S < 1 || S > 256) return MLB_EINVAL;
    if (reward && (metric < 0 || metric >= MLB_REWARD_COUNT_ || field < 0 || field >= MLB_OBS_COLS)) return MLB_EINVAL;
    if (E == 0) return MLB_OK;
    legacy_step_kernel<<<(E + 3) / 4, 128, 0, (cudaStream_t)stream>>>(mt_state, E, S, metric, field, obs, reward, status);
    CKL();
    return MLB_OK;
}

}  // extern "C"
