// The fused env-step kernel: one warp per (env, LB agent).
//
// Stands in for LoadBalanceEnv.step (reference:
// simulation-mode/problem-03-rl-environment/src/env.py:215-286) executed for
// E envs at once, with the flow-level dynamics of SURVEY.md App. B:
//   phase 0  action -> weights (env.py:334-353); per-server scalars -> shared memory
//   phase A  per arrival (time order): retire finished flows (n_flow_on--, fct
//            sample -> Algorithm-R add, reservoir.py:50-85), choose a server
//            (SED / LSQ / alias, src/vpp/lb/node.c:393-460) with REDUX argmin,
//            push on that server's FIFO ring
//   phase C  window end: retire, then one flow_duration sample per active flow
//   phase B  reservoir statistics of every reservoir touched this step
//            (reservoir.py:105-196) -> obs columns (features.py:256-286)
//   phase R  reward over the env's active servers (rewards.py:329-381), done flag
//
// Every per-server scalar is owned by lane (server % 32) of the agent's warp, so
// phases A/C need no intra-warp locking; phase B is warp-cooperative per
// reservoir with warp-uniform control flow.
#pragma once
#include "mlb_common.cuh"
#include "mlb_env.cuh"
#include "mlb_features.cuh"
#include "../../include/marllb_b200.h"

namespace mlb {

constexpr int NF = 23;  // per-server shared-memory fields (4 bytes each)

struct WarpSmem {
    int32_t* n_on;
    float* last_fin;
    uint32_t* head;
    float* head_fin;
    uint32_t* score;
    float* speed;
    float* weight;
    uint32_t* cnt[2];
    uint32_t* cur[2];
    uint32_t* dropped;
    uint32_t* flags;
    uint32_t* nold[2];    // valid slots at step start (what the stored ranks describe)
    uint32_t* chg;        // [2][4][SP] bit mask of reservoir slots written this step
    int sp;
    // alias policy only
    double* prob;
    int32_t* alias;
    int32_t* stack;
};

__host__ __device__ inline size_t warp_smem_bytes(int SP, bool alias) {
    return (size_t)NF * SP * 4 + (alias ? (size_t)SP * (8 + 4 + 8) : 0);
}

__device__ __forceinline__ WarpSmem carve(unsigned char* base, int SP, bool alias) {
    WarpSmem w;
    unsigned char* p = base;
    if (alias) {  // doubles first: keep 8-byte alignment
        w.prob = reinterpret_cast<double*>(p);
        p += (size_t)SP * 8;
    } else {
        w.prob = nullptr;
    }
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    w.n_on = reinterpret_cast<int32_t*>(q + 0 * SP);
    w.last_fin = reinterpret_cast<float*>(q + 1 * SP);
    w.head = q + 2 * SP;
    w.head_fin = reinterpret_cast<float*>(q + 3 * SP);
    w.score = q + 4 * SP;
    w.speed = reinterpret_cast<float*>(q + 5 * SP);
    w.weight = reinterpret_cast<float*>(q + 6 * SP);
    w.cnt[0] = q + 7 * SP;
    w.cnt[1] = q + 8 * SP;
    w.cur[0] = q + 9 * SP;
    w.cur[1] = q + 10 * SP;
    w.dropped = q + 11 * SP;
    w.flags = q + 12 * SP;
    w.nold[0] = q + 13 * SP;
    w.nold[1] = q + 14 * SP;
    w.chg = q + 15 * SP;
    w.sp = SP;
    w.alias = alias ? reinterpret_cast<int32_t*>(q + NF * SP) : nullptr;
    w.stack = alias ? reinterpret_cast<int32_t*>(q + (NF + 1) * SP) : nullptr;
    return w;
}

// score of one server under SED / LSQ as an order-preserving uint
// node.c:395-404: f32 score = (n_flow_on + 1) / (1e-9 + weight), evaluated in double
template <int POLICY>
__device__ __forceinline__ uint32_t server_score(int n, float w) {
    if (POLICY == MLB_POLICY_SED) {
        const double s = (double)(n + 1) / (1e-9 + (double)w);
        return f32_orderable((float)s);
    } else {
        return f32_orderable((float)n);  // node.c:419-431
    }
}

// Index draw of ReservoirSampler.add (reservoir.py:65-85): returns the slot the
// sample goes to, or -1 when Algorithm R rejects it.  Replays
// RandomState(seed).randint(0, count+1): masked rejection over the raw MT19937
// words of that seed's row, `cur` = words consumed so far.
__device__ __forceinline__ int res_draw_slot(uint32_t cnt, uint32_t& cur, const uint32_t* __restrict__ row,
                                             int L, int K, int* status) {
    if (cnt < (uint32_t)K) return (int)cnt;  // fill phase, reservoir.py:65-73
    const uint32_t mask = 0xffffffffu >> __clz(cnt);
    uint32_t c = cur, v;
    do {
        if (c >= (uint32_t)L) {
            atomicOr(status, ST_ERR_RNG);
            v = 0xffffffffu;
            break;
        }
        v = __ldg(row + c) & mask;
        c++;
    } while (v > cnt);
    cur = c;
    return v < (uint32_t)K ? (int)v : -1;  // reservoir.py:78-85
}

// ReservoirSampler.add by the owning lane of server j, metric m.
__device__ __forceinline__ void res_add(const DevState& d, const WarpSmem& s, int m, int j,
                                        size_t srv, int seed_row, float value, float ts) {
    const uint32_t cnt = s.cnt[m][j];
    uint32_t c = s.cur[m][j];
    const int slot = res_draw_slot(cnt, c, d.mt_table + (size_t)seed_row * d.L, d.L, d.K, d.status);
    s.cur[m][j] = c;
    s.cnt[m][j] = cnt + 1;
    if (slot >= 0) {
        const size_t at = (srv * 2 + m) * d.KP + slot;
        d.res_val[at] = value;
        d.res_ts[at] = ts;
        s.flags[j] |= (1u << m);
        s.chg[(m * 4 + (slot >> 5)) * s.sp + j] |= 1u << (slot & 31);
    }
}

template <int POLICY>
__device__ __forceinline__ void retire(const DevState& d, const WarpSmem& s, int j, size_t srv,
                                       int seed_row, float now) {
    // pop while the oldest flow finished strictly before `now`
    while (s.head_fin[j] < now) {
        uint32_t h = s.head[j];
        const size_t rb = srv * d.Q;
        const float arr = d.ring_arr[rb + h];
        const float fin = s.head_fin[j];
        h = (h + 1 == (uint32_t)d.Q) ? 0u : h + 1;
        s.head[j] = h;
        const int n = s.n_on[j] - 1;  // src/vpp/lb/lbhash.h:120
        s.n_on[j] = n;
        s.head_fin[j] = n > 0 ? d.ring_fin[rb + h] : MLB_INF;
        res_add(d, s, 0, j, srv, seed_row, __fsub_rn(fin, arr), fin);  // lbhash.h:122-124
        if (POLICY != MLB_POLICY_ALIAS) s.score[j] = server_score<POLICY>(n, s.weight[j]);
    }
}

// numpy pairwise block sum (n <= 128) and its two-block extension (n <= 256)
__device__ inline double np_block_sum(const double* a, int n) {
    if (n < 8) {
        double r = -0.0;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    double r[8];
    for (int k = 0; k < 8; k++) r[k] = a[k];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; k++) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}
__device__ inline double np_sum_f64(const double* a, int n) {
    if (n <= 128) return np_block_sum(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_block_sum(a, n2) + np_block_sum(a + n2, n - n2);
}

// rl_controller.py:359-405 _build_alias_table over p = w / sum(w); one lane.
__device__ inline void alias_build(const WarpSmem& s, int Sa) {
    for (int k = 0; k < Sa; k++) s.prob[k] = (double)s.weight[k];
    const double tot = np_sum_f64(s.prob, Sa);
    int* small = s.stack;
    int* large = s.stack + Sa;
    int ns = 0, nl = 0;
    for (int k = 0; k < Sa; k++) {
        const double p = (s.prob[k] / tot) * (double)Sa;
        s.prob[k] = p;
        s.alias[k] = k;
        if (p < 1.0) small[ns++] = k; else large[nl++] = k;
    }
    while (ns > 0 && nl > 0) {
        const int l = small[--ns];
        const int g = large[--nl];
        s.alias[l] = g;
        const double pg = s.prob[g] + s.prob[l] - 1.0;
        s.prob[g] = pg;
        if (pg < 1.0) small[ns++] = g; else large[nl++] = g;
    }
}

// Reward metric over the staged reward-field values of one env (leader warp).
// rewards.py:21-287; float64 throughout like the reference.
template <typename T>
__device__ inline double reward_staged(int metric, const T* rv, const uint32_t* ra, int S) {
    const int lane = lane_id();
    const double eps = 1e-10;
    int na = 0;
    double sum = 0.0, sumsq = 0.0, mx = -1.0e308, mn = 1.0e308, slog = 0.0;
    for (int i = lane; i < S; i += 32) {
        if (!ra || ra[i]) {
            const double x = (double)rv[i];
            na++;
            sum += x;
            sumsq += x * x;
            mx = fmax(mx, x);
            mn = fmin(mn, x);
            if (metric == MLB_REWARD_PRODUCT) slog += log(x + eps);
        }
    }
    na = __reduce_add_sync(MLB_FULL, na);
    if (na == 0) return 0.0;  // rewards.py:364-365
    const double n = (double)na;
    sum = warp_sum(sum);
    switch (metric) {
    case MLB_REWARD_JAIN: {
        if (sum < eps) return 1.0;
        sumsq = warp_sum(sumsq);
        if (sumsq < eps) return 1.0;
        const double j = (sum * sum) / (n * sumsq);
        return fmin(fmax(j, 1.0 / n), 1.0);
    }
    case MLB_REWARD_MAX: return -warp_max(mx);
    case MLB_REWARD_MIN: return warp_min(mn);
    case MLB_REWARD_RANGE: return -(warp_max(mx) - warp_min(mn));
    case MLB_REWARD_PRODUCT: return warp_sum(slog);
    default: break;
    }
    const double mean = sum / n;
    if (metric == MLB_REWARD_GINI) {
        if (mean == 0.0) return 0.0;
        double ds = 0.0;
        for (int i = lane; i < S; i += 32) {
            if (ra && !ra[i]) continue;
            const double xi = (double)rv[i];
            for (int k = 0; k < S; k++)
                if (!ra || ra[k]) ds += fabs(xi - (double)rv[k]);
        }
        ds = warp_sum(ds);
        return -(ds / (2.0 * n * n * mean));
    }
    double ss = 0.0;  // two-pass variance like np.var
    for (int i = lane; i < S; i += 32) {
        if (!ra || ra[i]) {
            const double dlt = (double)rv[i] - mean;
            ss += dlt * dlt;
        }
    }
    const double var = warp_sum(ss) / n;
    if (metric == MLB_REWARD_VARIANCE) return -var;
    if (metric == MLB_REWARD_STD) return -sqrt(var);
    // MLB_REWARD_CV
    if (mean < eps) return 0.0;
    return -(sqrt(var) / (mean + eps));
}

// SINGLE = one agent per env (A == 1): blocks are exactly 4 warps = 4 envs and no
// inter-warp barrier is needed, so the kernel can be register-capped for occupancy.
template <int POLICY, bool SINGLE>
__global__ void __launch_bounds__(SINGLE ? 128 : 1024, SINGLE ? 8 : 1)
step_kernel(const __grid_constant__ DevState d, const void* __restrict__ action) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int A = SINGLE ? 1 : d.A, Sa = d.Sa, S = d.S;
    const int nwarps = blockDim.x >> 5;
    const int epb = nwarps / A;
    const int env_in_blk = warp / A;
    const int agent = warp - env_in_blk * A;
    const int e = blockIdx.x * epb + env_in_blk;
    if (e >= d.E) return;  // whole env (all its warps) leaves together

    const int SP = (Sa + 31) & ~31;
    constexpr bool kAlias = (POLICY == MLB_POLICY_ALIAS);
    const size_t wbytes = warp_smem_bytes(SP, kAlias);
    const WarpSmem s = carve(smem_raw + (size_t)warp * wbytes, SP, kAlias);
    // per-env reward staging after all warp areas
    float* rv = reinterpret_cast<float*>(smem_raw + (size_t)nwarps * wbytes) + (size_t)env_in_blk * 2 * S;
    uint32_t* ra = reinterpret_cast<uint32_t*>(rv + S);
    // per-warp 1 KB scratch for the rank-ordered statistics, after the reward staging
    float* scr = reinterpret_cast<float*>(smem_raw + (size_t)nwarps * wbytes) + (size_t)epb * 2 * S + (size_t)warp * 256;
    const WarpScratch scratch{scr, scr + 128};

    const int step = d.step[e] + 1;                       // env.py:230
    const float t1 = __fmul_rn((float)step, d.dt);        // window end
    const size_t sbase = (size_t)e * S + (size_t)agent * Sa;
    const int seed0 = agent * Sa;                         // replay row = server index in env

    // ---------------- phase 0: load state, action -> weights -----------------
    for (int j = lane; j < Sa; j += 32) {
        const size_t g = sbase + j;
        const int n = d.n_on[g];
        const uint32_t h = d.head[g];
        s.n_on[j] = n;
        s.last_fin[j] = d.last_fin[g];
        s.head[j] = h;
        s.head_fin[j] = n > 0 ? d.ring_fin[g * d.Q + h] : MLB_INF;
        s.speed[j] = d.speed[g];
        s.dropped[j] = d.dropped[g];
        s.flags[j] = 0;
        float w;
        if (d.action_kind == MLB_ACTION_CONTINUOUS_F32) {
            const float x = reinterpret_cast<const float*>(action)[g];
            w = fminf(fmaxf(x, d.min_w), d.max_w);        // env.py:349-351
        } else {
            int a = d.action_kind == MLB_ACTION_DISCRETE_U8
                        ? (int)reinterpret_cast<const uint8_t*>(action)[g]
                        : reinterpret_cast<const int32_t*>(action)[g];
            if ((unsigned)a >= (unsigned)d.n_discrete) {
                atomicOr(d.status, ST_ERR_ACTION);
                a = 0;
            }
            w = d.dw[a];                                  // env.py:346
        }
        s.weight[j] = w;
        if (!kAlias) s.score[j] = server_score<POLICY>(n, w);
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const size_t c = ((size_t)e * 2 + m) * S + seed0 + j;
            const uint32_t cnt = d.res_count[c];
            s.cnt[m][j] = cnt;
            s.cur[m][j] = d.res_cursor[c];
            s.nold[m][j] = cnt < (uint32_t)d.K ? cnt : (uint32_t)d.K;
#pragma unroll
            for (int k = 0; k < 4; k++) s.chg[(m * 4 + k) * s.sp + j] = 0;
        }
    }
    __syncwarp();
    if (kAlias) {
        if (lane == 0) alias_build(s, Sa);
        __syncwarp();
    }

    // ---------------- phase A: arrivals of this window, in time order --------
    const int ea = e * A + agent;
    const int64_t aoff = d.arr_off[ea];
    const int an = d.arr_n[ea];
    int cur = d.arr_cur[ea];
    while (true) {
        const int idx = cur + lane;
        const float at = idx < an ? __ldcs(d.arr_time + aoff + idx) : MLB_INF;
        const bool inw = at < t1;
        const unsigned bal = __ballot_sync(MLB_FULL, inw);
        const int nv = (bal == MLB_FULL) ? 32 : (__ffs(~bal) - 1);
        if (nv == 0) break;
        const float awk = lane < nv ? __ldcs(d.arr_work + aoff + idx) : 0.f;
        int abk = 0;
        float au = 0.f;
        if (kAlias && lane < nv) {
            abk = __ldcs(d.arr_bucket + aoff + idx);
            au = __ldcs(d.arr_u + aoff + idx);
        }
        for (int i = 0; i < nv; i++) {
            const float a = __shfl_sync(MLB_FULL, at, i);
            const float wk = __shfl_sync(MLB_FULL, awk, i);
            for (int j = lane; j < Sa; j += 32) retire<POLICY>(d, s, j, sbase + j, seed0 + j, a);
            __syncwarp();
            int kstar;
            if (kAlias) {
                const int b = __shfl_sync(MLB_FULL, abk, i);
                const float u = __shfl_sync(MLB_FULL, au, i);
                kstar = ((double)u < s.prob[b]) ? b : s.alias[b];  // test_integration.py:57-63
            } else {
                uint32_t bkey = 0xffffffffu;
                int bj = 0x7fffffff;
                for (int j = lane; j < Sa; j += 32) {
                    const uint32_t key = s.score[j];
                    if (key < bkey) { bkey = key; bj = j; }  // strict <: first minimum wins
                }
                const uint32_t mkey = __reduce_min_sync(MLB_FULL, bkey);
                kstar = (int)__reduce_min_sync(MLB_FULL, (uint32_t)(bkey == mkey ? bj : 0x7fffffff));
            }
            if (lane == (kstar & 31)) {
                const int k = kstar;
                const int n = s.n_on[k];
                if (n >= d.Q) {
                    s.dropped[k] += 1;  // [B] drop-and-count
                } else {
                    const float start = fmaxf(s.last_fin[k], a);
                    const float fin = __fadd_rn(start, __fdiv_rn(wk, s.speed[k]));
                    uint32_t pos = s.head[k] + (uint32_t)n;
                    if (pos >= (uint32_t)d.Q) pos -= (uint32_t)d.Q;
                    const size_t rb = (sbase + k) * d.Q;
                    d.ring_arr[rb + pos] = a;
                    d.ring_fin[rb + pos] = fin;
                    if (n == 0) s.head_fin[k] = fin;
                    s.last_fin[k] = fin;
                    s.n_on[k] = n + 1;  // lbhash.h:142,167
                    if (!kAlias) s.score[k] = server_score<POLICY>(n + 1, s.weight[k]);
                }
            }
            if (d.record_assign && lane == 0) d.assign[aoff + cur + i] = seed0 + kstar;
            __syncwarp();
        }
        cur += nv;
        if (nv < 32) break;
    }
    if (lane == 0) d.arr_cur[ea] = cur;

    // ---------------- phase C: window end ------------------------------------
    for (int j = lane; j < Sa; j += 32) {
        const size_t srv = sbase + j;
        retire<POLICY>(d, s, j, srv, seed0 + j, t1);
        const int n = s.n_on[j];
        uint32_t pos = s.head[j];
        const size_t rb = srv * d.Q;
        for (int q = 0; q < n; q++) {  // lbhash.h:131-135, one sample per active flow per step
            const float arr = d.ring_arr[rb + pos];
            res_add(d, s, 1, j, srv, seed0 + j, __fsub_rn(t1, arr), t1);
            pos = (pos + 1 == (uint32_t)d.Q) ? 0u : pos + 1;
        }
        // write back per-server state and the n_flow_on column (features.py:274)
        d.n_on[srv] = n;
        d.last_fin[srv] = s.last_fin[j];
        d.head[srv] = s.head[j];
        d.dropped[srv] = s.dropped[j];
        d.obs[srv * MLB_OBS_COLS] = (float)n;
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const size_t c = ((size_t)e * 2 + m) * S + seed0 + j;
            d.res_count[c] = s.cnt[m][j];
            d.res_cursor[c] = s.cur[m][j];
        }
    }
    __syncwarp();

    // ---------------- phase B: statistics of touched reservoirs --------------
    for (int jb = 0; jb < SP; jb += 32) {
        const int j = jb + lane;
        uint32_t fl = 0;
        if (j < Sa) fl = d.feature_cache ? s.flags[j] : 3u;  // mode 0: recompute every reservoir
#pragma unroll
        for (int m = 0; m < 2; m++) {
            unsigned todo = __ballot_sync(MLB_FULL, (fl >> m) & 1u);
            while (todo) {
                const int jj = jb + __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t cnt = s.cnt[m][jj];
                const int n = cnt < (uint32_t)d.K ? (int)cnt : d.K;
                const size_t rid = ((sbase + jj) * 2 + m) * d.KP;
                uint32_t mw[4];
#pragma unroll
                for (int k = 0; k < 4; k++) mw[k] = s.chg[(m * 4 + k) * s.sp + jj];
                const int nchg = __popc(mw[0]) + __popc(mw[1]) + __popc(mw[2]) + __popc(mw[3]);
                float f[5];
                warp_features_cached(d.res_val + rid, d.res_ts + rid, d.res_rank + rid, n, (int)s.nold[m][jj],
                                     mw, nchg, d.feature_cache != 1, t1, d.decay, d.log2_decay, scratch, f);
                float mine = f[0];
#pragma unroll
                for (int q = 1; q < 5; q++) mine = lane == q ? f[q] : mine;
                if (lane < 5) d.obs[(sbase + jj) * MLB_OBS_COLS + 1 + 5 * m + lane] = mine;
            }
        }
    }
    __syncwarp();

    // ---------------- phase R: reward over the env's active servers ----------
    for (int j = lane; j < Sa; j += 32) {
        const float* row = d.obs + (sbase + j) * MLB_OBS_COLS;
        bool active = false;  // env.py:410-413
        float x = 0.f;
#pragma unroll
        for (int c = 0; c < MLB_OBS_COLS; c++) {
            const float o = row[c];
            active |= o > 0.f;
            x = c == d.reward_field ? o : x;
        }
        rv[seed0 + j] = x;
        ra[seed0 + j] = active ? 1u : 0u;
    }
    if (SINGLE) {
        __syncwarp();
    } else if (env_in_blk == 0) {  // at most 2 envs per block when A > 1: static barrier ids
        asm volatile("bar.sync 1, %0;" ::"r"(A * 32) : "memory");
    } else {
        asm volatile("bar.sync 2, %0;" ::"r"(A * 32) : "memory");
    }
    if (agent == 0) {
        const double r = reward_staged(d.reward_metric, rv, ra, S);
        if (lane == 0) {
            d.reward[e] = r;                               // multi_agent_env.py:143-145: same scalar for all agents
            d.done[e] = step >= d.max_steps ? 1 : 0;       // env.py:267
            d.step[e] = step;
        }
    }
}

}  // namespace mlb
