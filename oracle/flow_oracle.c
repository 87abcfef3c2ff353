/*
 * flow_oracle.c -- CPU restatement of the MARLLB simulation-mode hot path.
 * TEST INFRASTRUCTURE ONLY (see flow_oracle.h).  Build: oracle/Makefile.
 * Compile with -ffp-contract=off: float results must not depend on FMA fusion.
 */
#include "flow_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* =========================================================================
 * numpy legacy RandomState = MT19937 (numpy/random/src/mt19937, Matsumoto &
 * Nishimura 2002).  Third-party arithmetic not under /root/reference: numpy
 * 2.3.5 here (pinned original 1.16.4 uses the same stream).
 * Call sites in the reference: reservoir.py:45,76; env.py:127,436-444.
 * ========================================================================= */
void ora_mt_seed(ora_mt_t *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->mti = 624;
}

static void mt_refill(ora_mt_t *s) {
    uint32_t *mt = s->mt;
    int k;
    for (k = 0; k < 624 - 397; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
        mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; k < 623; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
        mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->mti = 0;
}

uint32_t ora_mt_u32(ora_mt_t *s) {
    if (s->mti >= 624) mt_refill(s);
    uint32_t y = s->mt[s->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* RandomState.randint(low, low+rng+1): smallest 2^k-1 mask >= rng, redraw one
 * 32-bit word until (word & mask) <= rng.  rng == 0 consumes nothing. */
uint32_t ora_mt_randint(ora_mt_t *s, uint32_t rng) {
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = ora_mt_u32(s) & mask; } while (v > rng);
    return v;
}

double ora_mt_double(ora_mt_t *s) {
    uint32_t a = ora_mt_u32(s) >> 5, b = ora_mt_u32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

void ora_mt_fill(uint32_t seed, uint32_t *out, int64_t n) {
    ora_mt_t s;
    ora_mt_seed(&s, seed);
    for (int64_t i = 0; i < n; i++) out[i] = ora_mt_u32(&s);
}

/* =========================================================================
 * numpy add.reduce semantics: out = pairwise_sum(a) with numpy's
 * 8-accumulator pairwise blocks (numpy/_core/src/umath/loops_utils.h.src).
 * Restated so means/stds agree with the reference's np.mean / np.std / np.sum
 * to the last bit where possible.
 * ========================================================================= */
#define DEFINE_PAIRWISE(NAME, T)                                                  \
    static T NAME(const T *a, long n) {                                           \
        if (n < 8) {                                                              \
            T res = (T)-0.0;                                                      \
            for (long i = 0; i < n; i++) res += a[i];                             \
            return res;                                                           \
        } else if (n <= 128) {                                                    \
            T r[8];                                                               \
            long i;                                                               \
            for (int j = 0; j < 8; j++) r[j] = a[j];                              \
            for (i = 8; i < n - (n % 8); i += 8)                                  \
                for (int j = 0; j < 8; j++) r[j] += a[i + j];                     \
            T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])); \
            for (; i < n; i++) res += a[i];                                       \
            return res;                                                           \
        } else {                                                                  \
            long n2 = n / 2;                                                      \
            n2 -= n2 % 8;                                                         \
            return NAME(a, n2) + NAME(a + n2, n - n2);                            \
        }                                                                         \
    }                                                                             \
    static T NAME##_reduce(const T *a, long n) {                                  \
        if (n <= 0) return (T)0;                                                  \
        return NAME(a, n); /* initial = identity 0: whole array is pairwise-summed */ \
    }
DEFINE_PAIRWISE(pw_f32, float)
DEFINE_PAIRWISE(pw_f64, double)

/* =========================================================================
 * Philox4x32-10: counter-based stream of rng_mode 1 (twin of csrc/mlb_step_kernel.cuh)
 * ========================================================================= */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t ora_philox_word(uint32_t c, uint32_t j, uint32_t g, uint32_t key0) {
    const uint32_t ctr[4] = {c >> 2, j, g, 0x52535652u}, key[2] = {key0, 0x4d4c4232u};
    uint32_t out[4];
    ora_philox4x32_10(ctr, key, out);
    return out[c & 3u];
}

/* =========================================================================
 * ReservoirSampler (reservoir.py:17-233)
 * ========================================================================= */
static int g_last_slot_dummy;

ora_reservoir_t *ora_reservoir_create(int capacity, uint32_t seed) {
    ora_reservoir_t *r = (ora_reservoir_t *)calloc(1, sizeof(*r));
    r->capacity = capacity;
    r->values = (float *)calloc((size_t)capacity, sizeof(float));          /* :41 */
    r->timestamps = (double *)calloc((size_t)capacity, sizeof(double));    /* :42 */
    ora_mt_seed(&r->rng, seed);                                            /* :45 */
    (void)g_last_slot_dummy;
    return r;
}

void ora_reservoir_set_philox(ora_reservoir_t *r, uint32_t key0, uint32_t server, uint32_t env) {
    r->rng_mode = 1;
    r->ph_key0 = key0; r->ph_server = server; r->ph_env = env;
    r->ph_cursor = 0;
}

void ora_reservoir_destroy(ora_reservoir_t *r) {
    if (!r) return;
    free(r->values);
    free(r->timestamps);
    free(r);
}

void ora_reservoir_reset(ora_reservoir_t *r) { /* :220-225 (rng is NOT reseeded) */
    r->count = 0;
    memset(r->values, 0, sizeof(float) * (size_t)r->capacity);
    memset(r->timestamps, 0, sizeof(double) * (size_t)r->capacity);
}

static int reservoir_add_slot(ora_reservoir_t *r, float value, double ts) {
    /* returns slot written, or -1 when rejected */
    if (r->count < (uint64_t)r->capacity) {                                /* :65-73 */
        int slot = (int)r->count;
        r->values[slot] = value;
        r->timestamps[slot] = ts;
        r->count++;
        return slot;
    }
    uint32_t j;
    if (r->rng_mode == 1) { /* same masked rejection as RandomState.randint, over the counter-based words */
        const uint32_t rng = (uint32_t)r->count;
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        do { j = ora_philox_word(r->ph_cursor++, r->ph_server, r->ph_env, r->ph_key0) & mask; } while (j > rng);
    } else {
        j = ora_mt_randint(&r->rng, (uint32_t)r->count);                   /* :76 randint(0, count+1) */
    }
    r->count++;                                                            /* :81/:84 */
    if (j < (uint32_t)r->capacity) {                                       /* :78-80 */
        r->values[j] = value;
        r->timestamps[j] = ts;
        return (int)j;
    }
    return -1;
}

static __thread int tl_last_slot = -1;

int ora_reservoir_add(ora_reservoir_t *r, float value, double ts) {
    tl_last_slot = reservoir_add_slot(r, value, ts);
    return tl_last_slot >= 0;
}

int ora_reservoir_last_slot(const ora_reservoir_t *r) {
    (void)r;
    return tl_last_slot;
}

void ora_reservoir_get(const ora_reservoir_t *r, float *values, double *ts, uint64_t *count) {
    if (values) memcpy(values, r->values, sizeof(float) * (size_t)r->capacity);
    if (ts) memcpy(ts, r->timestamps, sizeof(double) * (size_t)r->capacity);
    if (count) *count = r->count;
}

typedef struct { float v; double w; int i; } vw_t;

static int cmp_vw(const void *pa, const void *pb) {
    const vw_t *a = (const vw_t *)pa, *b = (const vw_t *)pb;
    if (a->v < b->v) return -1;
    if (a->v > b->v) return 1;
    return (a->i > b->i) - (a->i < b->i);
}

/* get_features (reservoir.py:105-163) + _weighted_percentile (:165-196) with
 * numpy's arithmetic restated: np.mean / np.std on float32 arrays accumulate
 * pairwise in float32; np.percentile(.,90) is the 'linear' method evaluated in
 * float64; np.average(values, weights) and the weighted percentile in float64. */
void ora_features(const float *values, const double *ts, int n, double decay,
                  double now, double out5[5]) {
    if (n <= 0) {                                                          /* :127-134 */
        for (int k = 0; k < 5; k++) out5[k] = 0.0;
        return;
    }
    /* mean = np.mean(values): f32 pairwise sum / n in f32                  :143 */
    float mean = pw_f32_reduce(values, n) / (float)n;
    /* std = np.std(values): sqrt(mean(|x-mean|^2)) in f32                  :145 */
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; i++) {
        float d = values[i] - mean;
        tmp[i] = d * d;
    }
    float var = pw_f32_reduce(tmp, n) / (float)n;
    float std = sqrtf(var);
    free(tmp);
    /* sort once by (value, index) */
    vw_t *vw = (vw_t *)malloc(sizeof(vw_t) * (size_t)n);
    for (int i = 0; i < n; i++) {
        vw[i].v = values[i];
        vw[i].w = pow(decay, now - ts[i]);                                 /* :148-149 */
        vw[i].i = i;
    }
    /* mean_decay = np.average(values, weights): sum(v*w)/sum(w) in f64     :152 */
    double *prod = (double *)malloc(sizeof(double) * (size_t)n * 2);
    double *wts = prod + n;
    for (int i = 0; i < n; i++) {
        wts[i] = vw[i].w;
        prod[i] = (double)values[i] * vw[i].w;
    }
    double scl = pw_f64_reduce(wts, n);
    double mean_decay = pw_f64_reduce(prod, n) / scl;
    free(prod);
    qsort(vw, (size_t)n, sizeof(vw_t), cmp_vw);
    /* p90 = np.percentile(values, 90), method 'linear'                     :144
     * numpy >= 2: q = 90 / float32(100) is float32 for a float32 array, and the
     * virtual index, gamma and the lerp are all evaluated in float32
     * (numpy/lib/_function_base_impl.py: percentile, _quantile, _lerp). */
    float qf = 90.0f / 100.0f;
    float vidx = (float)(n - 1) * qf;
    long lo = (long)floorf(vidx);
    long hi = lo + 1;
    if (vidx >= (float)(n - 1)) { lo = n - 1; hi = n - 1; }
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;
    float gamma = vidx - floorf(vidx);
    float a = vw[lo].v, b = vw[hi].v;
    float diff = b - a;
    float p90f = a + diff * gamma;
    if (gamma >= 0.5f) p90f = b - diff * (1.0f - gamma);
    double p90 = (double)p90f;
    /* p90_decay: cumulative weights in value order, searchsorted-left      :181-196 */
    double c = 0.0;
    double total = 0.0;
    for (int i = 0; i < n; i++) total += vw[i].w;   /* cumsum[-1]: sequential adds */
    double cutoff = 0.9 * total;
    int idx = n;
    for (int i = 0; i < n; i++) {
        c += vw[i].w;
        if (c >= cutoff) { idx = i; break; }
    }
    if (idx >= n) idx = n - 1;
    double p90_decay = (double)vw[idx].v;
    free(vw);
    out5[0] = (double)mean;
    out5[1] = p90;
    out5[2] = (double)std;
    out5[3] = mean_decay;
    out5[4] = p90_decay;
}

void ora_reservoir_features(const ora_reservoir_t *r, double decay, double now, double out5[5]) {
    int n = r->count < (uint64_t)r->capacity ? (int)r->count : r->capacity; /* :87-89 */
    ora_features(r->values, r->timestamps, n, decay, now, out5);
}

/* =========================================================================
 * rewards.py
 * ========================================================================= */
static double np_var(const double *v, int n) {
    double mean = pw_f64_reduce(v, n) / (double)n;
    double *t = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double d = v[i] - mean;
        t[i] = d * d;
    }
    double r = pw_f64_reduce(t, n) / (double)n;
    free(t);
    return r;
}

double ora_reward_metric(int metric, const double *v, int n) {
    const double eps = 1e-10;
    switch (metric) {
    case ORA_JAIN: {                                                       /* rewards.py:21-67 */
        if (n == 0) return 1.0;
        double s = pw_f64_reduce(v, n);
        if (s < eps) return 1.0;
        double *sq = (double *)malloc(sizeof(double) * (size_t)n);
        for (int i = 0; i < n; i++) sq[i] = v[i] * v[i];
        double s2 = pw_f64_reduce(sq, n);
        free(sq);
        if (s2 < eps) return 1.0;
        double j = (s * s) / ((double)n * s2);
        double lo = 1.0 / (double)n;
        return j < lo ? lo : (j > 1.0 ? 1.0 : j);
    }
    case ORA_VARIANCE:                                                     /* :70-94 */
        if (n == 0) return 0.0;
        return -np_var(v, n);
    case ORA_STD:                                                          /* :97-114 */
        if (n == 0) return 0.0;
        return -sqrt(np_var(v, n));
    case ORA_CV: {                                                         /* :117-144 */
        if (n == 0) return 0.0;
        double mean = pw_f64_reduce(v, n) / (double)n;
        if (mean < eps) return 0.0;
        return -(sqrt(np_var(v, n)) / (mean + eps));
    }
    case ORA_MAX: {                                                        /* :147-171 */
        if (n == 0) return 0.0;
        double m = v[0];
        for (int i = 1; i < n; i++) if (v[i] > m) m = v[i];
        return -m;
    }
    case ORA_MIN: {                                                        /* :174-191 */
        if (n == 0) return 0.0;
        double m = v[0];
        for (int i = 1; i < n; i++) if (v[i] < m) m = v[i];
        return m;
    }
    case ORA_PRODUCT: {                                                    /* :194-225 */
        if (n == 0) return 0.0;
        double *l = (double *)malloc(sizeof(double) * (size_t)n);
        for (int i = 0; i < n; i++) l[i] = log(v[i] + eps);
        double s = pw_f64_reduce(l, n);
        free(l);
        return s;
    }
    case ORA_RANGE: {                                                      /* :228-246 */
        if (n == 0) return 0.0;
        double mx = v[0], mn = v[0];
        for (int i = 1; i < n; i++) {
            if (v[i] > mx) mx = v[i];
            if (v[i] < mn) mn = v[i];
        }
        return -(mx - mn);
    }
    case ORA_GINI: {                                                       /* :249-287 */
        if (n == 0) return 0.0;
        double mean = pw_f64_reduce(v, n) / (double)n;
        if (mean == 0) return 0.0;
        double ds = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) ds += fabs(v[i] - v[j]);
        return -(ds / (2.0 * n * n * mean));
    }
    /* ---- original testbed rewards, src/lb/env.py:73-156 (fair_fn); Python's builtin sum() is a
     * strictly sequential float64 sum, np.prod a sequential product, np.var as above ---- */
    case ORA_FAIR_JAIN: {                                                  /* src/lb/env.py:73-85 */
        if (n == 0) return 1.0;
        double s = 0.0, s2 = 0.0;
        for (int i = 0; i < n; i++) s += v[i];
        if (s == 0.0) return 1.0;
        for (int i = 0; i < n; i++) s2 += v[i] * v[i];
        return (s * s) / ((double)n * s2);
    }
    case ORA_FAIR_PRODUCT: {                                               /* src/lb/env.py:87-96 */
        if (n == 0) return 0.0;
        double m = v[0], p = 1.0;
        for (int i = 1; i < n; i++) if (v[i] > m) m = v[i];
        for (int i = 0; i < n; i++) p *= v[i] / (m + 1e-6);
        return p;
    }
    case ORA_VAR_EXP:                                                      /* src/lb/env.py:108-115, k = 10000 */
        if (n == 0) return 0.0;
        return exp(-10000.0 * np_var(v, n));
    case ORA_VAR_LOG:                                                      /* src/lb/env.py:118-125 */
        if (n == 0) return 0.0;
        return -log(np_var(v, n));
    case ORA_MAX_EXP: case ORA_MAX_LOG: {                                  /* src/lb/env.py:135-149 */
        if (n == 0) return 0.0;
        double m = v[0];
        for (int i = 1; i < n; i++) if (v[i] > m) m = v[i];
        return metric == ORA_MAX_EXP ? exp(-10000.0 * m) : -log(m);
    }
    default:
        return NAN;
    }
}

double ora_reward_from_obs(int metric, int field, const float *obs, int S) {
    /* env.py:410-417: active <=> any(obs[s] > 0); rewards.py:361-381 */
    double *vals = (double *)malloc(sizeof(double) * (size_t)(S > 0 ? S : 1));
    int n = 0;
    for (int s = 0; s < S; s++) {
        int active = 0;
        for (int c = 0; c < 11; c++) if (obs[s * 11 + c] > 0.0f) active = 1;
        if (active) vals[n++] = (double)obs[s * 11 + field];
    }
    double r = (n == 0) ? 0.0 : ora_reward_metric(metric, vals, n);          /* :364-365 */
    free(vals);
    return r;
}

/* =========================================================================
 * rl_controller.py:359-405  _build_alias_table (stack-pop pairing)
 * ========================================================================= */
void ora_alias_build(const double *p, int n, double *prob, int32_t *alias) {
    int *small = (int *)malloc(sizeof(int) * (size_t)n * 2);
    int *large = small + n;
    int ns = 0, nl = 0;
    for (int i = 0; i < n; i++) {
        prob[i] = p[i] * (double)n;                                        /* :378 */
        alias[i] = i;                                                      /* :379 */
    }
    for (int i = 0; i < n; i++) {                                          /* :385-389 */
        if (prob[i] < 1.0) small[ns++] = i; else large[nl++] = i;
    }
    while (ns > 0 && nl > 0) {                                             /* :392-403 */
        int l = small[--ns];
        int g = large[--nl];
        alias[l] = g;
        prob[g] = prob[g] + prob[l] - 1.0;
        if (prob[g] < 1.0) small[ns++] = g; else large[nl++] = g;
    }
    free(small);
}

/* =========================================================================
 * env.py:425-448  _simulate_observation (legacy random features)
 * The obs array is float32, so columns 4,5,9,10 are computed from the value
 * ALREADY ROUNDED to float32.  Under numpy >= 2 (NEP 50; this image has 2.3.5)
 * `np.float32 * 0.9` is a float32 multiply by float32(0.9); numpy 1.x used
 * float64 and rounded again.  The oracle follows the numpy in this image, which
 * is what generated tests/golden/.
 * ========================================================================= */
void ora_legacy_obs(ora_mt_t *g, int S, float *obs) {
    for (int s = 0; s < S; s++) {
        float *o = obs + s * 11;
        o[0] = (float)(5 + (int)ora_mt_randint(g, 14));                    /* :436 randint(5,20) */
        o[1] = (float)(5.0 + (15.0 - 5.0) * ora_mt_double(g));             /* :437 */
        o[2] = (float)(10.0 + (25.0 - 10.0) * ora_mt_double(g));           /* :438 */
        o[3] = (float)(1.0 + (5.0 - 1.0) * ora_mt_double(g));              /* :439 */
        o[4] = o[1] * 0.9f;                                               /* :440 */
        o[5] = o[2] * 0.9f;                                               /* :441 */
        o[6] = (float)(8.0 + (18.0 - 8.0) * ora_mt_double(g));             /* :442 */
        o[7] = (float)(15.0 + (30.0 - 15.0) * ora_mt_double(g));           /* :443 */
        o[8] = (float)(2.0 + (8.0 - 2.0) * ora_mt_double(g));              /* :444 */
        o[9] = o[6] * 0.85f;                                              /* :445 */
        o[10] = o[6] * 0.9f;                                              /* :446 */
    }
}

/* =========================================================================
 * Composed flow-level env (mirrors oracle/ref_flow_env.py)
 * ========================================================================= */
typedef struct { float arr, fin; } flow_t;

struct ora_env {
    ora_env_cfg cfg;
    int S;
    int64_t cur_step;
    float *speeds;
    int32_t *n_on;
    float *last_fin;
    flow_t *ring;      /* [S][Q] circular */
    uint32_t *head, *tail;
    int64_t *dropped;
    ora_reservoir_t **res; /* [S][2] */
    const float **a_time, **a_work, **a_u;
    const int32_t **a_bucket;
    int64_t *a_n, *a_cursor;
    double *alias_prob;
    int32_t *alias_idx;
};

ora_env *ora_env_create(const ora_env_cfg *cfg, const float *speeds) {
    ora_env *e = (ora_env *)calloc(1, sizeof(*e));
    e->cfg = *cfg;
    int S = cfg->num_agents * cfg->servers_per_agent, A = cfg->num_agents;
    e->S = S;
    e->speeds = (float *)malloc(sizeof(float) * (size_t)S);
    memcpy(e->speeds, speeds, sizeof(float) * (size_t)S);
    e->n_on = (int32_t *)calloc((size_t)S, sizeof(int32_t));
    e->last_fin = (float *)calloc((size_t)S, sizeof(float));
    e->ring = (flow_t *)calloc((size_t)S * (size_t)cfg->queue_cap, sizeof(flow_t));
    e->head = (uint32_t *)calloc((size_t)S, sizeof(uint32_t));
    e->tail = (uint32_t *)calloc((size_t)S, sizeof(uint32_t));
    e->dropped = (int64_t *)calloc((size_t)S, sizeof(int64_t));
    e->res = (ora_reservoir_t **)calloc((size_t)S * 2, sizeof(void *));
    for (int j = 0; j < S; j++)
        for (int m = 0; m < 2; m++) /* same seed for both metrics: reservoir.py:261-265; seed=j: basic_usage.py:157-163 */
        {
            ora_reservoir_t *r = ora_reservoir_create(cfg->reservoir_k, cfg->seed_base + (uint32_t)j);
            r->rng_mode = cfg->rng_mode;
            r->ph_key0 = cfg->seed_base; r->ph_server = (uint32_t)j; r->ph_env = cfg->env_id;
            e->res[j * 2 + m] = r;
        }
    e->a_time = (const float **)calloc((size_t)A, sizeof(void *));
    e->a_work = (const float **)calloc((size_t)A, sizeof(void *));
    e->a_u = (const float **)calloc((size_t)A, sizeof(void *));
    e->a_bucket = (const int32_t **)calloc((size_t)A, sizeof(void *));
    e->a_n = (int64_t *)calloc((size_t)A, sizeof(int64_t));
    e->a_cursor = (int64_t *)calloc((size_t)A, sizeof(int64_t));
    e->alias_prob = (double *)calloc((size_t)S, sizeof(double));
    e->alias_idx = (int32_t *)calloc((size_t)S, sizeof(int32_t));
    return e;
}

void ora_env_destroy(ora_env *e) {
    if (!e) return;
    for (int i = 0; i < e->S * 2; i++) ora_reservoir_destroy(e->res[i]);
    free(e->speeds); free(e->n_on); free(e->last_fin); free(e->ring);
    free(e->head); free(e->tail); free(e->dropped); free(e->res);
    free(e->a_time); free(e->a_work); free(e->a_u); free(e->a_bucket);
    free(e->a_n); free(e->a_cursor); free(e->alias_prob); free(e->alias_idx);
    free(e);
}

void ora_env_set_arrivals(ora_env *e, int agent, const float *time, const float *work,
                          const int32_t *bucket, const float *u, int64_t n) {
    e->a_time[agent] = time;
    e->a_work[agent] = work;
    e->a_bucket[agent] = bucket;
    e->a_u[agent] = u;
    e->a_n[agent] = n;
    e->a_cursor[agent] = 0;
}

void ora_env_reset(ora_env *e) {
    int S = e->S;
    e->cur_step = 0;
    memset(e->n_on, 0, sizeof(int32_t) * (size_t)S);
    memset(e->last_fin, 0, sizeof(float) * (size_t)S);
    memset(e->head, 0, sizeof(uint32_t) * (size_t)S);
    memset(e->tail, 0, sizeof(uint32_t) * (size_t)S);
    memset(e->dropped, 0, sizeof(int64_t) * (size_t)S);
    for (int j = 0; j < S; j++)
        for (int m = 0; m < 2; m++) { /* a fresh MultiMetricReservoir(seed=j) per episode */
            ora_reservoir_reset(e->res[j * 2 + m]);
            ora_mt_seed(&e->res[j * 2 + m]->rng, e->cfg.seed_base + (uint32_t)j);
            e->res[j * 2 + m]->ph_cursor = 0;
        }
    for (int i = 0; i < e->cfg.num_agents; i++) e->a_cursor[i] = 0;
}

static void env_retire(ora_env *e, int j, float now) {
    const int Q = e->cfg.queue_cap;
    flow_t *ring = e->ring + (size_t)j * (size_t)Q;
    while (e->head[j] != e->tail[j]) {
        flow_t f = ring[e->head[j] % (uint32_t)Q];
        if (!(f.fin < now)) break;
        e->head[j]++;
        e->n_on[j] -= 1;                                                   /* lbhash.h:120 */
        float fct = f.fin - f.arr;
        reservoir_add_slot(e->res[j * 2 + 0], fct, (double)f.fin);         /* lbhash.h:122-124 */
    }
}

int64_t ora_env_step(ora_env *e, const void *action, float *obs, double *reward,
                     uint8_t *done, int32_t *assign, int64_t assign_cap) {
    const ora_env_cfg *c = &e->cfg;
    const int S = e->S, Sa = c->servers_per_agent, A = c->num_agents, Q = c->queue_cap;
    e->cur_step += 1;                                                      /* env.py:230 */
    const float t1 = (float)e->cur_step * c->dt;
    /* weights: env.py:334-353 */
    float *w = (float *)malloc(sizeof(float) * (size_t)S);
    if (c->action_kind == ORA_DISCRETE) {
        const int32_t *a = (const int32_t *)action;
        for (int j = 0; j < S; j++) w[j] = c->discrete_weights[a[j]];      /* :346 */
    } else {
        const float *a = (const float *)action;
        for (int j = 0; j < S; j++) {                                      /* :349-351 np.clip */
            float x = a[j];
            w[j] = x < c->min_weight ? c->min_weight : (x > c->max_weight ? c->max_weight : x);
        }
    }
    if (c->policy == ORA_ALIAS) {
        double *p = (double *)malloc(sizeof(double) * (size_t)Sa);
        for (int i = 0; i < A; i++) {
            double tot = 0.0; /* wi.sum(): numpy pairwise over f64 */
            for (int k = 0; k < Sa; k++) p[k] = (double)w[i * Sa + k];
            tot = pw_f64_reduce(p, Sa);
            for (int k = 0; k < Sa; k++) p[k] = p[k] / tot;
            ora_alias_build(p, Sa, e->alias_prob + i * Sa, e->alias_idx + i * Sa);
        }
        free(p);
    }
    int64_t n_flows = 0;
    for (int i = 0; i < A; i++) {
        int64_t cur = e->a_cursor[i];
        const int lo = i * Sa;
        while (cur < e->a_n[i] && e->a_time[i][cur] < t1) {
            const float a = e->a_time[i][cur];
            const float wk = e->a_work[i][cur];
            for (int j = lo; j < lo + Sa; j++) env_retire(e, j, a);
            int best = lo;
            if (c->policy == ORA_SED) {                                    /* node.c:395-404 */
                float bs = 0.f;
                for (int j = lo; j < lo + Sa; j++) {
                    float s = (float)((double)(e->n_on[j] + 1) / (1e-9 + (double)w[j]));
                    if (j == lo || s < bs) { best = j; bs = s; }
                }
            } else if (c->policy == ORA_LSQ) {                             /* node.c:419-431 */
                float bs = 0.f;
                for (int j = lo; j < lo + Sa; j++) {
                    float s = (float)e->n_on[j];
                    if (j == lo || s < bs) { best = j; bs = s; }
                }
            } else if (c->policy == ORA_SED2 || c->policy == ORA_LSQ2) {   /* node.c:408-417, 433-441: power of two */
                /* asindex0 = new_flow_table[hash & mask], asindex1 = new_flow_table[(hash+1) & mask]:
                 * [B] the pre-drawn bucket stands for the hash and the table is the identity mod Sa */
                const int c0 = e->a_bucket[i][cur];
                const int c1 = (c0 + 1 == Sa) ? 0 : c0 + 1;
                float s0, s1;
                if (c->policy == ORA_SED2) {
                    s0 = (float)((double)(e->n_on[lo + c0] + 1) / (1e-9 + (double)w[lo + c0]));
                    s1 = (float)((double)(e->n_on[lo + c1] + 1) / (1e-9 + (double)w[lo + c1]));
                } else {
                    s0 = (float)e->n_on[lo + c0];
                    s1 = (float)e->n_on[lo + c1];
                }
                best = lo + (s1 < s0 ? c1 : c0);
            } else {                                                       /* node.c:442-460; rule of test_integration.py:57-63 */
                int b = e->a_bucket[i][cur];
                double u = (double)e->a_u[i][cur];
                best = lo + ((u < e->alias_prob[lo + b]) ? b : e->alias_idx[lo + b]);
            }
            if (assign && n_flows < assign_cap) assign[n_flows] = best;
            n_flows++;
            if (e->tail[best] - e->head[best] >= (uint32_t)Q) {
                e->dropped[best]++;
            } else {
                float start = e->last_fin[best] > a ? e->last_fin[best] : a;
                float svc = wk / e->speeds[best];
                float fin = start + svc;
                flow_t *ring = e->ring + (size_t)best * (size_t)Q;
                ring[e->tail[best] % (uint32_t)Q].arr = a;
                ring[e->tail[best] % (uint32_t)Q].fin = fin;
                e->tail[best]++;
                e->last_fin[best] = fin;
                e->n_on[best] += 1;                                        /* lbhash.h:142,167 */
            }
            cur++;
        }
        e->a_cursor[i] = cur;
    }
    for (int j = 0; j < S; j++) {
        env_retire(e, j, t1);
        flow_t *ring = e->ring + (size_t)j * (size_t)Q;
        for (uint32_t h = e->head[j]; h != e->tail[j]; h++) {              /* lbhash.h:131-135 */
            float dur = t1 - ring[h % (uint32_t)Q].arr;
            reservoir_add_slot(e->res[j * 2 + 1], dur, (double)t1);
        }
        double f5[5];
        float *o = obs + j * 11;
        o[0] = (float)e->n_on[j];                                          /* features.py:274 */
        for (int m = 0; m < 2; m++) {                                      /* reservoir.py:295-308 */
            ora_reservoir_features(e->res[j * 2 + m], c->decay, (double)t1, f5);
            for (int k = 0; k < 5; k++) o[1 + m * 5 + k] = (float)f5[k];   /* :212-218 float32 vector */
        }
    }
    *reward = ora_reward_from_obs(c->reward_metric, c->reward_field, obs, S); /* env.py:259-262 */
    *done = (uint8_t)(e->cur_step >= c->max_steps);                        /* env.py:267 */
    free(w);
    return n_flows;
}

void ora_env_dump(const ora_env *e, int32_t *n_flow_on, float *res_values, float *res_ts,
                  int64_t *res_count, int64_t *dropped) {
    const int S = e->S, K = e->cfg.reservoir_k;
    for (int j = 0; j < S; j++) {
        if (n_flow_on) n_flow_on[j] = e->n_on[j];
        if (dropped) dropped[j] = e->dropped[j];
        for (int m = 0; m < 2; m++) {
            const ora_reservoir_t *r = e->res[j * 2 + m];
            if (res_count) res_count[j * 2 + m] = (int64_t)r->count;
            for (int k = 0; k < K; k++) {
                if (res_values) res_values[((size_t)j * 2 + m) * K + k] = r->values[k];
                if (res_ts) res_ts[((size_t)j * 2 + m) * K + k] = (float)r->timestamps[k];
            }
        }
    }
}

int ora_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    ora_env **envs;
    int begin, end;
    const char *actions;
    int stride;
    float *obs;
    double *reward;
    uint8_t *done;
    int64_t flows;
} batch_job_t;

static void *batch_worker(void *arg) {
    batch_job_t *j = (batch_job_t *)arg;
    for (int i = j->begin; i < j->end; i++) {
        int S = j->envs[i]->S;
        j->flows += ora_env_step(j->envs[i], j->actions + (size_t)i * (size_t)j->stride,
                                 j->obs + (size_t)i * (size_t)S * 11, j->reward + i, j->done + i,
                                 NULL, 0);
    }
    return NULL;
}

/* CPU-baseline driver: n independent envs split over `nthreads` pthreads. */
int64_t ora_env_step_batch(ora_env **envs, int n, const void *actions, int action_stride_bytes,
                           float *obs, double *reward, uint8_t *done, int nthreads) {
    if (nthreads <= 0) nthreads = ora_max_threads();
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    batch_job_t *jobs = (batch_job_t *)calloc((size_t)nthreads, sizeof(batch_job_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].envs = envs;
        jobs[t].begin = (int)((int64_t)n * t / nthreads);
        jobs[t].end = (int)((int64_t)n * (t + 1) / nthreads);
        jobs[t].actions = (const char *)actions;
        jobs[t].stride = action_stride_bytes;
        jobs[t].obs = obs;
        jobs[t].reward = reward;
        jobs[t].done = done;
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    int64_t total = jobs[0].flows;
    for (int t = 1; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].flows;
    }
    free(th);
    free(jobs);
    return total;
}
