/*
 * marllb_b200.h -- C ABI of the B200-native MARLLB simulation-mode hot path.
 *
 * Drop-in boundary.  The reference is pure Python with no FFI
 * of its own; each entry point below names the reference interface it stands
 * in for (paths relative to the reference tree).  Plain pointers and sizes
 * only -- no torch / numpy types.  Every function returns 0 on success or a
 * negative MLB_E* code and never throws; mlb_last_error() gives the text.
 *
 * Memory locations are explicit: MLB_HOST pointers are ordinary (ideally
 * pinned) host memory, MLB_DEVICE pointers are CUDA device memory on the
 * handle's device.  `stream` is a cudaStream_t passed as void* (NULL = legacy
 * default stream).  Calls are asynchronous on `stream` unless they involve
 * pageable host memory or say otherwise.
 */
#ifndef MARLLB_B200_H
#define MARLLB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLB_ABI_VERSION 1

/* status codes */
enum {
    MLB_OK = 0,
    MLB_EINVAL = -1,     /* bad argument / unsupported configuration            */
    MLB_ECUDA = -2,      /* CUDA runtime error (text in mlb_last_error)          */
    MLB_ENOMEM = -3,
    MLB_ESTATE = -4,     /* call order violated (e.g. step before arrivals)      */
    MLB_ERNG = -5,       /* replayed MT19937 stream exhausted (raise rng_table_len) */
    MLB_EACTION = -6     /* discrete action outside [0, n_discrete)              */
};

enum { MLB_HOST = 0, MLB_DEVICE = 1 };

/* server-assignment rule per new flow: src/vpp/lb/node.c:393-460 */
enum {
    MLB_POLICY_SED = 0,   /* argmin (n_flow_on+1)/(1e-9+w) over all servers, first minimum   node.c:393-406 */
    MLB_POLICY_LSQ = 1,   /* argmin n_flow_on                                                 node.c:419-431 */
    MLB_POLICY_ALIAS = 2, /* alias-table draw from pre-drawn (bucket, u)                      node.c:442-460 */
    /* power of two choices: candidates c0 = pre-drawn bucket (new_flow_table[hash]) and
     * c1 = (c0 + 1) mod servers_per_agent (new_flow_table[hash + 1]); c1 wins on a strictly lower score */
    MLB_POLICY_SED2 = 3,  /* SED score on the two candidates                                  node.c:408-417 */
    MLB_POLICY_LSQ2 = 4   /* n_flow_on on the two candidates                                  node.c:433-441 */
};

/* action encodings: problem-03-rl-environment/src/env.py:334-353 */
enum {
    MLB_ACTION_DISCRETE_I32 = 0, /* int32 index into discrete_weights            */
    MLB_ACTION_CONTINUOUS_F32 = 1, /* float weight, clipped to [min_w, max_w]    */
    MLB_ACTION_DISCRETE_U8 = 2   /* same as DISCRETE_I32, one byte per server    */
};

/* Reservoir index stream (the `j = rng.randint(0, count + 1)` of reservoir.py:76).
 *   MLB_RNG_REPLAY: every reservoir of server j replays np.random.RandomState(rng_seed_base + j) from a table of
 *     rng_table_len raw MT19937 words -- bit-for-bit the reference's stream, shared by all envs (reference parity;
 *     needs S * rng_table_len * 4 bytes of table and fails with MLB_ERNG when a row is used up).
 *   MLB_RNG_PHILOX: word c of the stream of (global env g, server j) is
 *     philox4x32_10(counter = {c >> 2, j, g, 0x52535652}, key = {rng_seed_base, 0x4d4c4232})[c & 3],
 *     consumed with the same masked-rejection rule.  Independent per env, no table, no length cap; results do not
 *     depend on how envs are sharded (g = env_id_base + local index).  CPU twin: oracle/flow_oracle.c. */
enum { MLB_RNG_REPLAY = 0, MLB_RNG_PHILOX = 1 };

/* reward metrics 0-8: problem-03-rl-environment/src/rewards.py:297-307;
 * 9-14: the original testbed's fair_fn table, src/lb/env.py:73-156 ('var' and 'max' of that table
 * are MLB_REWARD_VARIANCE and MLB_REWARD_MAX) */
enum {
    MLB_REWARD_JAIN = 0, MLB_REWARD_VARIANCE, MLB_REWARD_STD, MLB_REWARD_CV,
    MLB_REWARD_MAX, MLB_REWARD_MIN, MLB_REWARD_PRODUCT, MLB_REWARD_RANGE,
    MLB_REWARD_GINI,
    MLB_REWARD_FAIR_JAIN,     /* (sum x)^2 / (n sum x^2), 1 when sum x == 0, unclipped   src/lb/env.py:73-85   */
    MLB_REWARD_FAIR_PRODUCT,  /* prod(x / (max x + 1e-6))                                src/lb/env.py:87-96   */
    MLB_REWARD_VAR_EXP,       /* exp(-10000 var x)                                       src/lb/env.py:108-115 */
    MLB_REWARD_VAR_LOG,       /* -log(var x)                                             src/lb/env.py:118-125 */
    MLB_REWARD_MAX_EXP,       /* exp(-10000 max x)                                       src/lb/env.py:142-149 */
    MLB_REWARD_MAX_LOG,       /* -log(max x)                                             src/lb/env.py:135-139 */
    MLB_REWARD_COUNT_
};

/* state fields readable through mlb_get_state (parity dumps) */
enum {
    MLB_F_N_FLOW_ON = 0, /* int32  [E][S]                                        */
    MLB_F_RES_VALUES,    /* float  [E][S][2][K]   reservoir.py:41                */
    MLB_F_RES_TS,        /* float  [E][S][2][K]   reservoir.py:42 (f32 seconds)  */
    MLB_F_RES_COUNT,     /* uint32 [E][2][S]      reservoir.py:40                */
    MLB_F_RES_CURSOR,    /* uint32 [E][2][S]      words consumed from the MT stream */
    MLB_F_DROPPED,       /* uint32 [E][S]                                        */
    MLB_F_LAST_FIN,      /* float  [E][S]                                        */
    MLB_F_HEAD,          /* uint32 [E][S]  ring position of the oldest in-system flow */
    MLB_F_STEP,          /* int32  [E]            env.py:146 current_step        */
    MLB_F_OBS,           /* float  [E][S][11]     env.py:46-49                   */
    MLB_F_ARR_CURSOR,    /* int32  [E][A]                                        */
    MLB_F_COUNT_
};

/*
 * Configuration.  Mirrors LoadBalanceEnv.__init__ (env.py:71-87) and
 * MultiAgentLoadBalanceEnv.__init__ (problem-05-qmix/src/multi_agent_env.py:44-53)
 * plus the flow-level knobs of SURVEY App. B.  S = num_agents*servers_per_agent.
 */
typedef struct mlb_config {
    int32_t abi_version;       /* = MLB_ABI_VERSION                              */
    int32_t device;            /* CUDA device ordinal                             */
    int32_t num_envs;          /* E  independent env instances                    */
    int32_t num_agents;        /* A  LB agents per env (multi_agent_env.py:57)    */
    int32_t servers_per_agent; /* Sa (multi_agent_env.py:58); <= 256              */
    int32_t reservoir_k;       /* reservoir capacity, reservoir.py:31; <= 128     */
    int32_t queue_cap;         /* per-server in-system cap Q (paper 4.2: 160)     */
    int32_t policy;            /* MLB_POLICY_*                                    */
    int32_t action_kind;       /* MLB_ACTION_*                                    */
    int32_t n_discrete;        /* len(discrete_weights) <= 8, env.py:69,113       */
    float discrete_weights[8]; /* default {1.0,1.5,2.0}                           */
    float min_weight;          /* env.py:77                                       */
    float max_weight;          /* env.py:76                                       */
    float dt;                  /* step_interval, env.py:80                        */
    double decay;              /* decay_factor, reservoir.py:106                  */
    int32_t reward_metric;     /* MLB_REWARD_*                                    */
    int32_t reward_field;      /* obs column 0..10 (env.py:377-381); default 10   */
    int32_t max_steps;         /* env.py:81                                       */
    uint32_t rng_seed_base;    /* reservoirs of server j replay RandomState(seed_base+j) */
    int32_t rng_table_len;     /* 32-bit words replayed per seed (0 -> 65536)     */
    int32_t feature_cache;     /* 1: reuse features of reservoirs untouched this step */
    int32_t record_assign;     /* 1: keep the server id chosen for every flow     */
    int32_t env_id_base;       /* global id of env 0 (multi-GPU sharding): keys the
                                  synthetic-arrival streams so results do not depend
                                  on how envs are split over GPUs                  */
    int32_t rng_mode;          /* MLB_RNG_* (0 = replay: the default, and what zeroed reserved words meant) */
    int32_t reserved[5];
} mlb_config;

typedef struct mlb_env mlb_env; /* opaque handle */

/* Fill *cfg with the reference's defaults (env.py:71-87). */
int mlb_config_default(mlb_config *cfg);

/* Create / destroy.  Allocates all device state; replaces constructing E
 * LoadBalanceEnv objects (env.py:71-154). */
int mlb_create(const mlb_config *cfg, mlb_env **out);
int mlb_destroy(mlb_env *h);
const char *mlb_last_error(const mlb_env *h); /* h may be NULL: last create error */
int mlb_abi_version(void);

/* Per-server processing speed v_j (paper Alg.1 l.5).  n = S (broadcast to all
 * envs) or E*S.  Default 1.0. */
int mlb_set_speeds(mlb_env *h, const float *speeds, int64_t n, int loc, void *stream);

/* Arrival streams, one per (env, agent), time-sorted float32 seconds since
 * reset and float32 work units; CSR offsets[E*A+1].  bucket/u (nullable) are
 * the pre-drawn alias-method randoms (node.c:443-445).  Replaces
 * TrainingPipeline._load_traces/_generate_poisson_trace
 * (problem-06-vpp-integration/src/training_pipeline.py:98-155). Data is copied. */
int mlb_load_arrivals(mlb_env *h, const float *time, const float *work,
                      const int32_t *bucket, const float *u,
                      const int64_t *offsets, int loc, void *stream);

/* Streamed traces (hour-long trace files do not have to be resident).  The arrivals of the
 * NEXT chunk of env steps -- same layout as mlb_load_arrivals, HOST pointers (pinned for a truly asynchronous
 * copy; they must stay valid until `copy_stream` has passed the call) -- are copied into a second buffer set
 * on `copy_stream` while the envs keep stepping through the current chunk.  mlb_commit_arrivals, enqueued on
 * the stepping stream between two steps, waits for that copy, swaps the two sets and rewinds the arrival
 * cursors.  A chunk must hold exactly the arrivals of a whole number of step windows (time < t1 of its last
 * step, in float32 like the kernel's comparison), so no flow is left behind at the swap.
 * Replays what src/client/replay_fork_io.py:95-143 does against the real testbed.
 * Threading: mlb_stage_arrivals may be called from a second host thread while the first one is inside
 * mlb_step / mlb_reset on the same handle (it only touches the staging buffer set); every other pair of calls
 * on one handle must be serialised by the caller, like the reference env (not thread-safe). */
int mlb_stage_arrivals(mlb_env *h, const float *time, const float *work, const int32_t *bucket,
                       const float *u, const int64_t *offsets, void *copy_stream);
int mlb_commit_arrivals(mlb_env *h, void *stream);

/* Device-side synthetic Poisson arrivals (training_pipeline.py:141-155
 * semantics: exponential inter-arrivals at `rate`/s, kept while < horizon;
 * exponential work with mean `mean_work`), Philox4x32-10 keyed by
 * (seed, env*A+agent). */
int mlb_gen_poisson(mlb_env *h, double rate, double mean_work, double horizon,
                    uint64_t seed, void *stream);

/* The same generator for the time window [t_start, t_end) of a running episode: replaces the resident arrivals by a
 * fresh Poisson stream that starts at t_start (the process is memoryless) and rewinds the arrival cursors; call it
 * between two steps, with t_start = the end of the last stepped window.  `window` keys the Philox counter so that
 * successive windows draw different numbers (window 0, t_start 0 = mlb_gen_poisson).  Long episodes (env.py:81
 * defaults to 10 000 steps) then need only one window of arrivals in HBM at a time. */
int mlb_gen_poisson_window(mlb_env *h, double rate, double mean_work, double t_start, double t_end,
                           uint64_t seed, uint32_t window, void *stream);

/* Copy the arrival stream of (env, agent) back to the host (for CPU baselines).
 * Returns the flow count in *n; copies at most cap entries. Synchronous. */
int mlb_get_arrivals(mlb_env *h, int32_t env, int32_t agent, float *time, float *work,
                     int32_t *bucket, float *u, int64_t cap, int64_t *n);

/* reset(): env.py:186-213.  env_mask (nullable, host uint8[E]) selects envs. */
int mlb_reset(mlb_env *h, const uint8_t *env_mask, void *stream);

/*
 * step(): env.py:215-286 for all E envs at once.
 *   action   [E][S] in cfg.action_kind encoding, at action_loc
 *   out_obs  float  [E][S][11]  (nullable)   env.py:46-49
 *   out_reward double [E]       (nullable)   rewards.py:329-381
 *   out_done uint8  [E]         (nullable)   env.py:267
 * Outputs are copied to out_loc memory; pass NULL and use mlb_device_ptr() to
 * read the extension-owned device buffers in place (valid until the next step).
 */
int mlb_step(mlb_env *h, const void *action, int action_loc,
             float *out_obs, double *out_reward, uint8_t *out_done, int out_loc,
             void *stream);

/*
 * The same step for a HOST consumer that keeps its observation array across steps: instead of all E*S*11 floats,
 * only what changed crosses PCIe -- the n_flow_on column (2 bytes per row) and one 48-byte record {row index, the 11
 * floats} per row in which Algorithm R wrote a reservoir slot this step (untouched reservoirs keep their features:
 * both decay-weighted statistics are invariant to `now`) -- and host threads apply them to `host_obs`.
 *   host_obs   float [E][S][11], PINNED, persistent: must hold the observation of the previous step (after
 *              mlb_reset: zeros; after steps taken through mlb_step: call mlb_step with out_obs = host_obs once)
 *   action     HOST, cfg.action_kind encoding;  out_reward / out_done: HOST, nullable
 *   threads    host threads that apply the records (<= 0: min(hardware threads, 16))
 *   d2h_bytes  (nullable) receives the device->host bytes this call moved
 * Chunks of envs are pipelined like in mlb_step; a chunk in which most rows changed (early in an episode) is copied
 * as a plain block straight into place, so the call never moves more than mlb_step does.  Synchronous: returns when
 * host_obs is up to date.  The result is bit-identical to mlb_step's out_obs.  (No reference counterpart: the
 * reference env returns a fresh (S, 11) array per step, env.py:283-286.)
 */
int mlb_step_changed(mlb_env *h, const void *action, float *host_obs, double *out_reward, uint8_t *out_done,
                     int threads, int64_t *d2h_bytes, void *stream);

/* Per-flow server ids chosen so far (parallel to the arrival arrays), int32. */
int mlb_get_assignments(mlb_env *h, int32_t *dst, int64_t n, int loc, void *stream);

/* Extension-owned device buffers: field = MLB_F_*; special ids below. */
enum { MLB_PTR_OBS = 100, MLB_PTR_REWARD = 101, MLB_PTR_DONE = 102, MLB_PTR_ASSIGN = 103 };
int mlb_device_ptr(mlb_env *h, int what, void **ptr, size_t *bytes);

/* Copy a state field to dst (host or device); bytes must match. Synchronises. */
int mlb_get_state(mlb_env *h, int field, void *dst, size_t bytes, int loc);

/* Sticky device-side status (MLB_OK / MLB_ERNG / MLB_EACTION). Synchronises stream. */
int mlb_status(mlb_env *h, void *stream);

/* Kernels launched by this handle since creation (bench.py "gpu_launches"). */
int64_t mlb_launch_count(const mlb_env *h);

/* Per-kernel timing of mlb_step (measurement only; no reference counterpart). After
 * mlb_profile_begin the next max_steps calls of mlb_step record CUDA events around their two
 * kernels on the step's stream; mlb_profile_end synchronises those events and returns the
 * summed durations (ms) of the event kernel and of the statistics pass (pair_kernel +
 * feature_kernel) and the steps covered. */
int mlb_profile_begin(mlb_env *h, int max_steps);
int mlb_profile_end(mlb_env *h, double *event_ms, double *feature_ms, int *steps);
/* Of the feature_ms of the last mlb_profile_end (the statistics pass = pair_kernel + feature_kernel),
 * the part spent in pair_kernel. */
double mlb_profile_pair_ms(const mlb_env *h);

/* ---- stand-alone pieces (P01 reservoir, P03 rewards) ------------------------ */

/* numpy RandomState(seed) raw 32-bit stream (MT19937), host side. */
int mlb_mt19937_fill(uint32_t seed, uint32_t *out, int64_t n);

/* Batched ReservoirSampler.add (reservoir.py:50-85): R reservoirs of capacity K
 * (device arrays values[R][Kp], ts[R][Kp], count[R], cursor[R]; Kp = K rounded
 * up to 32), each fed n_add[r] samples add_v/add_t[r][max_add] in order.
 * seeds row r replays RandomState(seed_row[r]) from mt_table[row][table_len]. */
int mlb_reservoir_add(float *values, float *ts, uint32_t *count, uint32_t *cursor,
                      const uint32_t *mt_table, const int32_t *seed_row, int32_t table_len,
                      int32_t R, int32_t K, const float *add_v, const float *add_t,
                      const int32_t *n_add, int32_t max_add, uint8_t *accepted /* nullable [R][max_add] */,
                      int32_t *status /* device int */, void *stream);

/* Batched ReservoirSampler.get_features (reservoir.py:105-196): out [R][5] float. */
int mlb_reservoir_features(const float *values, const float *ts, const uint32_t *count,
                           int32_t R, int32_t K, double decay, const float *now /* [R] */,
                           float *out, void *stream);

/* Batched reward metric over rows of doubles (rewards.py:21-287): values[B][stride],
 * n[B] valid entries each; out double[B]. Device pointers. */
int mlb_reward_metric(int metric, const double *values, const int32_t *n, int32_t B,
                      int32_t stride, double *out, void *stream);

/* Batched _simulate_observation (env.py:425-448): E independent RandomState(seed[e])
 * streams with state kept in mt_state[E][625]; writes obs [E][S][11]. */
int mlb_legacy_seed(uint32_t *mt_state, const uint32_t *seeds, int32_t E, void *stream);
int mlb_legacy_obs(uint32_t *mt_state, int32_t E, int32_t S, float *obs, void *stream);

/* The reference's whole simulation-mode step for E envs in one launch (env.py:254-262): `_simulate_observation`
 * (env.py:425-448) from each env's own RandomState stream, then -- when `reward` is not NULL -- the reward metric
 * over column `field` of all S servers (every server is active in this mode: n_flow_on >= 5).  One warp per env,
 * chunked warp-parallel MT19937 twist; same state layout as mlb_legacy_seed / mlb_legacy_obs (the three can be
 * mixed on one state).  S <= 256.  `status` (device int, caller-zeroed) receives ST bits on a corrupt stream. */
int mlb_legacy_step(uint32_t *mt_state, int32_t E, int32_t S, int32_t metric, int32_t field, float *obs,
                    double *reward, int32_t *status, void *stream);

/* Batched _normalize_observation (env.py:450-470): running mean / std update and normalisation of
 * n = E*S*11 observation entries in float64, the reference's arithmetic operation for operation
 * (count = obs_count AFTER the increment of env.py:461, >= 1; mean starts at 0, std at 1: env.py:152-153).
 * mean, std are updated in place; out [n] double = (obs - mean) / (std + 1e-8).  Device pointers. */
int mlb_normalize_obs(const float *obs, double *mean, double *std, int64_t count, double *out,
                      int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MARLLB_B200_H */
