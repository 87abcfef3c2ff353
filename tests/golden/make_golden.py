"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The fixtures pin the C oracle (oracle/flow_oracle.c) and, through it, the CUDA
kernels.  Nothing here is imported by the product.
numpy version that produced the committed files: see golden_meta.json.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_flow_env as rfe  # noqa: E402
import ref_import  # noqa: E402

REF = ref_import.REF_ROOT


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


# --------------------------------------------------------------------------
def gold_reservoir():
    """ReservoirSampler(128, seed=42), add(i, i*1e-3), i<1128."""
    res = ref_import.load("reservoir")
    r = res.ReservoirSampler(capacity=128, seed=42)
    accepted = np.zeros(1128, np.uint8)
    for i in range(1128):
        accepted[i] = r.add(float(i), timestamp=i * 1e-3)
    f = r.get_features(0.9, current_time=1.128)
    feats = np.array([f[k] for k in ("mean", "p90", "std", "mean_decay", "p90_decay")])
    assert int(accepted[128:].sum()) == 297 and r.count == 1128
    assert hashlib.sha256(r.values.tobytes()).hexdigest()[:16] == "4ef41e7995d7e99e"
    save("reservoir_seed42", values=r.values, timestamps=r.timestamps, accepted=accepted,
         count=np.int64(r.count), features=feats,
         feature_vector=r.get_feature_vector(0.9, current_time=1.128))
    # a second stream with float32-exact timestamps and several seeds/capacities
    out = {}
    rng = np.random.RandomState(7)
    for cap, seed, n in ((128, 0, 3000), (128, 3, 700), (16, 5, 400), (7, 11, 90), (128, 63, 100)):
        r = res.ReservoirSampler(capacity=cap, seed=seed)
        v = rng.exponential(0.3, n).astype(np.float32)
        t = np.cumsum(rng.exponential(0.01, n)).astype(np.float32)
        acc = np.array([r.add(float(v[i]), timestamp=float(t[i])) for i in range(n)], np.uint8)
        now = float(np.float32(t[-1] + 0.125))
        key = f"c{cap}_s{seed}"
        out[key + "_in_v"], out[key + "_in_t"], out[key + "_acc"] = v, t, acc
        out[key + "_values"], out[key + "_ts"] = r.values.copy(), r.timestamps.copy()
        out[key + "_now"] = np.float64(now)
        out[key + "_fv"] = r.get_feature_vector(0.9, current_time=now)
    save("reservoir_streams", **out)
    # wall-clock timestamps (the reference's default: time.time(), float64, reservoir.py:62,140)
    out = {}
    rng = np.random.RandomState(21)
    for cap, seed, n, t0 in ((128, 9, 900, 1.8e9), (32, 2, 300, 1.7605e9 + 0.123)):
        r = res.ReservoirSampler(capacity=cap, seed=seed)
        v = rng.exponential(0.3, n).astype(np.float32)
        t = t0 + np.cumsum(rng.exponential(0.05, n))            # float64 epoch seconds, ~45 s of samples
        acc = np.array([r.add(float(v[i]), timestamp=float(t[i])) for i in range(n)], np.uint8)
        now = float(t[-1] + 0.25)
        key = f"c{cap}_s{seed}"
        out[key + "_in_v"], out[key + "_in_t"], out[key + "_acc"] = v, t, acc
        out[key + "_values"], out[key + "_ts"] = r.values.copy(), r.timestamps.copy()
        out[key + "_now"] = np.float64(now)
        out[key + "_fv"] = r.get_feature_vector(0.9, current_time=now)
    save("reservoir_wallclock", **out)


def gold_features():
    res = ref_import.load("reservoir")
    rng = np.random.RandomState(0)
    N = 256
    vals = np.zeros((N, 128), np.float32)
    ts = np.zeros((N, 128), np.float32)
    ns = np.zeros(N, np.int32)
    nows = np.zeros(N, np.float32)
    fv = np.zeros((N, 5), np.float32)
    f64 = np.zeros((N, 5), np.float64)
    for i in range(N):
        n = [1, 2, 3, 31, 32, 33, 64, 127, 128][i % 9] if i < 36 else rng.randint(1, 129)
        now = np.float32(rng.uniform(1, 200))
        v = rng.exponential(0.3, n).astype(np.float32)
        t = (now - rng.exponential(2.0, n)).astype(np.float32)
        if i % 5 == 0:
            t[:] = now
        if i % 7 == 0:
            t = (now - np.float32(0.25) * rng.randint(0, 6, n)).astype(np.float32)
        if i % 13 == 0:
            v = np.round(v * 8).astype(np.float32) / 8      # ties in value
        r = res.ReservoirSampler(capacity=128, seed=1)
        r.values[:n] = v
        r.timestamps[:n] = t
        r.count = n
        vals[i, :n], ts[i, :n], ns[i], nows[i] = v, t, n, now
        fv[i] = r.get_feature_vector(0.9, current_time=float(now))
        d = r.get_features(0.9, current_time=float(now))
        f64[i] = [d[k] for k in ("mean", "p90", "std", "mean_decay", "p90_decay")]
    save("features_cases", values=vals, ts=ts, n=ns, now=nows, fv=fv, f64=f64)


def gold_rewards():
    rw = ref_import.load("rewards")
    rng = np.random.RandomState(1)
    metrics = ["jain", "variance", "std", "cv", "max", "min", "product", "range", "gini"]
    cases = [np.array(x, np.float64) for x in
             ([10, 10, 10, 10], [40, 0, 0, 0], [15, 10, 10, 5], [25, 10, 10, 5], [40, 5, 5, 0],
              [0, 0, 0, 0], [7.5], [1e-12, 1e-12])]     # test_rewards.py:31-48,83-119,273-302
    for _ in range(120):
        n = rng.randint(1, 257)
        cases.append(rng.exponential(3.0, n))
    L = max(len(c) for c in cases)
    vals = np.zeros((len(cases), L))
    ns = np.array([len(c) for c in cases], np.int32)
    out = np.zeros((len(cases), len(metrics)))
    for i, c in enumerate(cases):
        vals[i, :len(c)] = c
        for m, name in enumerate(metrics):
            out[i, m] = float(rw.RewardFunction.SUPPORTED_METRICS[name](c))
    assert out[0, 0] == 1.0 and out[1, 0] == 0.25 and abs(out[2, 0] - 8 / 9) < 1e-12
    assert out[1, 1] == -300.0 and out[1, 4] == -40.0 and out[4, 7] == -40.0
    save("rewards_cases", values=vals, n=ns, out=out)


def gold_fair_fn():
    """The original testbed's reward table (src/lb/env.py:73-161).  That module cannot be imported here
    (it needs gym 0.17 and a live /dev/shm region at import), so the `calcul_*` function definitions and
    the `fair_fn` dict are compiled out of the reference source, unmodified, with ast."""
    import ast
    path = os.path.join(ref_import.REF_ROOT, "src", "lb", "env.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body
            if (isinstance(n, ast.FunctionDef) and n.name.startswith("calcul_"))
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "fair_fn")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    ref_fn = ns["fair_fn"]
    names = ["jain", "product", "var", "var_exp", "var_log", "max", "max_exp", "max_log"]
    assert sorted(names) == sorted(ref_fn)
    rng = np.random.RandomState(11)
    cases = [np.array(x, np.float64) for x in
             ([10, 10, 10, 10], [40, 0, 0, 0], [15, 10, 10, 5], [0, 0, 0, 0], [7.5], [1e-4, 2e-4, 1.5e-4])]
    for i in range(90):
        n = rng.randint(1, 257)
        scale = [3.0, 1e-2, 1e-3][i % 3]          # exp(-10000 x) is only non-zero for small x
        cases.append(rng.exponential(scale, n))
    L = max(len(c) for c in cases)
    vals = np.zeros((len(cases), L))
    nn = np.array([len(c) for c in cases], np.int32)
    out = np.zeros((len(cases), len(names)))
    with np.errstate(all="ignore"):
        for i, c in enumerate(cases):
            vals[i, :len(c)] = c
            for m, name in enumerate(names):
                out[i, m] = float(ref_fn[name](c))
    assert out[0, 0] == 1.0 and out[1, 0] == 0.25 and out[3, 0] == 1.0
    save("fair_fn_cases", values=vals, n=nn, out=out, names=np.array(names))


def gold_normalize():
    """_normalize_observation (env.py:450-470): the same seeded reference env with and without
    normalize_obs gives (raw float32 obs, normalised float64 obs) pairs, reset + 12 steps."""
    env = ref_import.load("env")
    S, T = 6, 12
    raw_env = env.LoadBalanceEnv(num_servers=S, step_interval=0.0, seed=7, normalize_obs=False, max_steps=100)
    nrm_env = env.LoadBalanceEnv(num_servers=S, step_interval=0.0, seed=7, normalize_obs=True, max_steps=100)
    raw, nrm = [raw_env.reset()], [nrm_env.reset()]
    act = np.random.RandomState(3).randint(0, 3, (T, S))
    for k in range(T):
        raw.append(raw_env.step(act[k])[0])
        nrm.append(nrm_env.step(act[k])[0])
    raw, nrm = np.stack(raw), np.stack(nrm)
    assert raw.dtype == np.float32 and nrm.dtype == np.float64
    save("normalize_cases", raw=raw, normalized=nrm, mean=nrm_env.obs_mean, std=nrm_env.obs_std,
         count=np.int64(nrm_env.obs_count))


def gold_alias():
    rng = np.random.RandomState(2)
    ps, probs, aliases, ns = [], [], [], []
    for i in range(64):
        n = [1, 2, 4, 16, 32, 64, 256][i % 7] if i < 14 else rng.randint(1, 65)
        w = rng.choice([1.0, 1.5, 2.0], n) if i % 2 else rng.uniform(0.1, 10, n)
        if i == 0:
            w = np.array([0.1, 0.2, 0.3, 0.4])            # test_integration.py:42
        p = w / w.sum()
        t = rfe.build_alias_table_ref(p)
        ps.append(p)
        probs.append(np.array([x[0] for x in t]))
        aliases.append(np.array([x[1] for x in t], np.int32))
        ns.append(len(p))
    L = max(ns)
    pad = lambda xs, dt: np.stack([np.pad(x.astype(dt), (0, L - len(x))) for x in xs])
    save("alias_cases", p=pad(ps, np.float64), prob=pad(probs, np.float64),
         alias=pad(aliases, np.int32), n=np.array(ns, np.int32))


def gold_legacy():
    env = ref_import.load("env")
    # SURVEY App. D
    e = env.LoadBalanceEnv(num_servers=4, step_interval=0.0, seed=42)
    o0 = e.reset()
    assert hashlib.sha256(o0.tobytes()).hexdigest()[:16] == "c0eb394e78c5dcb6"
    o1, r1, d1, info = e.step([0, 1, 2, 1])
    out = {"s4_reset": o0, "s4_step_obs": o1, "s4_step_reward": np.float64(r1),
           "s4_weights": np.array(info["weights"], np.float32)}
    for S, seed, metric in ((16, 7, "jain"), (64, 99, "variance"), (5, 3, "gini")):
        e = env.LoadBalanceEnv(num_servers=S, step_interval=0.0, seed=seed, reward_metric=metric,
                               max_steps=6)
        obs = [e.reset()]
        rew, dones = [], []
        rng = np.random.RandomState(seed)
        for k in range(6):
            o, r, d, _ = e.step(rng.randint(0, 3, S))
            obs.append(o)
            rew.append(r)
            dones.append(d)
        out[f"s{S}_obs"] = np.stack(obs)
        out[f"s{S}_rew"] = np.array(rew, np.float64)
        out[f"s{S}_done"] = np.array(dones, np.uint8)
    # normalisation path (env.py:450-470)
    e = env.LoadBalanceEnv(num_servers=4, step_interval=0.0, seed=5, normalize_obs=True)
    norm = [e.reset()] + [e.step([1, 1, 1, 1])[0] for _ in range(4)]
    out["norm_obs"] = np.stack(norm)
    save("legacy_env", **out)


# --------------------------------------------------------------------------
def load_trace(path, horizon):
    t, w = [], []
    with open(path) as f:
        next(f)
        for line in f:
            a, b = line.rstrip("\n").split("\t")
            if float(a) >= horizon:
                break
            t.append(float(a))
            w.append(int(b.split("n=")[1]))
    return np.array(t, np.float64), np.array(w, np.float64)


def run_flow(name, A, Sa, steps, arrivals, speeds, actions, **kw):
    S = A * Sa
    env = rfe.RefFlowEnv(A, Sa, speeds, arrivals, max_steps=steps, **kw)
    env.reset()
    obs = np.zeros((steps, S, 11), np.float32)
    rew = np.zeros(steps, np.float64)
    done = np.zeros(steps, np.uint8)
    n_on = np.zeros((steps, S), np.int32)
    assign = [[] for _ in range(A)]
    for k in range(steps):
        o, r, d, info = env.step(actions[k])
        obs[k], rew[k], done[k] = o, r, d
        n_on[k] = env.psf.n_flow_on
        for i in range(A):
            assign[i] += info["assign"][i]
    d = env.dump()
    out = {"obs": obs, "reward": rew, "done": done, "n_flow_on": n_on,
           "speeds": np.asarray(speeds, np.float32), "actions": np.asarray(actions),
           "res_values": d["res_values"], "res_ts": d["res_ts"], "res_count": d["res_count"],
           "dropped": d["dropped"],
           "cfg": np.array(json.dumps(dict(A=A, Sa=Sa, steps=steps, **kw)))}
    for i in range(A):
        out[f"arr_time_{i}"] = np.asarray(arrivals[i]["time"], np.float32)
        out[f"arr_work_{i}"] = np.asarray(arrivals[i]["work"], np.float32)
        if "bucket" in arrivals[i]:
            out[f"arr_bucket_{i}"] = np.asarray(arrivals[i]["bucket"], np.int32)
            out[f"arr_u_{i}"] = np.asarray(arrivals[i]["u"], np.float32)
        out[f"assign_{i}"] = np.array(assign[i], np.int16)
    save(name, **out)


def synth(rng, A, rate, horizon, mean_work, Sa):
    arr = []
    for _ in range(A):
        n = int(rate * horizon * 1.3) + 10
        t = np.cumsum(rng.exponential(1.0 / rate, n))      # training_pipeline.py:141-155
        t = t[t < horizon].astype(np.float32)
        arr.append({"time": t, "work": rng.exponential(mean_work, len(t)).astype(np.float32),
                    "bucket": rng.randint(0, Sa, len(t)).astype(np.int32),
                    "u": rng.random_sample(len(t)).astype(np.float32)})
    return arr


def gold_flow(only=None):
    """`only`: optional list of fixture names to (re)generate."""
    # C1: unittest topology, 1 LB x 4 servers, data/trace rate_500, first 60 s
    steps = 240
    t, n = load_trace(os.path.join(REF, "data/trace/poisson_for_loop/rate_500.csv"), steps * 0.25)
    work = (n / 1e6).astype(np.float32)                     # Mloops
    rate = len(t) / (steps * 0.25)
    v0 = rate * float(work.mean()) / (6 * 0.8)              # rho ~= 0.8 over speeds [1,1,2,2]*v0
    speeds = (np.array([1, 1, 2, 2]) * v0).astype(np.float32)
    actions = np.random.RandomState(0).randint(0, 3, (steps, 4)).astype(np.int32)
    if only is None or "flow_c1_trace" in only:
        run_flow("flow_c1_trace", 1, 4, steps, [{"time": t.astype(np.float32), "work": work}], speeds, actions)

    def case(name, A, Sa, steps, rate, mean_work, speeds, seed, **kw):
        if only is not None and name not in only:
            return
        rng = np.random.RandomState(seed)
        arr = synth(rng, A, rate, steps * 0.25 + 1, mean_work, Sa)
        S = A * Sa
        if kw.get("action_type", "discrete") == "discrete":
            act = rng.randint(0, 3, (steps, S)).astype(np.int32)
        else:
            act = rng.uniform(-1, 12, (steps, S)).astype(np.float32)
        run_flow(name, A, Sa, steps, arr, speeds, act, **kw)

    case("flow_sed_a2s3", 2, 3, 40, 60, 0.04, [1, 2, 1, 2, 1, 2], 1)
    case("flow_lsq_s5", 1, 5, 40, 80, 0.05, [1, 1, 1, 2, 2], 2, policy="lsq")
    case("flow_alias_a2s4", 2, 4, 40, 50, 0.05, [1, 2] * 4, 3, policy="alias")
    case("flow_cont_s4", 1, 4, 40, 100, 0.03, [1, 1, 2, 2], 4, action_type="continuous")
    case("flow_drops_q4", 1, 3, 40, 100, 0.06, [1, 1, 1], 5, queue_cap=4)
    case("flow_k8_s4", 1, 4, 60, 100, 0.03, [1, 1, 2, 2], 6, reservoir_k=8)
    case("flow_s40_var", 1, 40, 30, 160, 0.4, [1, 2] * 20, 8, reward_metric="variance", reward_field="fct_mean")
    case("flow_a4s16_gini", 4, 16, 24, 64, 0.4, [1, 2] * 32, 9, reward_metric="gini")
    # power of two choices (node.c:408-417, 433-441)
    case("flow_sed2_a2s5", 2, 5, 40, 70, 0.05, [1, 2, 1, 2, 2] * 2, 10, policy="sed2")
    case("flow_lsq2_s36", 1, 36, 30, 200, 0.25, [1, 2] * 18, 12, policy="lsq2", reward_metric="max",
         reward_field="fct_mean")


if __name__ == "__main__":
    gold_reservoir()
    gold_features()
    gold_rewards()
    gold_fair_fn()
    gold_normalize()
    gold_alias()
    gold_legacy()
    gold_flow(sys.argv[1:] or None)
    meta = {"numpy": np.__version__, "python": sys.version.split()[0],
            "reference": REF, "generator": "tests/golden/make_golden.py"}
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(meta)
