"""GPU: stand-alone reservoir / reward kernels against the reference-generated fixtures
(SURVEY App. D goldens, reference tests' known answers)."""
import hashlib

import numpy as np
import pytest

from conftest import OBS_RTOL, load_golden

pytestmark = pytest.mark.gpu


def test_reservoir_appendix_d_golden():
    from marllb_b200 import ReservoirSampler
    g = load_golden("reservoir_seed42")
    r = ReservoirSampler(capacity=128, seed=42)
    # timestamps i*1e-3 are stored as float32 on the device; slots/accept flags must still match
    acc = r.add_many(np.arange(1128, dtype=np.float32), np.arange(1128) * 1e-3)
    assert np.array_equal(acc.astype(np.uint8), g["accepted"]) and int(acc[128:].sum()) == 297
    assert r.count == 1128 and r.get_size() == 128 and r.is_full()
    v = r.values
    assert np.array_equal(v, g["values"])
    assert v[:16].tolist() == [1087, 378, 1017, 1067, 1117, 1064, 1016, 993, 679, 1096, 10, 768, 1100, 195, 261, 588]
    assert hashlib.sha256(v.tobytes()).hexdigest()[:16] == "4ef41e7995d7e99e"
    f = r.get_feature_vector(0.9, current_time=1.128)
    np.testing.assert_allclose(f, g["feature_vector"], rtol=OBS_RTOL)
    assert f[4] == 1041.0 and abs(f[1] - 1040.30005) < 1e-3


def test_single_adds_follow_reference_semantics():
    from marllb_b200 import ReservoirSampler
    r = ReservoirSampler(capacity=128, seed=42)
    for i in range(128):                       # test_reservoir.py:40-47: fill phase always accepts
        assert r.add(float(i), timestamp=float(i))
    assert r.is_full() and len(r) == 128
    g = load_golden("reservoir_seed42")
    for i in range(128, 160):
        assert r.add(float(i), timestamp=float(i)) == bool(g["accepted"][i])
    r.reset()                                  # test_reservoir.py:133-146
    assert r.count == 0 and len(r) == 0 and not r.values.any()
    assert r.get_features() == {k: 0.0 for k in ("mean", "p90", "std", "mean_decay", "p90_decay")}


def test_reservoir_streams():
    from marllb_b200.reservoir import BatchedReservoirs
    g = load_golden("reservoir_streams")
    keys = sorted({k.rsplit("_in_v", 1)[0] for k in g if k.endswith("_in_v")})
    for key in keys:
        cap, seed = int(key.split("_")[0][1:]), int(key.split("_")[1][1:])
        b = BatchedReservoirs(1, cap, seeds=[seed])
        acc = b.add(g[key + "_in_v"][None], g[key + "_in_t"][None])
        b.check_status()
        assert np.array_equal(acc[0].cpu().numpy(), g[key + "_acc"])
        assert np.array_equal(b.values[0, :cap].cpu().numpy(), g[key + "_values"])
        assert np.array_equal(b.timestamps[0, :cap].cpu().numpy().astype(np.float64), g[key + "_ts"])
        f = b.features(0.9, float(g[key + "_now"]))[0].cpu().numpy()
        np.testing.assert_allclose(f, g[key + "_fv"], rtol=OBS_RTOL)


def test_reservoir_wallclock_timestamps():
    """The reference's default timestamps are float64 time.time() values (reservoir.py:42,62,140): ~1.8e9, where
    float32 resolves 128 s.  The facade rebases them on a float64 host-side epoch, so decay weights still see the
    true differences between samples (fixture: unmodified reference fed epoch-scale timestamps)."""
    from marllb_b200.reservoir import ReservoirSampler
    g = load_golden("reservoir_wallclock")
    keys = sorted({k.rsplit("_in_v", 1)[0] for k in g if k.endswith("_in_v")})
    assert keys
    for key in keys:
        cap, seed = int(key.split("_")[0][1:]), int(key.split("_")[1][1:])
        s = ReservoirSampler(capacity=cap, seed=seed)
        acc = s.add_many(g[key + "_in_v"], g[key + "_in_t"])
        assert np.array_equal(acc.astype(np.uint8), g[key + "_acc"])
        assert np.array_equal(s.values, g[key + "_values"])
        # rebased float32 seconds: < 1e-5 s over the ~45 s the samples span (128 s without the epoch)
        np.testing.assert_allclose(s.timestamps, g[key + "_ts"], rtol=0, atol=1e-5)
        f = s.get_feature_vector(0.9, float(g[key + "_now"]))
        np.testing.assert_allclose(f, g[key + "_fv"], rtol=OBS_RTOL)


def test_features_cases_batched():
    import torch
    from marllb_b200.reservoir import BatchedReservoirs
    g = load_golden("features_cases")
    N = len(g["n"])
    b = BatchedReservoirs(N, 128)
    b.values.copy_(torch.as_tensor(g["values"]))
    b.timestamps.copy_(torch.as_tensor(g["ts"]))
    b.count.copy_(torch.as_tensor(g["n"]))
    f = b.features(0.9, torch.as_tensor(g["now"])).cpu().numpy()
    np.testing.assert_allclose(f[:, [0, 2, 3]], g["fv"][:, [0, 2, 3]], rtol=OBS_RTOL)   # mean, std, mean_decay
    assert np.array_equal(f[:, 1], g["fv"][:, 1])       # p90: exact order statistics + same f32 lerp
    assert np.array_equal(f[:, 4], g["fv"][:, 4])       # p90_decay: an element of the reservoir


def test_multi_metric_and_per_server_state():
    from marllb_b200 import MultiMetricReservoir, PerServerFeatures
    m = MultiMetricReservoir(metrics=["fct", "flow_duration"], capacity=128, seed=3)
    rng = np.random.RandomState(0)
    for i in range(40):
        m.add("fct", float(rng.exponential(0.1)), timestamp=i * 0.01)
        m.add("flow_duration", float(rng.uniform(0.01, 1.0)), timestamp=i * 0.01)
    with pytest.raises(ValueError):
        m.add("nope", 1.0)                              # test_reservoir.py:195-198
    v = m.get_feature_vector(0.9, current_time=0.5)
    assert v.shape == (10,) and v.dtype == np.float32   # test_reservoir.py:217-226
    p = PerServerFeatures(2)
    p.update_flow_count(0, 7)
    st = p.get_state_vector([v, v], active_servers=[0])
    assert st.shape == (2, 11) and st[0, 0] == 7 and not st[1].any()


def test_rewards_cases_and_reference_literals():
    import torch
    from marllb_b200 import _lib, rewards
    g = load_golden("rewards_cases")
    vals = torch.as_tensor(g["values"]).cuda()
    n = torch.as_tensor(g["n"]).cuda()
    for m, name in enumerate(list(_lib.METRICS)[:g["out"].shape[1]]):
        out = rewards.reward_metric_batch(name, vals, n).cpu().numpy()
        np.testing.assert_allclose(out, g["out"][:, m], rtol=1e-12, atol=1e-300, err_msg=name)
    # tests/test_rewards.py:31-48,83-119
    assert rewards.jain_fairness([10, 10, 10, 10]) == pytest.approx(1.0)
    assert rewards.jain_fairness([40, 0, 0, 0]) == pytest.approx(0.25)
    assert rewards.jain_fairness([15, 10, 10, 5]) == pytest.approx(0.888888, abs=1e-5)
    assert rewards.variance_fairness([40, 0, 0, 0]) == pytest.approx(-300.0)
    assert rewards.max_min_fairness([40, 0, 0, 0]) == pytest.approx(-40.0)
    assert rewards.range_fairness([40, 5, 5, 0]) == pytest.approx(-40.0)
    assert rewards.jain_fairness([]) == 1.0 and rewards.variance_fairness([]) == 0.0
    rf = rewards.RewardFunction("jain", "fct_mean")      # tests/test_rewards.py:159-182
    obs = {"active_servers": [0, 1, 2, 3], "server_stats": {i: {"fct_mean": v} for i, v in enumerate([10, 12, 11, 10])}}
    assert 0.99 < rf.compute(obs) <= 1.0 and rf(obs) == rf.compute(obs)


def test_fair_fn_table_of_the_original_testbed():
    """src/lb/env.py:73-161 (`fair_fn`, `calcul_fair`): fixtures from the reference's own functions."""
    import torch
    from marllb_b200 import rewards
    g = load_golden("fair_fn_cases")
    vals = torch.as_tensor(g["values"]).cuda()
    n = torch.as_tensor(g["n"]).cuda()
    here = {"jain": "fair_jain", "product": "fair_product"}
    for m, name in enumerate(g["names"]):
        name = str(name)
        out = rewards.reward_metric_batch(here.get(name, name), vals, n).cpu().numpy()
        np.testing.assert_allclose(out, g["out"][:, m], rtol=1e-11, atol=1e-300, err_msg=name)
    assert set(rewards.fair_fn) == set(str(x) for x in g["names"])
    assert rewards.calcul_fair([40, 0, 0, 0], "jain") == pytest.approx(0.25)
    assert rewards.calcul_fair([0, 0, 0, 0], "jain") == 1.0                 # src/lb/env.py:82-85
    assert rewards.calcul_fair([1e-4, 2e-4], "max_exp") == pytest.approx(np.exp(-2.0))
    assert rewards.RewardFunction("var_log", "fct_mean").metric == "var_log"


def test_env_step_with_fair_fn_reward():
    """The in-kernel reward of the flow-level step with a fair_fn metric equals the oracle's."""
    import flow_oracle as fo
    from marllb_b200 import VecLoadBalanceEnv
    rng = np.random.RandomState(5)
    E, S, steps = 3, 6, 12
    speeds = np.array([1, 2, 1, 2, 1, 2], np.float32)
    for metric in ("fair_jain", "fair_product", "var_exp", "var_log", "max_exp", "max_log"):
        streams = []
        for _ in range(E):
            t = np.cumsum(rng.exponential(1 / 80.0, 400))
            t = t[t < steps * 0.25 + 0.5].astype(np.float32)
            streams.append([{"time": t, "work": rng.exponential(0.002, len(t)).astype(np.float32)}])
        env = VecLoadBalanceEnv(E, num_servers=S, max_steps=steps, reward_metric=metric, reward_field=1)
        env.set_speeds(speeds)
        env.load_arrivals([s[0] for s in streams])
        env.reset()
        ora = [fo.FlowEnv(1, S, speeds, streams[e], max_steps=steps, reward_metric=metric, reward_field=1)
               for e in range(E)]
        for _ in range(steps):
            act = rng.randint(0, 3, (E, S)).astype(np.int32)
            _, rew, _ = env.step(act)
            _, r_ref, _, _ = fo.step_batch(ora, act)
            # exp(-10000 x) turns a relative feature error of 1e-6 into 1e-2 * x
            rtol = 1e-3 if metric.endswith("_exp") else 1e-5
            np.testing.assert_allclose(rew.cpu().numpy(), r_ref, rtol=rtol, atol=1e-300, err_msg=metric)
        env.close()


def test_reference_statistical_properties_of_the_sampler():
    """The reference's statistical tests restated on the batched device sampler (tests/test_reservoir.py:243-316):
    chi-square uniformity of Algorithm R over 500 seeded trials (capacity 100, stream of 1000 distinct items, one
    trial per reservoir, all trials in one launch) and convergence of the sample mean."""
    import torch
    from marllb_b200.reservoir import BatchedReservoirs
    trials, cap, n = 500, 100, 1000
    b = BatchedReservoirs(trials, cap, seeds=np.arange(trials))            # ReservoirSampler(capacity, seed=trial)
    stream = np.tile(np.arange(n, dtype=np.float32), (trials, 1))
    b.add(stream, stream * 1e-3)
    b.check_status()
    vals = b.values[:, :cap].cpu().numpy().astype(np.int64)
    counts = np.bincount(vals.reshape(-1), minlength=n).astype(np.float64)
    expected = trials * cap / n
    chi_square = float(((counts - expected) ** 2 / expected).sum())
    assert chi_square < 1100, chi_square                                   # :286-287 (dof 999, alpha 0.05: ~1073)
    for row in vals[:16]:
        assert len(set(row.tolist())) == cap                               # a reservoir never holds an item twice
    # :289-316: mean of a sample of N(100, 15) values stays close to the stream mean
    rng = np.random.RandomState(42)
    data = rng.normal(100.0, 15.0, (1, 10000)).astype(np.float32)
    s = BatchedReservoirs(1, 128, seeds=[42])
    s.add(data, np.arange(10000, dtype=np.float32)[None] * 1e-3)
    mean = float(s.features(0.9, 10.0)[0, 0])
    assert abs(mean - float(data.mean())) < 5.0
