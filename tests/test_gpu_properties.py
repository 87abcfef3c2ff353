"""GPU: size-independent properties of the env step at sizes the oracle cannot reach in seconds
(thousands of envs x 64 servers, 128-slot reservoirs, BASELINE.json's C5 shape)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E, S, STEPS = 4096, 64, 48


def _make(feature_cache=1, env_id_base=0, E_=E, **kw):
    from marllb_b200 import VecLoadBalanceEnv
    env = VecLoadBalanceEnv(E_, num_servers=S, num_agents=1, max_steps=STEPS, feature_cache=feature_cache,
                            env_id_base=env_id_base, action_dtype="uint8", **kw)
    env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0))
    env.gen_poisson(128.0, 0.6, STEPS * 0.25 + 1.0, seed=7)
    env.reset()
    return env


def test_flow_conservation_and_feature_cache_idempotence():
    """Every processed flow is finished (one fct sample), still active, or dropped; the
    flow_duration reservoir saw one sample per active flow per step; skipping untouched
    reservoirs (mode 1, incremental ranks) changes nothing versus recomputing everything
    from scratch (mode 0) or re-sorting the touched ones (mode 2)."""
    import torch
    envs = [_make(fc) for fc in (1, 0, 2)]
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    active_sum = torch.zeros((E, S), dtype=torch.float64, device="cuda")
    for k in range(STEPS):
        act = torch.randint(0, 3, (E, S), generator=g, device="cuda", dtype=torch.uint8)
        outs = [env.step(act) for env in envs]
        obs, rew, done = outs[0]
        for o2, r2, d2 in outs[1:]:
            # counts and order statistics (p90, p90_decay: elements / exact lerps of the reservoir) and the
            # reward, which reads column 10: bit-identical.  mean, std, mean_decay are float32 sums whose
            # reduction tree depends on the path (mode 1 keeps 4 slots per lane for every fill level, the
            # re-sorting modes use 1 / 2 / 4): equal within a few ulp.
            exact = [0, 2, 5, 7, 10]
            assert torch.equal(obs[..., exact], o2[..., exact]) and torch.equal(rew, r2) and torch.equal(done, d2)
            assert torch.allclose(obs[..., [1, 3, 4, 6, 8, 9]], o2[..., [1, 3, 4, 6, 8, 9]], rtol=2e-6, atol=1e-9)
        active_sum += obs[..., 0].double()
        assert bool((obs >= 0).all()) and bool(torch.isfinite(obs).all())
        assert bool(((rew >= 1.0 / S - 1e-12) & (rew <= 1.0 + 1e-12) | (rew == 0)).all())    # Jain in [1/n, 1]
        # order relations inside one reservoir: mean <= max-ish, p90 >= 0, std >= 0
        assert bool((obs[..., 3] >= 0).all()) and bool((obs[..., 8] >= 0).all())
    for env in envs:
        env.check_status()
    env = envs[0]
    cnt = env.get_state("res_count").astype(np.int64)          # (E, 2, S)
    n_on = env.get_state("n_flow_on").astype(np.int64)
    drp = env.get_state("dropped").astype(np.int64)
    cur = env.get_state("arr_cursor").astype(np.int64)[:, 0]
    assert np.array_equal(cur, cnt[:, 0].sum(1) + n_on.sum(1) + drp.sum(1))
    assert np.array_equal(cnt[:, 1], active_sum.cpu().numpy().astype(np.int64))
    assert bool(done.all()) and int(env.get_state("step")[0]) == STEPS
    # ranks are a permutation of 0..n-1 for every touched reservoir
    vals = env.get_state("res_values")
    assert np.isfinite(vals).all() and (vals >= 0).all()
    for e in envs:
        e.close()


def test_results_do_not_depend_on_the_shard():
    """Global env e computed as part of a 64-env block equals the same env computed alone with
    env_id_base=e (the multi-GPU split only moves envs between ranks)."""
    import torch
    big = _make(E_=64)
    ids = [0, 17, 63]
    small = [_make(E_=1, env_id_base=i) for i in ids]
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    for k in range(16):
        act = torch.randint(0, 3, (64, S), generator=g, device="cuda", dtype=torch.uint8)
        ob, rb, _ = big.step(act)
        for i, env in zip(ids, small):
            o, r, _ = env.step(act[i:i + 1].clone())
            assert torch.equal(o[0], ob[i]) and torch.equal(r[0], rb[i])
    for env in small + [big]:
        env.close()


def test_determinism_and_episode_replay():
    """reset() rewinds the episode: the same actions give bit-identical trajectories."""
    import torch
    env = _make(E_=256)
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    acts = [torch.randint(0, 3, (256, S), generator=g, device="cuda", dtype=torch.uint8) for _ in range(12)]
    first = []
    for a in acts:
        o, r, _ = env.step(a)
        first.append((o.clone(), r.clone()))
    env.reset()
    for a, (o1, r1) in zip(acts, first):
        o, r, _ = env.step(a)
        assert torch.equal(o, o1) and torch.equal(r, r1)
    env.close()


def test_c2_full_size_long_episode_modes_agree_and_flows_are_conserved():
    """BASELINE config C2 at its full size (4096 envs x 16 servers, K = 128, 32 flows/s, 1000 steps): deep in the
    episode every touched reservoir is full and takes the half-warp pair kernels; their output must keep
    agreeing with the re-sorting modes (order statistics and rewards exactly, float32 sums to 2e-6), and
    the flow accounting must close at the end."""
    import torch
    from marllb_b200 import VecLoadBalanceEnv
    E2, S2, T2 = 4096, 16, 1000
    envs = []
    for fc in (1, 2):
        env = VecLoadBalanceEnv(E2, num_servers=S2, max_steps=T2, feature_cache=fc, action_dtype="uint8")
        env.set_speeds(np.where(np.arange(S2) % 2 == 0, 1.0, 2.0))
        env.gen_poisson(32.0, 0.8 * 1.5 * S2 / 32.0, T2 * 0.25 + 1.0, seed=21)
        env.reset()
        envs.append(env)
    g = torch.Generator(device="cuda")
    g.manual_seed(8)
    exact, approx = [0, 2, 5, 7, 10], [1, 3, 4, 6, 8, 9]
    active_sum = torch.zeros((E2, S2), dtype=torch.float64, device="cuda")
    for k in range(T2):
        act = torch.randint(0, 3, (E2, S2), generator=g, device="cuda", dtype=torch.uint8)
        (o1, r1, d1), (o2, r2, d2) = [env.step(act) for env in envs]
        active_sum += o1[..., 0].double()
        if k % 25 == 24 or k > T2 - 20:
            assert torch.equal(o1[..., exact], o2[..., exact]) and torch.equal(r1, r2), k
            assert torch.allclose(o1[..., approx], o2[..., approx], rtol=2e-6, atol=1e-9), k
    assert bool(d1.all())
    env = envs[0]
    env.check_status()
    cnt = env.get_state("res_count").astype(np.int64)
    n_on, drp = env.get_state("n_flow_on").astype(np.int64), env.get_state("dropped").astype(np.int64)
    cur = env.get_state("arr_cursor").astype(np.int64)[:, 0]
    assert np.array_equal(cur, cnt[:, 0].sum(1) + n_on.sum(1) + drp.sum(1))        # every flow: done, active or dropped
    assert np.array_equal(cnt[:, 1], active_sum.cpu().numpy().astype(np.int64))      # one duration sample per active flow per step
    assert (cnt[:, 0] > 128).mean() > 0.9                                            # reservoirs really are in the replacement regime
    for e in envs:
        e.close()


def test_device_poisson_generator_statistics():
    """mlb_gen_poisson follows `_generate_poisson_trace` (training_pipeline.py:141-155): exponential gaps at `rate`,
    arrivals kept while < horizon, exponential work; streams are time-sorted, independent per (env, agent) and a
    function of the seed only."""
    from marllb_b200 import VecLoadBalanceEnv
    E, rate, mean_work, horizon = 512, 40.0, 0.3, 25.0
    env = VecLoadBalanceEnv(E, num_servers=8, num_agents=2, policy="alias")
    env.gen_poisson(rate, mean_work, horizon, seed=7)
    counts, gaps, works, buckets, us = [], [], [], [], []
    for e in range(0, E, 8):
        for a in range(2):
            s = env.get_arrivals(e, a)
            t = s["time"]
            assert t.dtype == np.float32 and np.all(np.diff(t) >= 0) and (len(t) == 0 or t[-1] < horizon)
            counts.append(len(t)); gaps.append(np.diff(t.astype(np.float64))); works.append(s["work"])
            buckets.append(s["bucket"]); us.append(s["u"])
    counts = np.array(counts, np.float64)
    n = len(counts)
    assert abs(counts.mean() - rate * horizon) < 4 * np.sqrt(rate * horizon / n)          # Poisson mean
    assert 0.7 < counts.var() / (rate * horizon) < 1.4                                    # ... and variance
    g = np.concatenate(gaps)
    assert abs(g.mean() * rate - 1) < 0.02 and abs(g.std() / g.mean() - 1) < 0.03         # exponential gaps (CV = 1)
    w = np.concatenate(works)
    assert abs(w.mean() / mean_work - 1) < 0.02 and w.min() >= 0
    b, u = np.concatenate(buckets), np.concatenate(us)
    assert b.min() == 0 and b.max() == 7 and abs(np.bincount(b, minlength=8) / len(b) - 0.125).max() < 0.01
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    first = env.get_arrivals(0, 0)["time"].copy()
    assert not np.array_equal(first[:16], env.get_arrivals(0, 1)["time"][:16])            # streams differ
    env.gen_poisson(rate, mean_work, horizon, seed=7)
    assert np.array_equal(first, env.get_arrivals(0, 0)["time"])                          # same seed, same stream
    env.gen_poisson(rate, mean_work, horizon, seed=8)
    assert not np.array_equal(first[:16], env.get_arrivals(0, 0)["time"][:16])
    env.close()
