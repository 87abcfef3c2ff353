"""GPU: the drop-in LoadBalanceEnv / MultiAgentLoadBalanceEnv.  Legacy mode must reproduce the
unmodified reference bit-for-bit for the same seed (fixtures from the reference); the remaining
tests restate the reference's own tests/test_env.py assertions against the new class."""
import hashlib

import numpy as np
import pytest

from conftest import OBS_RTOL, flow_case, load_golden

pytestmark = pytest.mark.gpu


def test_legacy_mode_is_bit_exact_with_reference():
    from marllb_b200 import LoadBalanceEnv
    g = load_golden("legacy_env")
    env = LoadBalanceEnv(num_servers=4, step_interval=0.0, seed=42)
    assert env.mode == "legacy"
    o0 = env.reset()
    assert o0.dtype == np.float32 and np.array_equal(o0, g["s4_reset"])
    assert hashlib.sha256(o0.tobytes()).hexdigest()[:16] == "c0eb394e78c5dcb6"
    o1, r1, d1, info = env.step([0, 1, 2, 1])
    assert np.array_equal(o1, g["s4_step_obs"]) and d1 is False or d1 == False
    assert r1 == pytest.approx(float(g["s4_step_reward"]), rel=1e-12) == pytest.approx(0.946240194901, rel=1e-10)
    assert info["weights"] == [1.0, 1.5, 2.0, 1.5] and info["step"] == 1 and info["active_servers"] == [0, 1, 2, 3]
    for S, seed, metric in ((16, 7, "jain"), (64, 99, "variance"), (5, 3, "gini")):
        env = LoadBalanceEnv(num_servers=S, step_interval=0.0, seed=seed, reward_metric=metric, max_steps=6)
        assert np.array_equal(env.reset(), g[f"s{S}_obs"][0])
        rng = np.random.RandomState(seed)
        for k in range(6):
            o, r, d, info = env.step(rng.randint(0, 3, S))
            assert np.array_equal(o, g[f"s{S}_obs"][k + 1])
            assert r == pytest.approx(g[f"s{S}_rew"][k], rel=1e-12)
            assert d == bool(g[f"s{S}_done"][k])
        assert info["episode"]["l"] == 6


def test_normalize_obs_matches_reference():
    from marllb_b200 import LoadBalanceEnv
    g = load_golden("legacy_env")
    env = LoadBalanceEnv(num_servers=4, step_interval=0.0, seed=5, normalize_obs=True)
    outs = [env.reset()] + [env.step([1, 1, 1, 1])[0] for _ in range(4)]
    np.testing.assert_allclose(np.stack(outs), g["norm_obs"], rtol=1e-12, atol=1e-12)
    assert env.obs_count == 5


def test_reference_test_env_assertions():
    """tests/test_env.py:30-123,129-175,281-310,341-364 restated."""
    from marllb_b200 import LoadBalanceEnv
    env = LoadBalanceEnv(num_servers=4, action_type="discrete", max_steps=100, use_shm=False, seed=42)
    assert env.num_servers == 4 and env.action_type == "discrete" and env.max_steps == 100 and not env.use_shm
    assert env.observation_space.shape == (4, 11)
    assert np.all(env.observation_space.low == 0) and np.all(env.observation_space.high == np.inf)
    assert len(env.action_space.nvec) == 4 and np.all(env.action_space.nvec == 3)
    c = LoadBalanceEnv(num_servers=4, action_type="continuous", max_steps=100)
    assert c.action_space.shape == (4,) and np.all(c.action_space.low == np.float32(0.1)) and np.all(c.action_space.high == 10.0)
    obs = env.reset()
    assert obs.shape == (4, 11) and np.all(np.isfinite(obs)) and env.current_step == 0
    nobs, reward, done, info = env.step(env.action_space.sample())
    assert nobs.shape == (4, 11) and isinstance(reward, float) and isinstance(done, (bool, np.bool_)) and isinstance(info, dict)
    assert {"step", "weights", "active_servers", "episode_return"} <= set(info)
    short = LoadBalanceEnv(num_servers=4, max_steps=5, seed=1)
    short.reset()
    for i in range(5):
        _, r, done, info = short.step(short.action_space.sample())
        assert done == (i == 4) and 0.25 <= r <= 1.0                       # test_env.py:100-112,281-296
    assert info["episode"]["l"] == 5 and info["episode"]["r"] == pytest.approx(short.episode_return)
    a, b = LoadBalanceEnv(num_servers=4, seed=123), LoadBalanceEnv(num_servers=4, seed=123)
    assert np.array_equal(a.reset(), b.reset())                            # test_env.py:114-123
    a.seed(7); b.seed(7)
    assert np.array_equal(a._simulate_observation(), b._simulate_observation())
    w = env._action_to_weights(np.array([0, 1, 2, 1]))
    assert w.dtype == np.float32 and w.tolist() == [1.0, 1.5, 2.0, 1.5]
    assert np.allclose(c._action_to_weights(np.array([0.01, 15.0, 5.0, 1.0])), [0.1, 10.0, 5.0, 1.0])
    d = env._array_to_dict(nobs)
    assert np.array_equal(env._dict_to_array(d), nobs)                     # round trip, test_env.py:254-275
    z = np.zeros((4, 11), np.float32); z[2, 3] = 1.0
    assert env._array_to_dict(z)["active_servers"] == [2]
    with pytest.raises(ValueError):
        LoadBalanceEnv(action_type="bogus")
    with pytest.raises(ValueError):
        LoadBalanceEnv(use_shm=True)
    env.render()


def test_flow_mode_single_env_matches_fixture():
    from marllb_b200 import LoadBalanceEnv
    cfg, arrivals, g = flow_case("flow_c1_trace")          # config C1: unittest topology, data/trace
    env = LoadBalanceEnv(num_servers=4, max_steps=cfg["steps"], arrivals=arrivals[0], server_speeds=g["speeds"])
    assert env.mode == "flow"
    assert not env.reset().any()
    for k in range(cfg["steps"]):
        o, r, d, info = env.step(g["actions"][k])
        np.testing.assert_allclose(o, g["obs"][k], rtol=OBS_RTOL, atol=1e-7)
        assert r == pytest.approx(g["reward"][k], rel=1e-9) and d == bool(g["done"][k])
    assert np.array_equal(env.assignments(), g["assign_0"].astype(np.int32))
    env.close()


def test_multi_agent_env_reference_shapes():
    """problem-05-qmix/src/multi_agent_env.py __main__ self-test + SURVEY App. C #2 measurements."""
    from marllb_b200 import MultiAgentLoadBalanceEnv
    env = MultiAgentLoadBalanceEnv(num_agents=4, servers_per_agent=4, action_type="continuous",
                                   reward_metric="jain", max_steps=20, global_reward=True, seed=0)
    assert env.total_servers == 16 and env.obs_dim == 20 and env.state_dim == 74
    obs = env.reset()
    assert len(obs) == 4 and all(o.shape == (128,) for o in obs)            # 4*Sa + 7*S_tot
    acts = [np.random.rand(4) for _ in range(4)]
    obs, rewards, done, info = env.step(acts)
    assert len(rewards) == 4 and len(set(rewards)) == 1 and not done
    assert env.get_state().shape == (74,)
    obs, rewards, done, info = env.step([1, 2, 0, 1])                        # integer actions are accepted
    clean = MultiAgentLoadBalanceEnv(num_agents=2, servers_per_agent=3, strict_reference=False, seed=0)
    o = clean.reset()
    assert o[0].shape == (33,) and clean.obs_dim == 33 and clean.get_state().shape == (6 * 11 + 10,)


def test_batched_normalize_observation_bit_exact():
    """_normalize_observation (env.py:450-470) for all envs in one kernel: the reference's float64 running
    statistics, bit for bit, against (raw, normalised) pairs produced by the reference env itself."""
    import torch
    from marllb_b200 import VecLoadBalanceEnv
    g = load_golden("normalize_cases")
    raw, ref = g["raw"], g["normalized"]
    T, S = raw.shape[0], raw.shape[1]
    E = 5
    env = VecLoadBalanceEnv(E, num_servers=S, normalize_obs=True)
    scale = np.arange(1, E + 1, dtype=np.float32).reshape(E, 1, 1)         # env e sees e+1 times the obs
    mean, std = np.zeros((E, S, 11)), np.ones((E, S, 11))
    for k in range(T):
        x = raw[k][None] * scale
        out = env._normalize_observation(torch.as_tensor(x).cuda()).cpu().numpy()
        # the reference formula, env by env
        n = k + 1
        delta = x - mean
        mean = mean + delta / n
        delta2 = x - mean
        std = np.sqrt(np.maximum((std ** 2 * (n - 1) + delta * delta2) / n, 1e-8))
        assert np.array_equal(out, (x - mean) / (std + 1e-8)), k
        assert np.array_equal(out[0], ref[k]), k                           # env 0 = the reference env's own output
    assert env.obs_count == int(g["count"])
    assert np.array_equal(env.obs_mean[0].cpu().numpy(), g["mean"])
    assert np.array_equal(env.obs_std[0].cpu().numpy(), g["std"])
    env.close()


def test_vec_legacy_env_matches_reference_streams():
    """The reference's actual simulation-mode step for many envs in one launch (mlb_legacy_step, one warp per env):
    every env reproduces its own RandomState(seed) stream -- against the reference fixtures (7 steps at 16 / 64 / 5
    servers: 9 MT19937 twists at 64 servers, two server batches) and against the one-thread-per-env kernel over 300
    steps including the final generator state."""
    import ctypes as C
    import torch
    from marllb_b200 import VecLegacyEnv, _lib
    g = load_golden("legacy_env")
    for S, seed, metric in ((16, 7, "jain"), (64, 99, "variance"), (5, 3, "gini")):
        seeds = np.array([seed, seed + 1, seed, 12345], np.uint32)
        env = VecLegacyEnv(4, num_servers=S, seeds=seeds, reward_metric=metric, max_steps=6)
        o = env.reset().cpu().numpy()
        assert np.array_equal(o[0], g[f"s{S}_obs"][0]) and np.array_equal(o[2], o[0]) and not np.array_equal(o[1], o[0])
        for k in range(6):
            o, r, d = env.step()
            o, r = o.cpu().numpy(), r.cpu().numpy()
            assert np.array_equal(o[0], g[f"s{S}_obs"][k + 1]) and np.array_equal(o[2], o[0])
            assert r[0] == pytest.approx(g[f"s{S}_rew"][k], rel=1e-12) and r[2] == r[0]
            assert bool(d[0]) == bool(g[f"s{S}_done"][k])
        env.check_status()
    # long run, many seeds, odd server count: warp kernel == thread kernel, word for word
    E, S, T = 37, 40, 300
    seeds = (np.arange(E) * 7919 + 5).astype(np.uint32)
    env = VecLegacyEnv(E, num_servers=S, seeds=seeds)
    L = _lib.load()
    st = torch.empty((E, 625), dtype=torch.int32, device="cuda")
    ref_obs = torch.empty((E, S, 11), dtype=torch.float32, device="cuda")
    sd = torch.as_tensor(seeds.view(np.int32)).cuda()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.mlb_legacy_seed(C.c_void_p(st.data_ptr()), C.c_void_p(sd.data_ptr()), E, stream))
    env.reset()
    _lib.check(L.mlb_legacy_obs(C.c_void_p(st.data_ptr()), E, S, C.c_void_p(ref_obs.data_ptr()), stream))
    assert torch.equal(env.obs, ref_obs)
    for k in range(T):
        o, r, d = env.step()
        _lib.check(L.mlb_legacy_obs(C.c_void_p(st.data_ptr()), E, S, C.c_void_p(ref_obs.data_ptr()), stream))
        assert torch.equal(o, ref_obs), k
    # same generator position; the thread kernel twists lazily (at the next draw), the warp kernel like numpy too:
    # both leave position 624 untwisted, so the state words agree as well
    assert torch.equal(env._mt, st)
    # jain of column 10 over all servers, float64 (rewards.py:21-67)
    x = o[:, :, 10].double()
    want = (x.sum(1) ** 2) / (S * (x * x).sum(1))
    np.testing.assert_allclose(r.cpu().numpy(), want.cpu().numpy(), rtol=1e-12)
    env.check_status()
