"""GPU: the host-buffer step and its changed-rows form (`mlb_step_changed`), and windowed Poisson generation.

step_host(obs="changed") must leave exactly the bytes in the persistent host observation array that the full copy
would: checked against a twin env stepped through step_host(obs="full") and, independently, against the device
observation, early in an episode (most rows change: block-copy fallback), deep into one (few rows change: record
path), with the chunked pipeline (E >= 8192) and without it, and across reset() / device-side steps (key frames).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(E, S, K, steps, seed=3, **kw):
    from marllb_b200 import VecLoadBalanceEnv
    env = VecLoadBalanceEnv(E, num_servers=S, reservoir_capacity=K, max_steps=10 ** 9, action_dtype="uint8", **kw)
    env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0).astype(np.float32))
    env.gen_poisson(2.0 * S, 0.8 * 1.5 * S / (2.0 * S), (steps + 2) * 0.25, seed=seed)
    env.reset()
    return env


@pytest.mark.parametrize("E,S,K,steps", [
    (8192, 16, 8, 260),       # chunked pipeline (8 chunks of 1024 envs); K = 8: the replacement regime comes early
    (300, 40, 16, 200),       # single chunk, 2 servers per lane, ragged sizes
    (1024, 64, 128, 60),      # fill phase only: every touched row changes
])
def test_changed_rows_equal_full_copy(E, S, K, steps):
    import torch
    a = _make(E, S, K, steps)
    b = _make(E, S, K, steps)
    rng = np.random.RandomState(0)
    moved_changed, moved_full, sparse_steps = 0, 0, 0
    for k in range(steps):
        act = rng.randint(0, 3, (E, S)).astype(np.uint8)
        oa, ra, da = a.step_host(act, obs="changed")
        ob, rb, db = b.step_host(act, obs="full")
        assert oa.tobytes() == ob.tobytes(), k                              # bit-equal, every step
        assert np.array_equal(ra, rb) and np.array_equal(da, db)
        moved_changed += a.last_d2h_bytes
        moved_full += b.last_d2h_bytes
        sparse_steps += a.last_d2h_bytes < b.last_d2h_bytes
        if k % 50 == 7:
            assert np.array_equal(oa, a.obs.cpu().numpy())                  # and equal to the device observation
    assert moved_changed <= moved_full + 64 * 4 * steps                     # never more than the full copy (+ the counts)
    if K <= 16:
        assert sparse_steps > steps // 4, (sparse_steps, steps)             # the record path was really exercised
        assert moved_changed < 0.8 * moved_full
    # a device-side step or a reset invalidates the host mirror: the next call re-keys it with a full copy
    act = rng.randint(0, 3, (E, S)).astype(np.uint8)
    a.step(torch.as_tensor(act).cuda()); b.step(torch.as_tensor(act).cuda())
    act = rng.randint(0, 3, (E, S)).astype(np.uint8)
    oa, _, _ = a.step_host(act, obs="changed")
    ob, _, _ = b.step_host(act, obs="full")
    assert a.last_d2h_bytes == b.last_d2h_bytes and oa.tobytes() == ob.tobytes()
    a.reset(); b.reset()
    for k in range(3):
        act = rng.randint(0, 3, (E, S)).astype(np.uint8)
        oa, _, _ = a.step_host(act, obs="changed")
        ob, _, _ = b.step_host(act, obs="full")
        assert oa.tobytes() == ob.tobytes()
    a.check_status(); b.check_status()
    a.close(); b.close()


def test_changed_rows_multi_agent_and_no_feature_cache():
    """A = 2 agents per env; feature_cache=False recomputes every reservoir each step, so every row is 'changed'."""
    from marllb_b200 import VecLoadBalanceEnv
    E, A, Sa, steps = 2048, 2, 8, 40
    envs = []
    for fc in (True, True, False):
        e = VecLoadBalanceEnv(E, num_servers=Sa, num_agents=A, reservoir_capacity=8, max_steps=10 ** 9, feature_cache=fc)
        e.gen_poisson(16.0, 0.5, (steps + 2) * 0.25, seed=9)
        e.reset()
        envs.append(e)
    rng = np.random.RandomState(1)
    for k in range(steps):
        act = rng.randint(0, 3, (E, A * Sa)).astype(np.int32)
        o0, r0, _ = envs[0].step_host(act, obs="changed")
        o1, r1, _ = envs[1].step_host(act, obs="full")
        o2, r2, _ = envs[2].step_host(act, obs="changed")
        assert o0.tobytes() == o1.tobytes() and np.array_equal(r0, r1)
        np.testing.assert_allclose(o2, o1, rtol=1e-5, atol=1e-7)            # recomputed-from-scratch features: float path
    for e in envs:
        e.close()


def test_poisson_windows_continue_an_episode():
    """mlb_gen_poisson_window: arrivals of [t_start, t_end) replace the resident ones between two steps.  Window 0 from
    t = 0 is mlb_gen_poisson; later windows start where the stepped time ends, keep the rate, and the episode goes on
    (flow conservation: arrivals consumed = flows in system + completed + dropped)."""
    E, S, K = 64, 16, 16
    env = _make(E, S, K, 40, seed=5)
    ref = _make(E, S, K, 40, seed=5)
    a0 = env.get_arrivals(3)
    env.gen_poisson(32.0, 0.6, 42 * 0.25, seed=5, t_start=0.0, window=0)
    assert np.array_equal(a0["time"], env.get_arrivals(3)["time"])           # window 0 == the plain generator
    rng = np.random.RandomState(2)
    consumed = np.zeros((E, 1), np.int64)
    for w in range(4):
        t0 = w * 40 * 0.25
        if w > 0:
            consumed += env.get_state("arr_cursor")
            env.gen_poisson(32.0, 0.6, t0 + 40 * 0.25 + 0.5, seed=5, t_start=t0, window=w)
            arr = env.get_arrivals(7)
            assert arr["time"].min() >= np.float32(t0) and np.all(np.diff(arr["time"]) >= 0)
            rate = len(arr["time"]) / (40 * 0.25 + 0.5)
            assert 20.0 < rate < 46.0, rate
            assert not np.array_equal(arr["time"][:8] - np.float32(t0), a0["time"][:8])   # a different draw per window
        for k in range(40):
            env.step(rng.randint(0, 3, (E, S)).astype(np.uint8))
    env.check_status()
    consumed += env.get_state("arr_cursor")
    n_on = env.get_state("n_flow_on").sum(axis=1)
    fct = env.get_state("res_count")[:, 0, :].astype(np.int64).sum(axis=1)       # one fct sample per completed flow
    drp = env.get_state("dropped").astype(np.int64).sum(axis=1)
    assert np.array_equal(consumed[:, 0], n_on + fct + drp)
    assert int(env.get_state("step")[0]) == 160 and consumed.min() > 3 * 40 * 0.25 * 32 * 0.6
    env.close(); ref.close()
