"""GPU: the fused rollout (env step + batched QMIX action selection) against (a) a float64 torch
evaluation of the same agent networks on the same observations and (b) a second env stepped with the
same per-server actions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(E, A, Sa):
    from marllb_b200 import VecLoadBalanceEnv
    env = VecLoadBalanceEnv(E, num_servers=Sa, num_agents=A, reservoir_capacity=128, max_steps=10 ** 6,
                            action_dtype="uint8")
    env.set_speeds(np.where(np.arange(Sa * A) % 2 == 0, 1.0, 2.0).astype(np.float32))
    env.gen_poisson(64.0, 0.8 * 1.5 * Sa / 64.0, 8.0, seed=11)
    env.reset()
    return env


def _q_ref(net, obs, h):
    """AgentQNetwork.forward in float64 torch (agent_network.py:63-87; nn.GRU gate order r, z, n)."""
    P = {k: v.double() for k, v in net.state_dict().items()}
    x, h = obs.double(), h.double()
    gi = x @ P["gru.weight_ih_l0"].T + P["gru.bias_ih_l0"]
    gh = h @ P["gru.weight_hh_l0"].T + P["gru.bias_hh_l0"]
    H = h.shape[1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    hn = (1 - z) * n + z * h
    a = torch.relu(hn @ P["fc1.weight"].T + P["fc1.bias"])
    a = torch.relu(a @ P["fc2.weight"].T + P["fc2.bias"])
    return a @ P["fc3.weight"].T + P["fc3.bias"], hn


@pytest.mark.parametrize("E", [96, 640])   # FFMA path and tensor-core path (M >= 512)
def test_qmix_rollout_matches_float64_policy_and_plain_env(E):
    from marllb_b200.policy import QMIXAgent
    from marllb_b200.rollout import QMIXRollout
    A, Sa = 2, 8
    torch.manual_seed(3)
    agent = QMIXAgent(num_agents=A, state_dim=4 * A * Sa + 10, obs_dim=Sa * 11, action_dim=Sa, hidden_dim=128, gru_dim=64)
    env, env2 = _env(E, A, Sa), _env(E, A, Sa)
    ro = QMIXRollout(env, agent)
    h_ref = [torch.zeros(E, 64, dtype=torch.float64, device="cuda") for _ in range(A)]
    g = torch.Generator(device="cuda").manual_seed(1)
    for k in range(12):
        obs_before = env.obs.clone().view(E, A, Sa * 11)
        u = torch.rand(E, A, device="cuda", generator=g)
        rnd = torch.randint(0, Sa, (E, A), device="cuda", generator=g, dtype=torch.int32)
        o, r, d, act = ro.step(epsilon=0.2, u=u, rnd=rnd)
        # (a) policy: greedy choices agree with float64 wherever the top-2 margin is not a rounding tie
        for a in range(A):
            q, h_ref[a] = _q_ref(agent.agent_networks[a], obs_before[:, a], h_ref[a])
            top2 = q.topk(2, dim=1).values
            clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * q.abs().max()
            greedy = u[:, a] >= 0.2
            want = torch.where(greedy, q.argmax(1).int(), rnd[:, a])
            ok = (act[:, a] == want) | (greedy & ~clear)
            assert bool(ok.all()), (k, a, int((~ok).sum()))
            np.testing.assert_allclose(ro.hidden[a].cpu().numpy(), h_ref[a].cpu().numpy(), rtol=1e-4, atol=2e-5)
            h_ref[a] = ro.hidden[a].double()       # keep following the float32 trajectory
        # (b) env: the same per-server actions on a second env give the same transition
        env_action = torch.zeros(E, A * Sa, dtype=torch.uint8, device="cuda")
        env_action.scatter_(1, (act.long() + torch.arange(A, device="cuda") * Sa), 2)
        o2, r2, d2 = env2.step(env_action)
        assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d, d2)
    env.check_status()


def test_sac_rollout_device_replay_and_update():
    from marllb_b200 import VecLoadBalanceEnv
    from marllb_b200.policy import SAC_GRU_Agent
    from marllb_b200.rollout import SACRollout
    E, S = 48, 16

    def mk():
        env = VecLoadBalanceEnv(E, num_servers=S, action_type="continuous", max_steps=10 ** 6)
        env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0).astype(np.float32))
        env.gen_poisson(64.0, 0.8 * 1.5 * S / 64.0, 10.0, seed=5)
        env.reset()
        return env
    env, env2 = mk(), mk()
    torch.manual_seed(11)
    agent = SAC_GRU_Agent(state_dim=S * 11, action_dim=S, hidden_dim=64, gru_dim=32, batch_size=64)
    ro = SACRollout(env, agent, replay_capacity=100)      # wraps around: 48 per step into 100 slots
    g = torch.Generator(device="cuda").manual_seed(2)
    for k in range(6):
        state_before = env.obs.clone().view(E, S * 11)
        h_before = ro.hidden[0].clone()
        eps = torch.randn(E, S, device="cuda", generator=g)
        o, r, d, a = ro.step(eps=eps)
        assert bool((a.abs() <= 1).all())                                        # tanh-squashed (networks.py:136)
        o2, r2, d2 = env2.step(a)
        assert torch.equal(o, o2) and torch.equal(r, r2)
        # the newest transitions sit right behind the write position
        pos = (ro.replay.pos - E) % 100
        idx = (torch.arange(E, device="cuda") + pos) % 100
        assert torch.equal(ro.replay.state[idx], state_before) and torch.equal(ro.replay.action[idx], a)
        assert torch.equal(ro.replay.next_state[idx], o.view(E, S * 11)) and torch.equal(ro.replay.hidden[idx], h_before)
        assert torch.equal(ro.replay.reward[idx, 0], r.float())
    assert len(ro.replay) == 100
    p0 = agent.policy.state_dict()["fc1.weight"].clone()
    out = ro.update(2, g)
    assert set(out) == {"q1", "q2", "policy", "alpha"} and all(np.isfinite(list(out.values())))
    assert not torch.equal(p0, agent.policy.state_dict()["fc1.weight"])


@pytest.mark.parametrize("overlap", [False, True])
def test_sac_rollout_graph_sequential_and_pipelined(overlap):
    """SACRollout.capture(): actor + env step + replay push + SAC update replayed from one CUDA graph.  overlap=True puts
    the update beside the env step (its batch is gathered before this step's push).  Either way: the ring keeps
    receiving the newest transitions, the learner moves, losses stay finite, the env stays healthy, and two identical
    runs end bit-identical (the branches of the graph do not race)."""
    from marllb_b200 import VecLoadBalanceEnv
    from marllb_b200.policy import SAC_GRU_Agent
    from marllb_b200.rollout import SACRollout
    E, S, CAP = 64, 16, 256

    def run():
        env = VecLoadBalanceEnv(E, num_servers=S, action_type="continuous", max_steps=10 ** 6)
        env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0).astype(np.float32))
        env.gen_poisson(64.0, 0.8 * 1.5 * S / 64.0, 20.0, seed=5)
        env.reset()
        torch.manual_seed(3)
        agent = SAC_GRU_Agent(state_dim=S * 11, action_dim=S, hidden_dim=64, gru_dim=32, batch_size=64)
        ro = SACRollout(env, agent, replay_capacity=CAP)
        for _ in range(CAP // E):                       # fill the ring eagerly (capture needs a full ring)
            ro.step()
        ro.capture(1, overlap=overlap)
        p0 = agent.policy.state_dict()["fc1.weight"].clone()
        for k in range(12):
            state_before = env.obs.clone().view(E, S * 11)
            (o, r, d, a), losses = ro.step_graph()
            assert bool((a.abs() <= 1).all()) and all(bool(torch.isfinite(v)) for v in losses.values())
            # the transitions of this step are in the ring (the device-held write position advanced by E)
            assert bool((ro.replay.state.view(CAP // E, E, -1) == state_before).all(-1).all(-1).any()), k
            assert bool((ro.replay.next_state.view(CAP // E, E, -1) == o.view(E, S * 11)).all(-1).all(-1).any()), k
        torch.cuda.synchronize()
        env.check_status()
        assert int(env.get_state("step")[0]) == CAP // E + 2 + 12          # fill + capture warm-up + replays
        assert not torch.equal(p0, agent.policy.state_dict()["fc1.weight"])
        out = (agent.policy.state_dict()["fc1.weight"].clone(), agent.q1.state_dict()["fc2.weight"].clone(), o.clone())
        del ro._graph
        env.close()
        return out
    a_, b_ = run(), run()
    for x, y in zip(a_, b_):
        assert torch.equal(x, y)
