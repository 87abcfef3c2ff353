"""Generate tests/golden/policy_*.npz by running the UNMODIFIED reference agents (torch CPU).

Run in the build container only:  python tests/golden/make_policy_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402


def sd_np(prefix, sd):
    return {f"{prefix}{k}": v.detach().cpu().numpy().copy() for k, v in sd.items()}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}.npz ({os.path.getsize(path) / 1024:.1f} KiB)")


def qmix():
    an = ref_import.load("agent_network", "problem-05-qmix")
    mn = ref_import.load("mixing_network", "problem-05-qmix")
    qa = ref_import.load("qmix_agent", "problem-05-qmix")
    torch.manual_seed(0)
    np.random.seed(0)
    out = {}
    # --- AgentQNetwork.forward, two chained steps
    net = an.AgentQNetwork(obs_dim=20, action_dim=6, hidden_dim=128, gru_dim=64)
    obs = torch.randn(5, 20)
    q1, h1 = net(obs, net.init_hidden(5))
    q2, h2 = net(obs * 0.5, h1)
    out.update(sd_np("net.", net.state_dict()))
    out.update(net_obs=obs.numpy(), net_q1=q1.detach().numpy(), net_h1=h1.detach().numpy(),
               net_q2=q2.detach().numpy(), net_h2=h2.detach().numpy())
    # --- QMixingNetwork.forward
    mix = mn.QMixingNetwork(num_agents=3, state_dim=9, mixing_embed_dim=32, hypernet_embed_dim=64)
    aq, st = torch.randn(7, 3), torch.randn(7, 9)
    out.update(sd_np("mix.", mix.state_dict()))
    out.update(mix_q=aq.numpy(), mix_state=st.numpy(), mix_out=mix(aq, st).detach().numpy())
    # --- select_actions (greedy + epsilon-greedy with numpy's global RNG)
    agent = qa.QMIXAgent(num_agents=3, state_dim=9, obs_dim=12, action_dim=5, hidden_dim=32, gru_dim=16,
                         mixing_embed_dim=8, hypernet_embed_dim=16, batch_size=4, max_seq_len=6, device="cpu",
                         target_update_interval=2)
    for i, n in enumerate(agent.agent_networks):
        out.update(sd_np(f"ag{i}.", n.state_dict()))
    out.update(sd_np("agmix.", agent.mixer.state_dict()))
    obs_list = [np.random.randn(12).astype(np.float32) for _ in range(3)]
    np.random.seed(123)
    acts, hids, qv = agent.select_actions(obs_list, None, evaluate=False, epsilon=0.5)
    acts_g, _, qv_g = agent.select_actions(obs_list, hids, evaluate=True)
    out.update(sel_obs=np.stack(obs_list), sel_acts=np.array(acts), sel_q=np.array(qv, np.float64),
               sel_acts_greedy=np.array(acts_g), sel_q_greedy=np.array(qv_g, np.float64),
               sel_h=np.stack([h.detach().numpy() for h in hids]))
    # --- update(): fixed batch, ragged lengths; two updates (second one triggers the hard target sync)
    rng = np.random.RandomState(5)
    B, T, A, K = 4, 6, 3, 5
    batch = {'observations': rng.randn(B, T, A, 12), 'actions': rng.randint(0, K, (B, T, A, 1)).astype(np.float64),
             'rewards': rng.rand(B, T, A), 'states': rng.randn(B, T, 9),
             'dones': np.zeros((B, T)), 'seq_lengths': np.array([6, 4, 6, 3], np.int32)}
    for b, L in enumerate(batch['seq_lengths']):
        batch['observations'][b, L:] = 0; batch['actions'][b, L:] = 0; batch['rewards'][b, L:] = 0
        batch['states'][b, L:] = 0
        batch['dones'][b, L - 1] = 1.0
    agent.episode_buffer.is_ready = lambda n: True
    agent.episode_buffer.sample_batch = lambda n, m: batch
    for k, v in batch.items():
        out["batch_" + k] = v
    for u in (1, 2):
        stats = agent.update()
        out[f"upd{u}_stats"] = np.array([stats['loss'], stats['q_tot'], stats['target_q_tot']])
        for i, n in enumerate(agent.agent_networks):
            out.update(sd_np(f"upd{u}.ag{i}.", n.state_dict()))
        out.update(sd_np(f"upd{u}.agmix.", agent.mixer.state_dict()))
        out.update(sd_np(f"upd{u}.tgt0.", agent.agent_networks_target[0].state_dict()))
    save("policy_qmix", **out)


def sac():
    nw = ref_import.load("networks", "problem-04-sac-gru")
    sa = ref_import.load("sac_agent", "problem-04-sac-gru")
    torch.manual_seed(1)
    np.random.seed(1)
    out = {}
    S, A, B = 22, 4, 8
    agent = sa.SAC_GRU_Agent(state_dim=S, action_dim=A, hidden_dim=64, gru_dim=32, batch_size=B, device="cpu")
    out.update(sd_np("policy.", agent.policy.state_dict()))
    out.update(sd_np("q1.", agent.q1.state_dict()))
    out.update(sd_np("q2.", agent.q2.state_dict()))
    # --- PolicyNetwork.sample / QNetwork.forward with known noise
    states = torch.randn(B, S)
    hidden = torch.randn(1, B, 32) * 0.3
    torch.manual_seed(77)
    eps = torch.randn(B, A)
    torch.manual_seed(77)
    with torch.no_grad():
        action, logp, mean_a, hn = agent.policy.sample(states, hidden)
        mean, log_std, _ = agent.policy.forward(states, hidden)
        assert torch.allclose(action, torch.tanh(mean + log_std.exp() * eps), atol=1e-6), "noise replay mismatch"
        q, qh = agent.q1.forward(states, action, hidden)
    out.update(fw_states=states.numpy(), fw_hidden=hidden.numpy(), fw_eps=eps.numpy(), fw_action=action.numpy(),
               fw_logp=logp.numpy(), fw_mean_action=mean_a.numpy(), fw_hn=hn.numpy(), fw_mean=mean.numpy(),
               fw_log_std=log_std.numpy(), fw_q=q.numpy(), fw_qh=qh.numpy())
    # --- update_parameters on a fixed batch, two updates
    rng = np.random.RandomState(9)
    fixed = (torch.as_tensor(rng.randn(B, S), dtype=torch.float32),
             torch.as_tensor(np.tanh(rng.randn(B, A)), dtype=torch.float32),
             torch.as_tensor(rng.rand(B, 1), dtype=torch.float32),
             torch.as_tensor(rng.randn(B, S), dtype=torch.float32),
             torch.as_tensor((rng.rand(B, 1) < 0.2).astype(np.float32)),
             torch.as_tensor(rng.randn(1, B, 32) * 0.2, dtype=torch.float32))
    for n, t in zip(("states", "actions", "rewards", "next_states", "dones", "hiddens"), fixed):
        out["batch_" + n] = t.numpy()
    agent.replay_buffer.is_ready = lambda n: True
    agent.replay_buffer.sample = lambda n, d: fixed
    for u in (1, 2):
        torch.manual_seed(100 + u)
        e_next, e_new = torch.randn(B, A), torch.randn(B, A)      # rsample order: next_states first, then states
        torch.manual_seed(100 + u)
        losses = agent.update_parameters(1)
        out[f"upd{u}_eps_next"], out[f"upd{u}_eps_new"] = e_next.numpy(), e_new.numpy()
        out[f"upd{u}_losses"] = np.array([losses['q1'], losses['q2'], losses['policy'], losses['alpha']])
        out[f"upd{u}_alpha"] = np.array([agent.alpha.item()])
        out.update(sd_np(f"upd{u}.policy.", agent.policy.state_dict()))
        out.update(sd_np(f"upd{u}.q1.", agent.q1.state_dict()))
        out.update(sd_np(f"upd{u}.q2.", agent.q2.state_dict()))
        out.update(sd_np(f"upd{u}.q1t.", agent.q1_target.state_dict()))
    save("policy_sac", **out)


def extras():
    """WeightedQMixingNetwork, QMixingNetwork.get_monotonicity_info (mixing_network.py:119-151,187-246) and
    PrioritizedReplayBuffer (replay_buffer.py:105-221)."""
    mn = ref_import.load("mixing_network", "problem-05-qmix")
    rb = ref_import.load("replay_buffer", "problem-04-sac-gru")
    torch.manual_seed(2)
    out = {}
    wq = mn.WeightedQMixingNetwork(num_agents=4, state_dim=11, hidden_dim=16)
    aq = torch.randn(9, 4, requires_grad=True)
    st = torch.randn(9, 11)
    q_tot, w = wq(aq, st)
    coef = torch.randn(9, 1)
    (q_tot * coef).sum().backward()
    out.update(sd_np("wq.", wq.state_dict()))
    out.update(wq_q=aq.detach().numpy(), wq_state=st.numpy(), wq_out=q_tot.detach().numpy(), wq_w=w.detach().numpy(),
               wq_coef=coef.numpy(), wq_dq=aq.grad.numpy())
    out.update({"wqg." + k: v.grad.numpy().copy() for k, v in wq.named_parameters()})
    mix = mn.QMixingNetwork(num_agents=3, state_dim=9, mixing_embed_dim=32, hypernet_embed_dim=64)
    aq2, st2 = torch.randn(7, 3, requires_grad=True), torch.randn(7, 9)
    out.update(sd_np("mono.", mix.state_dict()))
    out.update(mono_q=aq2.detach().numpy(), mono_state=st2.numpy(),
               mono_grad=mix.get_monotonicity_info(aq2, st2).detach().numpy())
    # prioritised replay: ring, max-priority insertion, numpy sampling, importance weights, beta schedule
    buf = rb.PrioritizedReplayBuffer(capacity=16, alpha=0.6, beta=0.4, beta_increment=0.01, seed=3)
    rng = np.random.RandomState(0)
    for i in range(21):
        buf.push(rng.randn(5).astype(np.float32), rng.randn(2).astype(np.float32), float(i), rng.randn(5).astype(np.float32),
                 float(i % 7 == 0), rng.randn(1, 1, 4).astype(np.float32))
        if i == 9:
            buf.update_priorities([0, 3, 5], [2.5, 0.1, 7.0])
    np.random.seed(11)
    res = buf.sample(6)
    out.update(per_rewards=res[2].numpy(), per_states=res[0].numpy(), per_hiddens=res[5].numpy(), per_weights=res[6].numpy(),
               per_indices=np.asarray(res[7]), per_beta=np.float64(buf.beta), per_priorities=buf.priorities.copy(),
               per_size=np.int64(len(buf)), per_position=np.int64(buf.position))
    save("policy_extras", **out)


def checkpoints():
    """Checkpoint FILES written by the unmodified reference agents after two updates (torch.optim.Adam state
    included: sac_agent.py:257-287, qmix_agent.py:309-322) + every parameter after a THIRD update run by the
    reference after its own load() of that file.  The product must resume from these files to the same numbers."""
    sa = ref_import.load("sac_agent", "problem-04-sac-gru")
    qa = ref_import.load("qmix_agent", "problem-05-qmix")
    out = {}
    # ---- SAC
    torch.manual_seed(3)
    S, A, B = 22, 4, 8
    kw = dict(state_dim=S, action_dim=A, hidden_dim=64, gru_dim=32, batch_size=B, device="cpu")
    agent = sa.SAC_GRU_Agent(**kw)
    rng = np.random.RandomState(19)
    fixed = (torch.as_tensor(rng.randn(B, S), dtype=torch.float32),
             torch.as_tensor(np.tanh(rng.randn(B, A)), dtype=torch.float32),
             torch.as_tensor(rng.rand(B, 1), dtype=torch.float32),
             torch.as_tensor(rng.randn(B, S), dtype=torch.float32),
             torch.as_tensor((rng.rand(B, 1) < 0.2).astype(np.float32)),
             torch.as_tensor(rng.randn(1, B, 32) * 0.2, dtype=torch.float32))
    for n, t in zip(("states", "actions", "rewards", "next_states", "dones", "hiddens"), fixed):
        out["sac_batch_" + n] = t.numpy()
    agent.replay_buffer.is_ready = lambda n: True
    agent.replay_buffer.sample = lambda n, d: fixed
    for u in (1, 2):
        torch.manual_seed(200 + u)
        agent.update_parameters(1)
    agent.save(os.path.join(HERE, "ref_sac_ckpt.pt"))
    resumed = sa.SAC_GRU_Agent(**kw)
    resumed.load(os.path.join(HERE, "ref_sac_ckpt.pt"))
    resumed.replay_buffer.is_ready = lambda n: True
    resumed.replay_buffer.sample = lambda n, d: fixed
    torch.manual_seed(203)
    e_next, e_new = torch.randn(B, A), torch.randn(B, A)
    torch.manual_seed(203)
    losses = resumed.update_parameters(1)
    out["sac_upd3_eps_next"], out["sac_upd3_eps_new"] = e_next.numpy(), e_new.numpy()
    out["sac_upd3_losses"] = np.array([losses['q1'], losses['q2'], losses['policy'], losses['alpha']])
    out["sac_upd3_alpha"] = np.array([resumed.alpha.item()])
    # Reference quirk (sac_agent.py:304-306): load() REBINDS self.log_alpha to the checkpoint tensor, so the
    # temperature optimiser is left holding the orphaned initial parameter (which never receives a gradient again) and
    # alpha stays frozen at its checkpoint value.  What a resume SHOULD give is what the never-reloaded agent gives on
    # the same third update: recorded as the expected temperature.
    out["sac_ckpt_log_alpha"] = np.array([float(torch.load(os.path.join(HERE, "ref_sac_ckpt.pt"))["log_alpha"].item())])
    torch.manual_seed(203)
    agent.update_parameters(1)
    out["sac_upd3_log_alpha_cont"] = np.array([float(agent.log_alpha.item())])
    for (k1, v1), (k2, v2) in zip(agent.policy.state_dict().items(), resumed.policy.state_dict().items()):
        assert torch.equal(v1, v2), k1           # everything else resumes identically
    for pre, net in (("policy.", resumed.policy), ("q1.", resumed.q1), ("q2.", resumed.q2), ("q1t.", resumed.q1_target)):
        out.update(sd_np("sac_upd3." + pre, net.state_dict()))
    # ---- QMIX
    torch.manual_seed(4)
    qkw = dict(num_agents=3, state_dim=9, obs_dim=12, action_dim=5, hidden_dim=32, gru_dim=16, mixing_embed_dim=8,
               hypernet_embed_dim=16, batch_size=4, max_seq_len=6, device="cpu", target_update_interval=2)
    qagent = qa.QMIXAgent(**qkw)
    rng = np.random.RandomState(6)
    Bq, T, Aq, K = 4, 6, 3, 5
    batch = {'observations': rng.randn(Bq, T, Aq, 12), 'actions': rng.randint(0, K, (Bq, T, Aq, 1)).astype(np.float64),
             'rewards': rng.rand(Bq, T, Aq), 'states': rng.randn(Bq, T, 9),
             'dones': np.zeros((Bq, T)), 'seq_lengths': np.array([5, 6, 3, 6], np.int32)}
    for b, L in enumerate(batch['seq_lengths']):
        batch['observations'][b, L:] = 0; batch['actions'][b, L:] = 0; batch['rewards'][b, L:] = 0
        batch['states'][b, L:] = 0
        batch['dones'][b, L - 1] = 1.0
    for k, v in batch.items():
        out["qmix_batch_" + k] = v
    qagent.episode_buffer.is_ready = lambda n: True
    qagent.episode_buffer.sample_batch = lambda n, m: batch
    qagent.update(); qagent.update()
    qagent.save(os.path.join(HERE, "ref_qmix_ckpt.pt"))
    qres = qa.QMIXAgent(**qkw)
    qres.load(os.path.join(HERE, "ref_qmix_ckpt.pt"))
    qres.episode_buffer.is_ready = lambda n: True
    qres.episode_buffer.sample_batch = lambda n, m: batch
    stats = qres.update()
    out["qmix_upd3_stats"] = np.array([stats['loss'], stats['q_tot'], stats['target_q_tot']])
    for i, n in enumerate(qres.agent_networks):
        out.update(sd_np(f"qmix_upd3.ag{i}.", n.state_dict()))
    out.update(sd_np("qmix_upd3.agmix.", qres.mixer.state_dict()))
    save("policy_ckpt", **out)


if __name__ == "__main__":
    if "ckpt" in sys.argv:
        checkpoints()
        sys.exit(0)
    if "extras" not in sys.argv:
        qmix()
        sac()
    extras()
    checkpoints()
