"""GPU: the original-paper agents (src/lb/sac_qmix.py RNNAgent / QMix / QMix_Trainer with
TD(lambda) targets, src/lb/sac_gru_discrete.py discrete SAC) against fixtures produced by the reference's own
classes on torch CPU (tests/golden/make_paper_golden.py).  Tolerance: 1e-5 relative for float32 network
outputs, 1e-4 for parameters after two optimiser steps (Adam divides by sqrt(v) ~ |g|, which amplifies
rounding of tiny gradients), with absolute floors for values near zero."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-6


def sd(g, prefix):
    return {k[len(prefix):]: torch.as_tensor(v) for k, v in g.items() if k.startswith(prefix)}


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def check_params(net, g, prefix, rtol=1e-4, atol=2e-5):
    ref, mine = sd(g, prefix), net.state_dict()
    assert set(ref) == set(mine), (sorted(ref), sorted(mine))
    for k in ref:
        close(mine[k], ref[k].numpy(), rtol, atol)


def test_new_kernels_against_torch():
    from marllb_b200.policy import ops
    torch.manual_seed(0)
    x = torch.randn(7, 5, 3, 6, device="cuda")
    y = ops.softmax_forward(x)
    close(y, torch.softmax(x.double(), -1).cpu().numpy(), 1e-6, 1e-7)
    dy = torch.randn_like(x)
    xr = x.double().requires_grad_(True)
    (torch.softmax(xr, -1) * dy.double()).sum().backward()
    close(ops.softmax_backward(y, dy), xr.grad.cpu().numpy(), 1e-5, 1e-6)
    st = torch.randn(11, 4, device="cuda")
    a = torch.randint(0, 3, (11, 2), device="cuda", dtype=torch.int32)
    want = torch.cat([st, torch.nn.functional.one_hot(a.long(), 3).float().reshape(11, 6)], -1)
    assert torch.equal(ops.concat_onehot(st, a, 3), want)
    p = torch.softmax(torch.randn(50, 4, device="cuda"), -1)
    u = torch.rand(50, device="cuda")
    act, logp, ps = ops.categorical(p, u=u, want_p=True)
    cdf = p.cumsum(-1)
    want_a = (u[:, None] * cdf[:, -1:] >= cdf).sum(-1).clamp_max(3)
    assert torch.equal(act.long(), want_a)
    close(logp, torch.log(p.gather(1, want_a[:, None])[:, 0]).cpu().numpy(), 1e-6, 1e-7)
    assert torch.equal(ops.categorical(p, want_logp=False)[0].long(), p.argmax(-1))
    r, q = torch.rand(6, 9, device="cuda"), torch.randn(6, 9, device="cuda")
    ret = torch.zeros_like(q)
    ret[:, -1] = q[:, -1]
    for t in range(7, -1, -1):
        ret[:, t] = 0.6 * 0.99 * ret[:, t + 1] + (r[:, t] + (1 - 0.6) * 0.99 * q[:, t + 1])
    close(ops.td_lambda_targets(r, q), ret.cpu().numpy(), 1e-6, 1e-7)
    close(ops.reward_normalize(r, 10.0), (10.0 * (r - r.mean(0)) / (r.std(0) + 1e-6)).cpu().numpy(), 1e-5, 1e-5)


def test_rnn_agent_and_mixer_forward_match_reference():
    from marllb_b200.policy.paper import QMix_Trainer
    g = load_golden("paper_qmix")
    H, A, Fd, heads, n, hyp, B, T = (int(x) for x in g["dims"])
    tr = QMix_Trainer(None, A, Fd, heads, n, H, hyp, lr=0.001)
    assert set(tr.agent.state_dict()) == set(sd(g, "agent0.")) and set(tr.mixer.state_dict()) == set(sd(g, "mixer0."))
    tr.agent.load_state_dict(sd(g, "agent0."))
    tr.mixer.load_state_dict(sd(g, "mixer0."))
    qs, hid = tr.agent(g["state"], g["last_action"], g["hidden_in"])
    assert qs.shape == (B, T, A, heads, n) and hid.shape == (1, B * A, H)
    close(qs, g["fwd_qs"])
    close(hid, g["fwd_hidden"])
    chosen = qs.gather(-1, torch.as_tensor(g["action"]).cuda().unsqueeze(-1)).squeeze(-1)
    qtot = tr.mixer(chosen, g["state"])
    close(qtot, g["fwd_qtot"], 1e-5, 1e-5)
    close(tr._build_td_lambda_targets(g["reward"][..., None], g["fwd_qtot"]), g["fwd_td_lambda"], 1e-6, 1e-6)
    a, h = tr.agent.get_action(g["state"][0, :, 0, :], g["last_action"][0, :, 0, :], np.zeros((1, 1, H), np.float32),
                               deterministic=True)
    assert np.array_equal(a, g["get_action_det"])
    close(h, g["get_action_hidden"])
    # stochastic path: inverse-CDF draws stay inside the support and follow the probabilities' argmax for u -> 0
    a0, _ = tr.agent.get_action(g["state"][0, :, 0, :], g["last_action"][0, :, 0, :], np.zeros((1, 1, H), np.float32),
                                u=np.zeros((T, heads), np.float32))
    assert a0.shape == (T, heads) and (a0 == 0).all()


def test_qmix_trainer_two_td_lambda_updates_match_reference():
    from marllb_b200.policy.paper import QMix_Trainer
    g = load_golden("paper_qmix")
    H, A, Fd, heads, n, hyp, B, T = (int(x) for x in g["dims"])
    tr = QMix_Trainer(None, A, Fd, heads, n, H, hyp, lr=0.001)
    tr.agent.load_state_dict(sd(g, "agent0."))
    tr.mixer.load_state_dict(sd(g, "mixer0."))
    tr._update_targets()
    batch = (g["hidden_in"], g["state"], g["action"], g["last_action"], g["reward"], g["next_state"])
    for k in range(2):
        out = tr.update(B, batch=batch)
        assert out["loss"] == pytest.approx(float(g["losses"][k]), rel=1e-4)
        check_params(tr.agent, g, f"agent{k + 1}.")
        check_params(tr.mixer, g, f"mixer{k + 1}.")


def test_replay_buffer_gru_ring_and_centre_crop():
    from marllb_b200.policy.paper import ReplayBufferGRU
    g = load_golden("paper_qmix")
    H, A, Fd, heads = (int(x) for x in g["dims"][:4])
    buf = ReplayBufferGRU(3, "/nonexistent/replay.pkl")
    for i, L in enumerate(g["buf_lens"]):
        L = int(L)
        buf.push(torch.full((1, 1, A, H), float(i)), np.full((L, A, Fd), i, np.float32) + np.arange(L)[:, None, None],
                 np.full((L, A, heads), i), np.full((L, A, heads), i), np.arange(L, dtype=np.float32) + 10 * i,
                 np.full((L, A, Fd), -i, np.float32))
    assert len(buf) == 3 and buf.get_length() == 3 and buf.position == int(g["buf_position"])
    random.seed(5)
    hi, s, a, la, r, ns = buf.sample(2)
    assert np.array_equal(hi.numpy(), g["buf_hidden"]) and np.array_equal(np.asarray(s), g["buf_state"])
    assert np.array_equal(np.asarray(r), g["buf_reward"])


def test_discrete_sac_forward_matches_reference():
    from marllb_b200.policy.paper import SAC_Trainer
    g = load_golden("paper_sac")
    H, Fd, heads, n, B, T = (int(x) for x in g["dims"])
    tr = SAC_Trainer(None, Fd, n, H, heads)
    for tag, net in (("q1", tr.soft_q_net1), ("q2", tr.soft_q_net2), ("pi", tr.policy_net)):
        assert set(net.state_dict()) == set(sd(g, f"{tag}_0."))
        net.load_state_dict(sd(g, f"{tag}_0."))
    probs, hid = tr.policy_net(g["state"], g["last_action"], g["hidden_in"])
    assert probs.shape == (B, T, heads, n) and hid.shape == (1, B, H)
    close(probs, g["fwd_probs"])
    close(hid, g["fwd_hidden"])
    q, qh = tr.soft_q_net1(g["state"], g["action"], g["hidden_in"])
    close(q, g["fwd_q"])
    close(qh, g["fwd_q_hidden"])
    a, _ = tr.policy_net.get_action(g["state"][0], g["last_action"][0], g["hidden_in"][:, :1], deterministic=True)
    assert np.array_equal(a, g["get_action_det"])
    act, lp, _ = tr.policy_net.evaluate(g["state"], g["last_action"], g["hidden_in"], given=g["sampled"][0])
    assert np.array_equal(act.cpu().numpy(), g["sampled"][0])
    close(lp, g["logps"][0], 1e-5, 1e-5)


def test_discrete_sac_two_updates_match_reference():
    from marllb_b200.policy.paper import SAC_Trainer
    g = load_golden("paper_sac")
    H, Fd, heads, n, B, T = (int(x) for x in g["dims"])
    tr = SAC_Trainer(None, Fd, n, H, heads)
    for tag, net in (("q1", tr.soft_q_net1), ("q2", tr.soft_q_net2), ("pi", tr.policy_net),
                     ("q1", tr.target_soft_q_net1), ("q2", tr.target_soft_q_net2)):
        net.load_state_dict(sd(g, f"{tag}_0."))
    batch = (g["hidden_in"], g["hidden_out"], g["state"], g["action"], g["last_action"], g["reward"], g["next_state"])
    for k in range(2):
        ret = tr.update(B, batch=batch, sampled=(g["sampled"][2 * k], g["sampled"][2 * k + 1]))
        assert ret == pytest.approx(float(g["returns"][k]), rel=1e-4, abs=1e-5)
        for tag, net in (("q1", tr.soft_q_net1), ("q2", tr.soft_q_net2), ("pi", tr.policy_net),
                         ("t1", tr.target_soft_q_net1), ("t2", tr.target_soft_q_net2)):
            check_params(net, g, f"{tag}_{k + 1}.")
        close(tr.log_alpha, g[f"log_alpha_{k + 1}"], 1e-5, 1e-7)
