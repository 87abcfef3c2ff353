"""GPU: the learners' update steps AT THE BENCHMARKED SIZES against the unmodified reference agents.

  config C4   SAC_GRU_Agent.update_parameters   state 2816, action 256, hidden 256, gru 128, batch 256  (sac_agent.py:151-255)
  config C3   QMIXAgent.update                  2 agents, obs 352, 32 actions, 32 episodes x 50 steps    (qmix_agent.py:192-307)

These are the shapes the tensor-core GEMMs (3xTF32, split-K, transposed operands), the skinny-GEMM kernels and the
fused GRU time loops actually run at; the toy-size fixtures of test_gpu_policy.py never reach them.  Fixture:
tests/golden/make_fullsize_golden.py ran the reference on torch CPU over the inputs of tests/fullsize_spec.py and kept
sampled entries.  Tolerances (north_star: 1e-5 relative on float32 outputs): losses 5e-5 relative; gradients (Adam's first
moment after one step = 0.1 x gradient) 1e-4 of the tensor's max-norm; parameters after each of three optimiser
steps 1e-4 relative with a 2e-6 absolute floor, the same bar as the toy-size tests.
"""
import numpy as np
import pytest
import torch

import fullsize_spec as spec
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _load_synth(net, seed):
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict({k: torch.as_tensor(v) for k, v in spec.synth_state_dict(shapes, seed).items()})


def _check(g, tag, named, salt, rtol, atol_of_absmax=0.0, atol=0.0):
    worst = 0.0
    for name, t in named:
        a = t.detach().cpu().numpy().reshape(-1)
        mine = a[spec.sample_index(name, a.size, salt)]
        ref = g[f"{tag}.{name}"]
        tol = atol + atol_of_absmax * float(g[f"{tag}.{name}.absmax"][0]) + rtol * np.abs(ref)
        err = np.abs(mine - ref)
        worst = max(worst, float((err / np.maximum(tol, 1e-30)).max()))
        assert (err <= tol).all(), (tag, name, float(err.max()), float(tol.min()))
    return worst


def test_sac_update_at_c4_size_matches_reference():
    from marllb_b200.policy import SAC_GRU_Agent
    g = load_golden("policy_fullsize")
    torch.manual_seed(0)
    agent = SAC_GRU_Agent(**spec.SAC)
    for k, net in enumerate((agent.policy, agent.q1, agent.q2)):
        _load_synth(net, 11 + k)
    agent.q1_target.load_state_dict(agent.q1.state_dict())
    agent.q2_target.load_state_dict(agent.q2.state_dict())
    for u in range(1, spec.N_UPDATES + 1):
        batch, eps_next, eps_new = spec.sac_batch(u)
        losses = agent.update_parameters(1, batch=tuple(torch.as_tensor(x) for x in batch),
                                         eps_next=torch.as_tensor(eps_next), eps_new=torch.as_tensor(eps_new))
        np.testing.assert_allclose([losses['q1'], losses['q2'], losses['policy'], losses['alpha']],
                                   g[f"sac.upd{u}.losses"], rtol=5e-5, atol=1e-6)
        assert float(agent.alpha.item()) == pytest.approx(float(g[f"sac.upd{u}.alpha"][0]), rel=1e-6)
        if u == 1:
            for tag, net, opt in (("policy", agent.policy, agent.policy_optimizer), ("q1", agent.q1, agent.q1_optimizer),
                                  ("q2", agent.q2, agent.q2_optimizer)):
                _check(g, f"sac.m1.{tag}", zip(net.P.p.keys(), opt.m), 99, rtol=1e-4, atol_of_absmax=1e-4)
        for tag, net in (("policy", agent.policy), ("q1", agent.q1), ("q2", agent.q2), ("q1t", agent.q1_target)):
            _check(g, f"sac.upd{u}.{tag}", net.state_dict().items(), u, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("graphed", [False, True])
def test_qmix_update_at_c3_size_matches_reference(graphed):
    from marllb_b200.policy import QMIXAgent
    g = load_golden("policy_fullsize")
    torch.manual_seed(0)
    agent = QMIXAgent(target_update_interval=2, **spec.QMIX)
    agent.graph_updates = graphed
    for i, n in enumerate(agent.agent_networks):
        _load_synth(n, 21 + i)
        agent.agent_networks_target[i].load_state_dict(n.state_dict())
    _load_synth(agent.mixer, 29)
    agent.mixer_target.load_state_dict(agent.mixer.state_dict())
    nets = [(f"ag{i}", n) for i, n in enumerate(agent.agent_networks)] + [("mixer", agent.mixer)]
    for u in range(1, spec.N_UPDATES + 1):
        stats = agent.update(batch=spec.qmix_batch(u))
        np.testing.assert_allclose([stats['loss'], stats['q_tot'], stats['target_q_tot']], g[f"qmix.upd{u}.stats"],
                                   rtol=5e-5, atol=1e-6)
        if u == 1:
            k = 0
            for tag, net in nets:                      # one optimiser over agents + mixer, bucket order (qmix_agent.py:108-113)
                names = list(net.P.p.keys())
                _check(g, f"qmix.m1.{tag}", zip(names, agent.optimizer.m[k:k + len(names)]), 99, rtol=1e-4, atol_of_absmax=1e-4)
                k += len(names)
        for tag, net in nets:
            _check(g, f"qmix.upd{u}.{tag}", net.state_dict().items(), u, rtol=1e-4, atol=2e-6)
