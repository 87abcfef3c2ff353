"""Generate tests/golden/policy_fullsize.npz: the UNMODIFIED reference learners (torch CPU) at the benchmarked sizes.

  C4: problem-04-sac-gru SAC_GRU_Agent.update_parameters, state 2816 / action 256 / hidden 256 / gru 128 / batch 256
  C3: problem-05-qmix   QMIXAgent.update, 2 agents / obs 352 / 32 actions / B = 32 episodes x T = 50

Inputs (parameters, batches, Gaussian noise) come from tests/fullsize_spec.py on both sides; the fixture keeps the
reference's losses, Adam first moments after the first update (= 0.1 x the gradient, per tensor) and the parameters
after each of N_UPDATES updates -- N_SAMPLES sampled entries per tensor plus each tensor's max-norm.

Run in the build container only:  python tests/golden/make_fullsize_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_import  # noqa: E402
import fullsize_spec as spec  # noqa: E402


def load_synth(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict({k: torch.as_tensor(v) for k, v in spec.synth_state_dict(shapes, seed).items()})


def record(out, tag, named_tensors, salt):
    for name, t in named_tensors:
        a = t.detach().cpu().numpy().reshape(-1)
        out[f"{tag}.{name}"] = a[spec.sample_index(name, a.size, salt)].copy()
        out[f"{tag}.{name}.absmax"] = np.array([np.abs(a).max()], np.float32)


def sac(out):
    sa = ref_import.load("sac_agent", "problem-04-sac-gru")
    c = spec.SAC
    torch.manual_seed(0)
    agent = sa.SAC_GRU_Agent(device="cpu", **c)
    for k, net in enumerate((agent.policy, agent.q1, agent.q2)):
        load_synth(net, 11 + k)
    agent.q1_target.load_state_dict(agent.q1.state_dict())
    agent.q2_target.load_state_dict(agent.q2.state_dict())
    agent.replay_buffer.is_ready = lambda n: True
    import torch.distributions.normal as tdn
    real = tdn._standard_normal
    for u in range(1, spec.N_UPDATES + 1):
        batch, eps_next, eps_new = spec.sac_batch(u)
        agent.replay_buffer.sample = lambda n, d, b=batch: tuple(torch.as_tensor(x) for x in b)
        queue = [torch.as_tensor(eps_next), torch.as_tensor(eps_new)]
        tdn._standard_normal = lambda shape, dtype, device: queue.pop(0).reshape(tuple(shape)).to(dtype)
        try:
            losses = agent.update_parameters(1)
        finally:
            tdn._standard_normal = real
        assert not queue, "the reference drew a different number of noise tensors"
        out[f"sac.upd{u}.losses"] = np.array([losses['q1'], losses['q2'], losses['policy'], losses['alpha']])
        out[f"sac.upd{u}.alpha"] = np.array([agent.alpha.item()])
        for tag, net in (("policy", agent.policy), ("q1", agent.q1), ("q2", agent.q2), ("q1t", agent.q1_target)):
            record(out, f"sac.upd{u}.{tag}", net.state_dict().items(), u)
        if u == 1:
            for tag, net, opt in (("policy", agent.policy, agent.policy_optimizer), ("q1", agent.q1, agent.q1_optimizer),
                                  ("q2", agent.q2, agent.q2_optimizer)):
                record(out, f"sac.m1.{tag}", [(n, opt.state[p]['exp_avg']) for n, p in net.named_parameters()], 99)
        print("SAC update", u, losses, flush=True)


def qmix(out):
    qa = ref_import.load("qmix_agent", "problem-05-qmix")
    c = spec.QMIX
    torch.manual_seed(0)
    agent = qa.QMIXAgent(device="cpu", target_update_interval=2, **c)
    for i, n in enumerate(agent.agent_networks):
        load_synth(n, 21 + i)
        agent.agent_networks_target[i].load_state_dict(n.state_dict())
    load_synth(agent.mixer, 29)
    agent.mixer_target.load_state_dict(agent.mixer.state_dict())
    agent.episode_buffer.is_ready = lambda n: True
    nets = [(f"ag{i}", n) for i, n in enumerate(agent.agent_networks)] + [("mixer", agent.mixer)]
    for u in range(1, spec.N_UPDATES + 1):
        batch = spec.qmix_batch(u)
        agent.episode_buffer.sample_batch = lambda n, m, b=batch: b
        stats = agent.update()
        out[f"qmix.upd{u}.stats"] = np.array([stats['loss'], stats['q_tot'], stats['target_q_tot']])
        for tag, net in nets:
            record(out, f"qmix.upd{u}.{tag}", net.state_dict().items(), u)
        if u == 1:
            for tag, net in nets:
                record(out, f"qmix.m1.{tag}", [(n, agent.optimizer.state[p]['exp_avg']) for n, p in net.named_parameters()], 99)
        print("QMIX update", u, stats, flush=True)


if __name__ == "__main__":
    out = {}
    sac(out)
    qmix(out)
    path = os.path.join(HERE, "policy_fullsize.npz")
    np.savez_compressed(path, **out)
    print(f"wrote policy_fullsize.npz ({os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays)")
