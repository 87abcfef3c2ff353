"""Generate tests/golden/paper_*.npz by running the reference's ORIGINAL-PAPER agents (torch CPU):
src/lb/sac_qmix.py (RNNAgent, QMix, QMix_Trainer, ReplayBufferGRU) and src/lb/sac_gru_discrete.py
(SoftQNetworkGRU, PolicyNetworkGRU, SAC_Trainer).

Those modules cannot be imported here (at import they load the VPP shared-memory layout and open
/dev/shm/shm_vip_1), so the class definitions are compiled out of the reference source,
unmodified, with `ast`, into a namespace that supplies the module-level names they use (`device`,
`DEBUG`, and the global `hidden_dim` the GRU layers read).

Run in the build container only:  python tests/golden/make_paper_golden.py
"""
import ast
import contextlib
import io
import math
import os
import pickle
import random
import sys
from os import path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.distributions import Categorical

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402


def load_classes(rel, names, hidden_dim):
    src = os.path.join(ref_import.REF_ROOT, rel)
    tree = ast.parse(open(src).read())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in names]
    assert len(keep) == len(names), [n.name for n in keep]
    ns = dict(torch=torch, nn=nn, F=F, optim=optim, np=np, Categorical=Categorical, random=random, path=path,
              pickle=pickle, math=math, device='cpu', DEBUG=False, hidden_dim=hidden_dim)
    exec(compile(ast.Module(body=keep, type_ignores=[]), src, "exec"), ns)
    return ns


def sd_np(prefix, sd):
    return {f"{prefix}{k}": v.detach().cpu().numpy().copy() for k, v in sd.items()}


def save(name, **arrs):
    p = os.path.join(HERE, name + ".npz")
    np.savez_compressed(p, **arrs)
    print(f"wrote {name}.npz ({os.path.getsize(p) / 1024:.1f} KiB)")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())      # the reference's forward() prints shapes


def qmix():
    H, A, Fd, heads, n, hyp, B, T = 32, 2, 6, 3, 3, 16, 4, 5
    ns = load_classes("src/lb/sac_qmix.py", ["ReplayBufferGRU", "RNNAgent", "QMix", "QMix_Trainer"], H)
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    out = dict(dims=np.array([H, A, Fd, heads, n, hyp, B, T]))
    tr = ns["QMix_Trainer"](None, A, Fd, heads, n, H, hyp, lr=0.001)
    out.update(sd_np("agent0.", tr.agent.state_dict()))
    out.update(sd_np("mixer0.", tr.mixer.state_dict()))
    state = rng.randn(B, T, A, Fd).astype(np.float32)
    next_state = rng.randn(B, T, A, Fd).astype(np.float32)
    action = rng.randint(0, n, (B, T, A, heads))
    last_action = rng.randint(0, n, (B, T, A, heads))
    reward = rng.rand(B, T).astype(np.float32)
    hidden_in = torch.as_tensor(rng.randn(1, B, A, H).astype(np.float32) * 0.1)
    out.update(state=state, next_state=next_state, action=action, last_action=last_action, reward=reward,
               hidden_in=hidden_in.numpy())
    # forward pieces
    with quiet(), torch.no_grad():
        qs, hid = tr.agent(torch.as_tensor(state), torch.as_tensor(last_action), hidden_in)
        chosen = torch.gather(qs, dim=-1, index=torch.as_tensor(action).unsqueeze(-1)).squeeze(-1)
        qtot = tr.mixer(chosen, torch.as_tensor(state))
        tl = tr._build_td_lambda_targets(torch.as_tensor(reward).unsqueeze(-1), qtot)
    out.update(fwd_qs=qs.numpy(), fwd_hidden=hid.numpy(), fwd_qtot=qtot.numpy(), fwd_td_lambda=tl.numpy())
    # deterministic get_action (sac_qmix.py:255-279)
    with quiet(), torch.no_grad():
        a_det, h_det = tr.agent.get_action(state[0, :, 0, :], last_action[0, :, 0, :],
                                           np.zeros((1, 1, H), np.float32), deterministic=True)
    out.update(get_action_det=np.asarray(a_det), get_action_hidden=h_det.numpy())

    class Buf:                                              # stands in for ReplayBufferGRU.sample
        def sample(self, batch_size):
            return hidden_in, state, action, last_action, reward, next_state
    tr.replay_buffer = Buf()
    losses = []
    crit = tr.criterion
    tr.criterion = lambda a, b: (losses.append(float(crit(a, b))), crit(a, b))[1]
    for k in range(2):
        with quiet():
            tr.update(B)
        out.update(sd_np(f"agent{k + 1}.", tr.agent.state_dict()))
        out.update(sd_np(f"mixer{k + 1}.", tr.mixer.state_dict()))
    out["losses"] = np.array(losses)
    # ReplayBufferGRU: ring + centre crop, under a fixed `random` seed
    buf = ns["ReplayBufferGRU"](3, "/nonexistent/replay.pkl")
    lens = [4, 6, 5, 7]
    for i, L in enumerate(lens):
        buf.push(torch.full((1, 1, A, H), float(i)), np.full((L, A, Fd), i, np.float32) + np.arange(L)[:, None, None],
                 np.full((L, A, heads), i), np.full((L, A, heads), i), np.arange(L, dtype=np.float32) + 10 * i,
                 np.full((L, A, Fd), -i, np.float32))
    random.seed(5)
    hi, s, a, la, r, ns_ = buf.sample(2)
    out.update(buf_lens=np.array(lens), buf_hidden=hi.numpy(), buf_state=np.asarray(s), buf_reward=np.asarray(r),
               buf_position=np.int64(buf.position))
    save("paper_qmix", **out)


def sac():
    H, Fd, heads, n, B, T = 32, 5, 3, 4, 4, 6
    ns = load_classes("src/lb/sac_gru_discrete.py", ["SoftQNetworkGRU", "PolicyNetworkGRU", "SAC_Trainer"], H)
    torch.manual_seed(1)
    rng = np.random.RandomState(1)
    out = dict(dims=np.array([H, Fd, heads, n, B, T]))
    tr = ns["SAC_Trainer"](None, Fd, n, H, heads)
    for tag, net in (("q1", tr.soft_q_net1), ("q2", tr.soft_q_net2), ("pi", tr.policy_net)):
        out.update(sd_np(f"{tag}_0.", net.state_dict()))
    state = rng.randn(B, T, Fd).astype(np.float32)
    next_state = rng.randn(B, T, Fd).astype(np.float32)
    action = rng.randint(0, n, (B, T, heads)).astype(np.float32)
    last_action = rng.randint(0, n, (B, T, heads)).astype(np.float32)
    reward = rng.rand(B, T).astype(np.float32)
    hidden_in = torch.as_tensor(rng.randn(1, B, H).astype(np.float32) * 0.1)
    hidden_out = torch.as_tensor(rng.randn(1, B, H).astype(np.float32) * 0.1)
    out.update(state=state, next_state=next_state, action=action, last_action=last_action, reward=reward,
               hidden_in=hidden_in.numpy(), hidden_out=hidden_out.numpy())
    with quiet(), torch.no_grad():
        probs, hid = tr.policy_net(torch.as_tensor(state), torch.as_tensor(last_action), hidden_in)
        q, qh = tr.soft_q_net1(torch.as_tensor(state), torch.as_tensor(action), hidden_in)
        a_det, _ = tr.policy_net.get_action(state[0], last_action[0], hidden_in[:, :1], deterministic=True)
    out.update(fwd_probs=probs.numpy(), fwd_hidden=hid.numpy(), fwd_q=q.numpy(), fwd_q_hidden=qh.numpy(),
               get_action_det=np.asarray(a_det))

    class Buf:
        def sample(self, batch_size):
            return hidden_in, hidden_out, state, action, last_action, reward, next_state
    tr.replay_buffer = Buf()
    sampled, logps = [], []
    ev = tr.policy_net.evaluate

    def recording_evaluate(*a, **k):                       # records what Categorical.sample() drew
        act, lp, h = ev(*a, **k)
        sampled.append(act.numpy().copy())
        logps.append(lp.detach().numpy().copy())
        return act, lp, h
    tr.policy_net.evaluate = recording_evaluate
    rets = []
    for k in range(2):
        with quiet():
            rets.append(float(tr.update(B)))
        for tag, net in (("q1", tr.soft_q_net1), ("q2", tr.soft_q_net2), ("pi", tr.policy_net),
                         ("t1", tr.target_soft_q_net1), ("t2", tr.target_soft_q_net2)):
            out.update(sd_np(f"{tag}_{k + 1}.", net.state_dict()))
        out[f"log_alpha_{k + 1}"] = tr.log_alpha.detach().numpy().copy()
    out.update(sampled=np.stack(sampled), logps=np.stack(logps), returns=np.array(rets))
    save("paper_sac", **out)


if __name__ == "__main__":
    qmix()
    sac()
