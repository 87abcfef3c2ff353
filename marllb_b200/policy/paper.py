"""The original MARLLB agents (the ones the paper evaluated) on the marllb_b200 policy kernels --
API mirror of the reference's testbed-side learners:

  ReplayBufferGRU     src/lb/sac_qmix.py:104-192, src/lb/sac_gru_discrete.py:46-124   episode replay (host)
  RNNAgent            src/lb/sac_qmix.py:195-279     linear-linear-GRU-linear-linear, softmax "Q" per head
  QMix                src/lb/sac_qmix.py:282-382     hypernet mixer with V(s) as the last-layer bias
  QMix_Trainer        src/lb/sac_qmix.py:385-470     TD(lambda = 0.6) targets, MSE, one Adam over agent + mixer
  SoftQNetworkGRU     src/lb/sac_gru_discrete.py:128-160
  PolicyNetworkGRU    src/lb/sac_gru_discrete.py:163-236   multi-head categorical policy
  SAC_Trainer         src/lb/sac_gru_discrete.py:241-359   discrete SAC with reward normalisation

Every forward AND backward runs on the hand-written kernels behind include/marllb_b200_policy.h
(mlb_gemm / mlb_linear_tc, GRU gates, softmax, categorical, mixer, TD(lambda), Adam); torch tensors
are device memory, torch initialisers are used once on the host in the reference's construction order
(same torch seed -> same weights), and state_dicts carry the reference's key names.

Two things the reference leaves to a module-level global or to torch's global generator are
arguments here: the GRU width (`hidden_dim = 128` at sac_qmix.py:844 / sac_gru_discrete.py -- here the
GRU is as wide as `hidden_size`, which is what that global is always set to) and the categorical draws
(`Categorical.sample()`: here an inverse-CDF draw from caller-supplied or torch-drawn uniforms, or
classes given by the caller; the deterministic path is np.argmax's first maximum).
"""
from __future__ import annotations

import pickle
import random
from os import path

import numpy as np
import torch

from . import ops
from .nn import Adam, FlatBucket, GRUCellSeq, Params, linear_backward
from .qmix import _device


class ReplayBufferGRU:
    """Episode replay of the original agents (sac_qmix.py:104-192; with `with_hidden_out=True` the
    7-field variant of sac_gru_discrete.py:46-124).  Host-side, same ring / centre-crop semantics."""

    def __init__(self, capacity, init_filename=None, logger=None, with_hidden_out=False):
        self.save2file = init_filename
        self.capacity = capacity
        self.with_hidden_out = with_hidden_out
        if init_filename and path.exists(init_filename):
            with open(init_filename, 'rb') as f:
                self.buffer, self.position = pickle.load(f)
            if logger:
                logger.info("replay_buffer: {}, {}".format(self.position, len(self.buffer)))
        else:
            self.buffer = []
            self.position = 0

    def push(self, *sample):
        if len(sample) != (7 if self.with_hidden_out else 6):
            raise TypeError("push(hidden_in, [hidden_out,] state, action, last_action, reward, next_state)")
        if len(self.buffer) < self.capacity:
            self.buffer.append(None)
        self.buffer[self.position] = tuple(sample)
        self.position = int((self.position + 1) % self.capacity)           # ring buffer

    def sample(self, batch_size):
        batch = random.sample(self.buffer, batch_size)
        nh = 2 if self.with_hidden_out else 1
        min_seq_len = min(len(s[nh]) for s in batch)
        # h_in: (1, 1, n_agents, hidden) -> cat along dim -3 (QMIX); (1, 1, hidden) -> dim -2 (SAC)
        cat_dim = -2 if self.with_hidden_out else -3
        hid = [torch.cat([torch.as_tensor(s[i]) for s in batch], dim=cat_dim).detach() for i in range(nh)]
        cols = [[] for _ in range(5)]
        for s in batch:
            n = len(s[nh])
            start = int((n - min_seq_len) / 2)                             # centre crop to the shortest episode
            for c in range(5):
                cols[c].append(s[nh + c][start:start + min_seq_len])
        return (*hid, *cols)

    def __len__(self):
        return len(self.buffer)

    def get_length(self):
        return len(self.buffer)

    def dump_buffer(self):
        with open(self.save2file, 'wb') as f:
            pickle.dump([self.buffer, self.position], f)


class _GRUTrunk:
    """linear1 -> ReLU -> linear2 -> ReLU -> nn.GRU(H, H) (sequence-major) -> post layers: the body shared by
    RNNAgent (sac_qmix.py:212-216,238-247), SoftQNetworkGRU and PolicyNetworkGRU (sac_gru_discrete.py:137-141,
    170-176).  `post` = [(name, out_dim, relu?)]."""

    def __init__(self, in_dim, hidden, post, device, final_init_w=None):
        self.device, self.hidden, self.post = device, hidden, post
        # the reference's construction order: linear1, linear2, rnn, then the post layers (default torch inits)
        lin1 = torch.nn.Linear(in_dim, hidden)
        lin2 = torch.nn.Linear(hidden, hidden)
        rnn = torch.nn.GRU(hidden, hidden)
        posts, d = [], hidden
        for name, out, _ in post:
            posts.append(torch.nn.Linear(d, out))
            d = out
        if final_init_w is not None:                                       # sac_gru_discrete.py:143-144
            posts[-1].weight.data.uniform_(-final_init_w, final_init_w)
            posts[-1].bias.data.uniform_(-final_init_w, final_init_w)
        self.P = Params(device)
        for n, lin in (("linear1", lin1), ("linear2", lin2)):
            self.P.add(n + ".weight", lin.weight)
            self.P.add(n + ".bias", lin.bias)
        for name, p in rnn.named_parameters():
            self.P.add("rnn." + name, p)
        for (name, _, _), lin in zip(post, posts):
            self.P.add(name + ".weight", lin.weight)
            self.P.add(name + ".bias", lin.bias)
        self.gru = GRUCellSeq(self.P, prefix="rnn.")
        self.out_dim = d
        self._tape = None

    def forward_seq(self, x, h0, save=False):
        """x [T, N, in] (contiguous), h0 [N, H] -> (y [T, N, out], h_T [N, H])."""
        P = self.P.p
        T, N, In = x.shape
        x2 = x.reshape(T * N, In)
        a1 = ops.linear(x2, P["linear1.weight"], P["linear1.bias"], ops.ACT_RELU)
        a2 = ops.linear(a1, P["linear2.weight"], P["linear2.bias"], ops.ACT_RELU)
        hs, tape = self.gru.forward_seq(a2.reshape(T, N, self.hidden), h0)
        acts, cur = [], hs.reshape(T * N, self.hidden)
        for name, _, relu in self.post:
            acts.append(cur)
            cur = ops.linear(cur, P[name + ".weight"], P[name + ".bias"], ops.ACT_RELU if relu else ops.ACT_NONE)
        if save:
            self._tape = (x2, a1, a2, tape, acts, cur)
        return cur.reshape(T, N, self.out_dim), hs[T - 1]

    def backward_seq(self, dy):
        """dy [T, N, out]: accumulates the gradients of every parameter (BPTT through the GRU)."""
        x2, a1, a2, tape, acts, y = self._tape
        T, N, _ = dy.shape
        d = dy.reshape(T * N, self.out_dim).contiguous()
        outs = acts[1:] + [y]
        for (name, _, relu), inp, out in zip(reversed(self.post), reversed(acts), reversed(outs)):
            if relu:
                d = ops.relu_backward(out, d)
            d = linear_backward(self.P, name + ".weight", name + ".bias", inp, d)
        dxs, _ = self.gru.backward_seq(d.reshape(T, N, self.hidden), tape, need_dx=True)
        da2 = ops.relu_backward(a2, dxs.reshape(T * N, self.hidden))
        da1 = ops.relu_backward(a1, linear_backward(self.P, "linear2.weight", "linear2.bias", a1, da2))
        linear_backward(self.P, "linear1.weight", "linear1.bias", x2, da1, need_dx=False)
        self._tape = None


class _Module:
    """state_dict / parameters plumbing shared by the network classes below."""

    def to(self, device):
        return self

    def state_dict(self):
        return self.P.state_dict()

    def load_state_dict(self, sd):
        self.P.load_state_dict(sd)

    def parameters(self):
        return self.P.tensors()


def _uniforms(shape, u, device):
    if u is None:
        return torch.rand(shape, device=device, dtype=torch.float32)
    return torch.as_tensor(u, dtype=torch.float32).to(device).reshape(shape).contiguous()


class RNNAgent(_Module):
    """sac_qmix.py:195-279: one network shared by all agents; per (agent, head) a softmax over `num_actions`
    discretised weights whose probabilities are used as Q values."""

    def __init__(self, num_inputs, num_heads, num_actions, hidden_size, device=None):
        self.num_inputs, self.num_heads, self.num_actions = num_inputs, num_heads, num_actions
        self.hidden_size = hidden_size
        self.device = _device(device)
        self.trunk = _GRUTrunk(num_inputs + num_heads * num_actions, hidden_size,
                               [("linear3", hidden_size, True), ("linear4", num_heads * num_actions, False)], self.device)
        self.P = self.trunk.P

    def forward(self, state, action, hidden_in, save=False):
        """state [B, T, A, F], action [B, T, A, heads] (class indices), hidden_in reshapeable to [B*A, H]
        -> (qs [B, T, A, heads, actions], hidden [1, B*A, H])                        sac_qmix.py:218-253"""
        dev = self.device
        state = torch.as_tensor(state, dtype=torch.float32).to(dev)
        B, T, A, F = state.shape
        H, n, heads = self.hidden_size, self.num_actions, self.num_heads
        st = state.permute(1, 0, 2, 3).reshape(T * B * A, F).contiguous()
        act = torch.as_tensor(action).to(dev).permute(1, 0, 2, 3).reshape(T * B * A, heads).to(torch.int32).contiguous()
        h0 = torch.as_tensor(hidden_in, dtype=torch.float32).to(dev).reshape(B * A, H).contiguous()
        x = ops.concat_onehot(st, act, n).reshape(T, B * A, F + heads * n)
        logits, hT = self.trunk.forward_seq(x, h0, save=save)
        probs = ops.softmax_forward(logits.reshape(T, B, A, heads, n))               # :250
        if save:
            self._probs = probs
        return probs.permute(1, 0, 2, 3, 4), hT.reshape(1, B * A, H)

    __call__ = forward

    def backward(self, dqs):
        """dqs [B, T, A, heads, actions]: gradient w.r.t. the softmax outputs of the last forward(save=True)."""
        p = self._probs
        T, B, A, heads, n = p.shape
        dlogits = ops.softmax_backward(p, dqs.permute(1, 0, 2, 3, 4).contiguous())
        self.trunk.backward_seq(dlogits.reshape(T, B * A, heads * n))
        self._probs = None

    def get_action(self, state, last_action, hidden_in, deterministic=False, u=None):
        """state [#batch, F], last_action [#batch, heads] -> (action [#batch, heads] numpy, hidden_out)
        sac_qmix.py:255-279 (the batch rides on the sequence axis there too).  `u`: optional pre-drawn uniforms
        [#batch, heads] standing in for Categorical.sample()'s use of torch's global generator."""
        state = torch.as_tensor(np.asarray(state), dtype=torch.float32)[None, :, None, :]
        last_action = torch.as_tensor(np.asarray(last_action), dtype=torch.int64)[None, :, None, :]
        hidden_in = torch.as_tensor(hidden_in, dtype=torch.float32)
        qs, hidden_out = self.forward(state, last_action, hidden_in)
        hidden_out = hidden_out.squeeze(-3)                                          # remove the agent dim :266
        probs = qs.squeeze(-3).contiguous()                                          # [1, #batch, heads, actions] :267
        uu = None if deterministic else _uniforms(probs.shape[:-1], u, self.device)
        act, _, _ = ops.categorical(probs, u=uu, want_logp=False)
        a = act.cpu().numpy().astype(np.int64)
        return (a if deterministic else a.squeeze()), hidden_out                     # :271-274


class QMix(_Module):
    """sac_qmix.py:282-382: two-layer hypernetworks, |.| on the mixing weights, V(s) instead of a last bias."""

    def __init__(self, state_dim, n_agents, num_heads, embed_dim=64, hypernet_embed=128, abs=True, device=None):
        self.n_agents, self.num_heads = n_agents, num_heads
        self.state_dim = state_dim * n_agents                                        # :291
        self.embed_dim, self.hypernet_embed, self.abs = embed_dim, hypernet_embed, abs
        self.device = _device(device)
        S, E, Hh = self.state_dim, embed_dim, hypernet_embed
        spec = [("hyper_w_1.0", S, Hh), ("hyper_w_1.2", Hh, num_heads * E * n_agents),     # construction order :302-319
                ("hyper_w_final.0", S, Hh), ("hyper_w_final.2", Hh, E),
                ("hyper_b_1", S, E), ("V.0", S, E), ("V.2", E, 1)]
        self.P = Params(self.device)
        for name, i, o in spec:
            lin = torch.nn.Linear(i, o)
            self.P.add(name + ".weight", lin.weight)
            self.P.add(name + ".bias", lin.bias)

    def forward(self, agent_qs, states, save=False):
        """agent_qs [B, T, A, heads], states [B, T, A, F] -> q_tot [B, T, 1]        sac_qmix.py:323-361"""
        P, dev = self.P.p, self.device
        bs = agent_qs.shape[0]
        s = torch.as_tensor(states, dtype=torch.float32).to(dev).reshape(-1, self.state_dim).contiguous()
        q = agent_qs.to(dev, torch.float32).reshape(-1, self.num_heads * self.n_agents).contiguous()
        h1 = ops.linear(s, P["hyper_w_1.0.weight"], P["hyper_w_1.0.bias"], ops.ACT_RELU)
        pre1 = ops.linear(h1, P["hyper_w_1.2.weight"], P["hyper_w_1.2.bias"])
        w1 = ops.abs_forward(pre1) if self.abs else pre1
        b1 = ops.linear(s, P["hyper_b_1.weight"], P["hyper_b_1.bias"])
        h2 = ops.linear(s, P["hyper_w_final.0.weight"], P["hyper_w_final.0.bias"], ops.ACT_RELU)
        pre2 = ops.linear(h2, P["hyper_w_final.2.weight"], P["hyper_w_final.2.bias"])
        w2 = ops.abs_forward(pre2) if self.abs else pre2
        h3 = ops.linear(s, P["V.0.weight"], P["V.0.bias"], ops.ACT_RELU)
        v = ops.linear(h3, P["V.2.weight"], P["V.2.bias"])
        q_tot, hidden = ops.mixer_forward(q, w1, b1, w2, v.reshape(-1), save_hidden=save)
        if save:
            self._tape = (q, s, h1, pre1, w1, h2, pre2, w2, h3, hidden)
        return q_tot.reshape(bs, -1, 1)

    __call__ = forward

    def backward(self, dq_tot):
        """dq_tot [B, T, 1] -> d agent_qs [B*T, A*heads]; accumulates the hypernet gradients."""
        q, s, h1, pre1, w1, h2, pre2, w2, h3, hidden = self._tape
        dq, dw1, db1, dw2, dv = ops.mixer_backward(dq_tot.reshape(-1).contiguous(), q, w1, w2, hidden)
        P = self.P
        dpre1 = ops.abs_backward(pre1, dw1) if self.abs else dw1
        dh1 = ops.relu_backward(h1, linear_backward(P, "hyper_w_1.2.weight", "hyper_w_1.2.bias", h1, dpre1))
        linear_backward(P, "hyper_w_1.0.weight", "hyper_w_1.0.bias", s, dh1, need_dx=False)
        linear_backward(P, "hyper_b_1.weight", "hyper_b_1.bias", s, db1, need_dx=False)
        dpre2 = ops.abs_backward(pre2, dw2) if self.abs else dw2
        dh2 = ops.relu_backward(h2, linear_backward(P, "hyper_w_final.2.weight", "hyper_w_final.2.bias", h2, dpre2))
        linear_backward(P, "hyper_w_final.0.weight", "hyper_w_final.0.bias", s, dh2, need_dx=False)
        dh3 = ops.relu_backward(h3, linear_backward(P, "V.2.weight", "V.2.bias", h3, dv))
        linear_backward(P, "V.0.weight", "V.0.bias", s, dh3, need_dx=False)
        self._tape = None
        return dq


class QMix_Trainer:
    """sac_qmix.py:385-470."""

    def __init__(self, replay_buffer, n_agents, state_dim, num_heads, action_dim, hidden_dim, hypernet_dim,
                 lr=0.001, logger=None, device=None, data_parallel=True):
        self.replay_buffer = replay_buffer
        self.device = _device(device)
        self.n_agents, self.num_heads, self.action_dim = n_agents, num_heads, action_dim
        self.agent = RNNAgent(state_dim, num_heads, action_dim, hidden_dim, self.device)
        self.target_agent = RNNAgent(state_dim, num_heads, action_dim, hidden_dim, self.device)
        self.mixer = QMix(state_dim, n_agents, num_heads, hidden_dim, hypernet_dim, device=self.device)
        self.target_mixer = QMix(state_dim, n_agents, num_heads, hidden_dim, hypernet_dim, device=self.device)
        self._update_targets()
        self._bucket = FlatBucket([self.agent.P, self.mixer.P])                      # one optimiser: :407-408
        self.optimizer = Adam(self._bucket, lr, data_parallel=data_parallel)
        self.logger = logger

    def update(self, batch_size, batch=None):
        """One TD(lambda) QMIX step on a sampled batch (or on `batch`, the 6-tuple ReplayBufferGRU.sample
        returns).  Returns {'loss', 'q_tot', 'targets'} (the reference returns nothing)."""
        dev, f32 = self.device, torch.float32
        hidden_in, state, action, last_action, reward, next_state = batch if batch is not None \
            else self.replay_buffer.sample(batch_size)
        state = torch.as_tensor(np.asarray(state), dtype=f32).to(dev)                # [B, T, A, F]
        next_state = torch.as_tensor(np.asarray(next_state), dtype=f32).to(dev)
        action = torch.as_tensor(np.asarray(action)).to(dev).to(torch.int32)         # [B, T, A, heads]
        last_action = torch.as_tensor(np.asarray(last_action)).to(dev).to(torch.int32)
        reward = torch.as_tensor(np.asarray(reward), dtype=f32).to(dev)              # [B, T]
        B, T, A, _ = state.shape
        heads, n = self.num_heads, self.action_dim

        agent_outs, _ = self.agent.forward(state, last_action, hidden_in, save=True)           # :428
        act_c = action.contiguous()
        _, _, chosen = ops.categorical(agent_outs.contiguous(), given=act_c, want_logp=False, want_p=True)   # :431-432
        qtot = self.mixer.forward(chosen, state, save=True)                                    # :433
        target_outs, _ = self.target_agent.forward(next_state, action, hidden_in)              # :436
        tmax, _ = ops.row_max(target_outs.reshape(-1, n).contiguous())                         # :438
        target_qtot = self.target_mixer.forward(tmax.reshape(B, T, A, heads), next_state)      # :439
        targets = ops.td_lambda_targets(reward.reshape(B, T).contiguous(),
                                        target_qtot.reshape(B, T).contiguous())                # :441
        loss, dq = ops.mse_loss(qtot.reshape(-1).contiguous(), targets.reshape(-1))            # :443

        self._bucket.flat_g.zero_()                                                            # :445
        dchosen = self.mixer.backward(dq.reshape(B, T, 1))                                     # [B*T, A*heads]
        dqs = ops.scatter_class(dchosen.reshape(B, T, A, heads).contiguous(), act_c, n)
        self.agent.backward(dqs)
        self.optimizer.step()                                                                  # :447
        return {"loss": float(loss.item()), "q_tot": qtot, "targets": targets.reshape(B, T, 1)}

    def _build_td_lambda_targets(self, rewards, target_qs, gamma=0.99, td_lambda=0.6):
        """rewards [B, T, 1], target_qs [B, T, 1] -> [B, T, 1]                       sac_qmix.py:449-460"""
        B, T = target_qs.shape[0], target_qs.shape[1]
        r = torch.as_tensor(rewards, dtype=torch.float32).to(self.device).reshape(B, T).contiguous()
        q = torch.as_tensor(target_qs, dtype=torch.float32).to(self.device).reshape(B, T).contiguous()
        return ops.td_lambda_targets(r, q, gamma, td_lambda).reshape(B, T, 1)

    def _update_targets(self):
        self.target_mixer.load_state_dict(self.mixer.state_dict())                   # :462-466
        self.target_agent.load_state_dict(self.agent.state_dict())

    def save_model(self, path):
        torch.save({k: v.cpu() for k, v in self.agent.state_dict().items()}, path + '_agent')   # :468-470
        torch.save({k: v.cpu() for k, v in self.mixer.state_dict().items()}, path + '_mixer')

    def load_model(self, path):
        self.agent.load_state_dict(torch.load(path + '_agent'))
        self.mixer.load_state_dict(torch.load(path + '_mixer'))


# ------------------------------------------------------------------------------------------------
class SoftQNetworkGRU(_Module):
    """sac_gru_discrete.py:128-160: Q(state, action) with the heads' class indices as float inputs."""

    def __init__(self, num_inputs, num_heads, hidden_size, init_w=3e-3, device=None):
        self.device = _device(device)
        self.hidden_size = hidden_size
        self.trunk = _GRUTrunk(num_inputs + num_heads, hidden_size,
                               [("linear3", hidden_size, True), ("linear4", 1, False)], self.device, final_init_w=init_w)
        self.P = self.trunk.P

    def forward(self, state, action, hidden_in, save=False):
        """state [B, T, F], action [B, T, heads] (float), hidden_in [1, B, H] -> (q [B, T, 1], hidden [1, B, H])"""
        dev = self.device
        state = torch.as_tensor(state, dtype=torch.float32).to(dev)
        action = torch.as_tensor(action).to(dev).to(torch.float32)
        B, T, _ = state.shape
        x = torch.cat([state.permute(1, 0, 2), action.permute(1, 0, 2)], -1).contiguous()     # :151-153
        h0 = torch.as_tensor(hidden_in, dtype=torch.float32).to(dev).reshape(B, self.hidden_size).contiguous()
        y, hT = self.trunk.forward_seq(x, h0, save=save)
        return y.permute(1, 0, 2), hT.reshape(1, B, self.hidden_size)

    __call__ = forward

    def backward(self, dq):
        """dq [B, T, 1]."""
        self.trunk.backward_seq(dq.permute(1, 0, 2).contiguous())


class PolicyNetworkGRU(_Module):
    """sac_gru_discrete.py:163-236: per head a categorical over `num_actions` weight levels."""

    def __init__(self, num_inputs, num_actions, hidden_size, num_heads, logger=None, device=None):
        self.logger = logger
        self.num_actions, self.num_heads, self.hidden_size = num_actions, num_heads, hidden_size
        self.device = _device(device)
        self.trunk = _GRUTrunk(num_inputs + num_heads, hidden_size,
                               [("linear3", hidden_size, True), ("linear4", hidden_size, True),
                                ("output", num_actions * num_heads, False)], self.device)
        self.P = self.trunk.P

    def forward(self, state, last_action, hidden_in, softmax_dim=-1, save=False):
        """state [B, T, F], last_action [B, T, heads] -> (probs [B, T, heads, actions], hidden [1, B, H])"""
        if softmax_dim != -1:
            raise ValueError("only softmax_dim=-1 (the reference's only use) is built")
        dev = self.device
        state = torch.as_tensor(state, dtype=torch.float32).to(dev)
        last_action = torch.as_tensor(last_action).to(dev).to(torch.float32)
        B, T, _ = state.shape
        x = torch.cat([state.permute(1, 0, 2), last_action.permute(1, 0, 2)], -1).contiguous()
        h0 = torch.as_tensor(hidden_in, dtype=torch.float32).to(dev).reshape(B, self.hidden_size).contiguous()
        logits, hT = self.trunk.forward_seq(x, h0, save=save)
        probs = ops.softmax_forward(logits.reshape(T, B, self.num_heads, self.num_actions))    # :189
        if save:
            self._probs = probs
        return probs.permute(1, 0, 2, 3), hT.reshape(1, B, self.hidden_size)

    __call__ = forward

    def evaluate(self, state, last_action, hidden_in, epsilon=1e-6, u=None, given=None, save=False):
        """-> (action [B, T, heads] int64, log_probs [B, T, 1] = sum over heads, hidden_out)   :193-211
        `u` / `given`: pre-drawn uniforms or classes [B, T, heads] in place of Categorical.sample()."""
        probs, hidden_out = self.forward(state, last_action, hidden_in, save=save)
        p = probs.contiguous()                                                       # [B, T, heads, n]
        g = None if given is None else torch.as_tensor(given).to(self.device).to(torch.int32).contiguous()
        uu = None if g is not None else _uniforms(p.shape[:-1], u, self.device)
        act, logp, _ = ops.categorical(p, u=uu, given=g)
        ones = torch.ones((1, self.num_heads), dtype=torch.float32, device=self.device)
        B, T = p.shape[0], p.shape[1]
        log_probs = ops.linear(logp.reshape(B * T, self.num_heads), ones).reshape(B, T, 1)     # sum over heads :208
        if save:
            self._act = act
        return act.to(torch.int64), log_probs, hidden_out

    def backward_logprob(self, g):
        """g [B, T, 1]: gradient w.r.t. log_probs of the last evaluate(save=True)."""
        p, act = self._probs, self._act                                              # p [T,B,heads,n], act [B,T,heads]
        T, B, heads, n = p.shape
        act_t = act.permute(1, 0, 2).contiguous()
        g_t = g.reshape(B, T).permute(1, 0).contiguous()
        dlogits = ops.logprob_backward(p, act_t, g_t, heads)
        self.trunk.backward_seq(dlogits.reshape(T, B, heads * n))
        self._probs = self._act = None

    def get_action(self, state, last_action, hidden_in, deterministic, u=None):
        """state [#batch, F], last_action [#batch, heads] -> (action numpy, hidden_out)       :213-236"""
        state = torch.as_tensor(np.asarray(state), dtype=torch.float32).unsqueeze(0)
        last_action = torch.as_tensor(np.asarray(last_action), dtype=torch.float32).unsqueeze(0)
        probs, hidden_out = self.forward(state, last_action, hidden_in)
        p = probs.contiguous()
        uu = None if deterministic else _uniforms(p.shape[:-1], u, self.device)
        act, _, _ = ops.categorical(p, u=uu, want_logp=False)
        a = act.cpu().numpy().astype(np.int64)
        return (a if deterministic else a.squeeze()), hidden_out


class SAC_Trainer:
    """sac_gru_discrete.py:241-359."""

    def __init__(self, replay_buffer, state_dim, action_dim, hidden_dim, head_dim, logger=None, device=None,
                 data_parallel=True):
        self.replay_buffer = replay_buffer
        dev = self.device = _device(device)
        self.soft_q_net1 = SoftQNetworkGRU(state_dim, head_dim, hidden_dim, device=dev)
        self.soft_q_net2 = SoftQNetworkGRU(state_dim, head_dim, hidden_dim, device=dev)
        self.target_soft_q_net1 = SoftQNetworkGRU(state_dim, head_dim, hidden_dim, device=dev)
        self.target_soft_q_net2 = SoftQNetworkGRU(state_dim, head_dim, hidden_dim, device=dev)
        self.policy_net = PolicyNetworkGRU(state_dim, action_dim, hidden_dim, head_dim, logger=logger, device=dev)
        self.log_alpha = torch.zeros(1, dtype=torch.float32, device=dev)
        self._log_alpha_g = torch.zeros(1, dtype=torch.float32, device=dev)
        self.target_soft_q_net1.load_state_dict(self.soft_q_net1.state_dict())       # :262-265
        self.target_soft_q_net2.load_state_dict(self.soft_q_net2.state_dict())
        lr = 3e-4                                                                    # :270-272
        self._b_q1, self._b_q2 = FlatBucket([self.soft_q_net1.P]), FlatBucket([self.soft_q_net2.P])
        self._b_pi = FlatBucket([self.policy_net.P])
        self._b_alpha = FlatBucket([], extra=[(self.log_alpha, self._log_alpha_g)])
        self.log_alpha, self._log_alpha_g = self._b_alpha.params[0], self._b_alpha.grads[0]
        self._b_t1, self._b_t2 = FlatBucket([self.target_soft_q_net1.P]), FlatBucket([self.target_soft_q_net2.P])
        self.soft_q_optimizer1 = Adam(self._b_q1, lr, data_parallel=data_parallel)
        self.soft_q_optimizer2 = Adam(self._b_q2, lr, data_parallel=data_parallel)
        self.policy_optimizer = Adam(self._b_pi, lr, data_parallel=data_parallel)
        self.alpha_optimizer = Adam(self._b_alpha, lr, data_parallel=data_parallel)
        self.alpha = 1.0

    def update(self, batch_size, reward_scale=10., auto_entropy=True, target_entropy=-2, gamma=0.99, soft_tau=1e-2,
               batch=None, sampled=None, u=None):
        """One discrete-SAC step (alpha, both critics, policy, soft target update), in the reference's order.
        `batch`: the 7-tuple ReplayBufferGRU(with_hidden_out=True).sample returns; `sampled` = (new_action,
        new_next_action) classes [B, T, heads] or `u` = the same pair as uniforms, standing in for the two
        Categorical.sample() calls.  Returns predicted_new_q_value.mean() like the reference (as a float)."""
        dev, f32 = self.device, torch.float32
        hidden_in, hidden_out, state, action, last_action, reward, next_state = batch if batch is not None \
            else self.replay_buffer.sample(batch_size)
        state = torch.as_tensor(np.asarray(state), dtype=f32).to(dev)                # [B, T, F]
        next_state = torch.as_tensor(np.asarray(next_state), dtype=f32).to(dev)
        action = torch.as_tensor(np.asarray(action), dtype=f32).to(dev)
        last_action = torch.as_tensor(np.asarray(last_action), dtype=f32).to(dev)
        reward = torch.as_tensor(np.asarray(reward), dtype=f32).to(dev)              # [B, T]
        B, T, _ = state.shape
        M = B * T
        g0, g1 = sampled if sampled is not None else (None, None)
        u0, u1 = u if u is not None else (None, None)

        q1, _ = self.soft_q_net1.forward(state, action, hidden_in, save=True)        # :292-293
        q2, _ = self.soft_q_net2.forward(state, action, hidden_in, save=True)
        new_action, log_prob, _ = self.policy_net.evaluate(state, last_action, hidden_in, u=u0, given=g0, save=True)
        pol_tape = (self.policy_net.trunk._tape, self.policy_net._probs, self.policy_net._act)
        new_next_action, next_log_prob, _ = self.policy_net.evaluate(next_state, action, hidden_out, u=u1, given=g1)
        rn = ops.reward_normalize(reward.reshape(B, T).contiguous(), reward_scale)   # :299-300

        if auto_entropy is True:                                                     # :303-309
            _, dla = ops.sac_alpha_loss(log_prob.reshape(M).contiguous(), self.log_alpha, float(target_entropy))
            self._log_alpha_g.copy_(dla)
            self.alpha_optimizer.step()
            alpha_t = ops.exp_scalar(self.log_alpha)
        else:
            alpha_t = torch.ones(1, dtype=f32, device=dev)
        self.alpha = alpha_t

        nna = new_next_action.to(f32)
        tq1, _ = self.target_soft_q_net1.forward(next_state, nna, hidden_out)        # :315-318
        tq2, _ = self.target_soft_q_net2.forward(next_state, nna, hidden_out)
        y = ops.dsac_q_target(rn.reshape(M), tq1.reshape(M).contiguous(), tq2.reshape(M).contiguous(),
                              next_log_prob.reshape(M).contiguous(), alpha_t, gamma)           # :319-321
        for net, q, bucket, opt in ((self.soft_q_net1, q1, self._b_q1, self.soft_q_optimizer1),
                                    (self.soft_q_net2, q2, self._b_q2, self.soft_q_optimizer2)):
            _, dq = ops.mse_loss(q.reshape(M).contiguous(), y)                       # :322-332
            bucket.flat_g.zero_()
            net.backward(dq.reshape(B, T, 1))
            opt.step()

        na = new_action.to(f32)
        pq1, _ = self.soft_q_net1.forward(state, na, hidden_in)                      # :335-338 (updated critics)
        pq2, _ = self.soft_q_net2.forward(state, na, hidden_in)
        # policy_loss = mean(alpha * log_prob - min(q1, q2)); the sampled classes carry no gradient, so only the
        # alpha * log_prob term reaches the policy parameters                        :343
        _, dlp, _, _ = ops.sac_policy_loss(log_prob.reshape(M).contiguous(), pq1.reshape(M).contiguous(),
                                           pq2.reshape(M).contiguous(), alpha_t)
        self._b_pi.flat_g.zero_()
        self.policy_net.trunk._tape, self.policy_net._probs, self.policy_net._act = pol_tape
        self.policy_net.backward_logprob(dlp.reshape(B, T, 1))
        self.policy_optimizer.step()                                                 # :345-347

        ops.axpby(soft_tau, self._b_q1.flat_p, 1.0 - soft_tau, self._b_t1.flat_p)    # :350-357
        ops.axpby(soft_tau, self._b_q2.flat_p, 1.0 - soft_tau, self._b_t2.flat_p)
        return float(torch.minimum(pq1, pq2).mean().item())                          # :358

    def save_model(self, path):
        torch.save({k: v.cpu() for k, v in self.soft_q_net1.state_dict().items()}, path + '_q1')   # :360-363
        torch.save({k: v.cpu() for k, v in self.soft_q_net2.state_dict().items()}, path + '_q2')
        torch.save({k: v.cpu() for k, v in self.policy_net.state_dict().items()}, path + '_policy')

    def load_model(self, path):
        self.soft_q_net1.load_state_dict(torch.load(path + '_q1'))
        self.soft_q_net2.load_state_dict(torch.load(path + '_q2'))
        self.policy_net.load_state_dict(torch.load(path + '_policy'))
