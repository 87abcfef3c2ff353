"""Device-resident rollout loops: the batched env step fused with batched policy inference.

Config C3 of BASELINE.json: E envs x A LB agents, QMIX action selection for every agent of every
env in one batch (reference: QMIXAgent.select_actions, problem-05-qmix/src/qmix_agent.py:126-170,
called once per env per step by the reference driver), the chosen server index of each agent turned
into the env's per-server action (rl_controller.py:314-321) and the env stepped -- observations,
hidden states, actions and rewards never leave the GPU.
"""
from __future__ import annotations

import torch

from .policy import ops
from .vec_env import VecLoadBalanceEnv


class QMIXRollout:
    """obs (E,S,11) -> per-agent obs [E, A, Sa*11] (the clean layout of SURVEY App. C #2) ->
    AgentQNetwork forward + epsilon-greedy -> one-hot env action -> env step."""

    def __init__(self, env: VecLoadBalanceEnv, agent, hot: int = 2, cold: int = 0):
        if env.num_agents != agent.num_agents:
            raise ValueError("env and agent disagree on the number of agents")
        self.env, self.agent = env, agent
        self.E, self.A = env.num_envs, env.num_agents
        self.Sa = env.total_servers // env.num_agents
        if agent.obs_dim != self.Sa * 11:
            raise ValueError(f"agent.obs_dim must be servers_per_agent*11 = {self.Sa * 11}")
        if agent.action_dim != self.Sa:
            raise ValueError("agent.action_dim must equal servers_per_agent (one server index per agent)")
        self.hot, self.cold = hot, cold
        self.hidden = None
        self._env_action = torch.empty((self.E, self.A * self.Sa), dtype=torch.uint8, device=env.device)

    def reset(self):
        self.hidden = None
        return self.env.reset()

    # ---- CUDA graph: the ~17 launches of a rollout step replayed as one graph launch
    def capture(self, epsilon: float = 0.0, warmup: int = 2):
        """Capture step(epsilon, u, rnd) into a CUDA graph.  The exploration draws are read from the
        static buffers `graph_u` [E, A] float32 / `graph_rnd` [E, A] int32 (fill them before each
        replay; with epsilon == 0 they are unused) and the GRU hidden state lives in a static
        buffer.  Note: the `warmup` + 1 steps run here advance the envs."""
        dev = self.env.device
        self.graph_u = torch.zeros((self.E, self.A), dtype=torch.float32, device=dev)
        self.graph_rnd = torch.zeros((self.E, self.A), dtype=torch.int32, device=dev)
        if self.hidden is None:
            self.hidden = torch.zeros((self.A, self.E, self.agent.agent_networks[0].gru_dim), dtype=torch.float32, device=dev)
        self._h_static = self.hidden.clone()

        def body():
            self.hidden = self._h_static
            out = self.step(epsilon, self.graph_u, self.graph_rnd)
            self._h_static.copy_(self.hidden)
            self.hidden = self._h_static
            return out

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from .policy import ops as _ops
        n0 = self.env.launch_count + _ops.LAUNCHES
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_out = body()
        self.graph_launches = self.env.launch_count + _ops.LAUNCHES - n0 + 1   # + the hidden-state copy
        return self

    def step_graph(self):
        """Replay the captured step (inputs: graph_u / graph_rnd).  Returns (obs, reward, done, actions)."""
        self._graph.replay()
        return self._graph_out

    def step(self, epsilon: float = 0.0, u=None, rnd=None):
        """One rollout step for every env.  u [E, A] uniforms / rnd [E, A] int32 random actions
        pre-drawn by the caller (None = greedy).  Returns (obs, reward, done, actions [E, A])."""
        obs = self.env.obs.view(self.E, self.A, self.Sa * 11)
        act, self.hidden, _ = self.agent.select_actions_batch(obs, self.hidden, epsilon, u, rnd)
        ops.onehot_action(act, self.Sa, self.hot, self.cold, out=self._env_action)
        o, r, d = self.env.step(self._env_action)
        return o, r, d, act


class DeviceReplay:
    """Replay ring held on the GPU for the batched rollout (the reference's ReplayBuffer,
    problem-04-sac-gru/src/replay_buffer.py:35-94, stores one transition per Python call in a host
    deque; at thousands of transitions per step that would be the bottleneck).  Same content per
    transition -- (state, action, reward, next_state, done, policy hidden state) -- and uniform
    sampling with replacement-free semantics left to the caller-supplied indices."""

    def __init__(self, capacity, state_dim, action_dim, gru_dim, device):
        f = dict(dtype=torch.float32, device=device)
        self.capacity, self.size, self.pos = capacity, 0, 0
        self._pos_dev = torch.zeros(1, dtype=torch.int64, device=device)   # write position, device copy (graph replays)
        self.state = torch.empty((capacity, state_dim), **f)
        self.next_state = torch.empty((capacity, state_dim), **f)
        self.action = torch.empty((capacity, action_dim), **f)
        self.reward = torch.empty((capacity, 1), **f)
        self.done = torch.empty((capacity, 1), **f)
        self.hidden = torch.empty((capacity, gru_dim), **f)

    def push_batch(self, state, action, reward, next_state, done, hidden):
        """state / next_state [n, state_dim], action [n, action_dim], hidden [n, gru] float32; reward [n] float64 and
        done [n] uint8 as the env returns them.  One kernel writes the n rows at the DEVICE-held position (a captured
        push lands in the right place on every replay), a second one advances it."""
        n = state.shape[0]
        if n > self.capacity:
            raise ValueError("batch larger than the replay capacity")
        ops.replay_push(self._ring(), self._pos_dev, state.contiguous(), action.contiguous(),
                        reward.reshape(n).to(torch.float64), next_state.contiguous(),
                        done.reshape(n).to(torch.uint8), hidden.contiguous())
        self.pos = (self.pos + n) % self.capacity      # host mirror (not advanced by graph replays)
        self.size = min(self.size + n, self.capacity)

    def _ring(self):
        return (self.state, self.action, self.reward, self.next_state, self.done, self.hidden)

    def __len__(self):
        return self.size

    def sample(self, batch_size, generator=None):
        """-> the tuple SAC_GRU_Agent.update_parameters(batch=...) takes (hidden as [1, B, gru])."""
        idx = torch.randint(0, self.size, (batch_size,), device=self.state.device, generator=generator)
        s, a, r, ns, d, h = ops.replay_gather(self._ring(), idx)
        return (s, a, r, ns, d, h.unsqueeze(0))


class SACRollout:
    """Config C4 of BASELINE.json: one SAC-GRU learner over E envs of S servers (the reference's SAC
    is single-agent: state = the flattened (S, 11) observation, one continuous weight per server,
    problem-04-sac-gru/src/trainer.py:96-133).  step(): actor forward + tanh-Gaussian sample for all
    envs -> env step -> transition into the device replay; update(): SAC_GRU_Agent.update_parameters
    on a device-sampled batch (its flat gradient buckets are all-reduced when a process group exists)."""

    def __init__(self, env: VecLoadBalanceEnv, agent, replay_capacity: int = 65536):
        if env.num_agents != 1 or env.action_type != "continuous":
            raise ValueError("SACRollout needs a single-agent env with continuous actions")
        self.env, self.agent = env, agent
        self.E, self.S = env.num_envs, env.total_servers
        if agent.state_dim != self.S * 11 or agent.action_dim != self.S:
            raise ValueError("agent dims must be (servers*11, servers)")
        self.hidden = agent.policy.init_hidden(self.E)
        self.replay = DeviceReplay(replay_capacity, agent.state_dim, agent.action_dim, agent.policy.gru_dim, env.device)
        self._state = torch.empty((self.E, self.S * 11), dtype=torch.float32, device=env.device)

    def reset(self):
        self.hidden = self.agent.policy.init_hidden(self.E)
        return self.env.reset()

    def step(self, eps=None, evaluate=False, store=True):
        """eps [E, S] standard-normal draws (None: drawn on the device).  Returns (obs, reward, done, action)."""
        self._state.copy_(self.env.obs.view(self.E, self.S * 11))      # obs is overwritten by the env step
        action, h_new = self.agent.select_action_batch(self._state, self.hidden, evaluate=evaluate, eps=eps)
        action = action.contiguous()
        obs, rew, done = self.env.step(action)
        if store:
            self.replay.push_batch(self._state, action, rew, obs.view(self.E, self.S * 11), done, self.hidden[0])
        self.hidden = h_new
        return obs, rew, done, action

    def update(self, updates=1, generator=None, sync_stats=True):
        if len(self.replay) < self.agent.batch_size:
            return None
        out = None
        for _ in range(updates):
            out = self.agent.update_parameters(1, batch=self.replay.sample(self.agent.batch_size, generator),
                                               sync_stats=sync_stats)
        return out

    # ---- CUDA graph: actor sampling + env step + replay push + one SAC update = one graph launch
    def _step_overlapped(self, updates):
        """One step with the update beside the env step (capture(overlap=True)).  What depends on what:
        the actor reads the policy parameters the update will overwrite -> the update starts after the actor forward;
        the update's batch is gathered from the ring BEFORE this step's transitions are pushed (the push waits for
        the gather), i.e. it samples the transitions of steps < t where the sequential order samples steps <= t;
        the next step's actor uses the parameters this update leaves behind, exactly as in the sequential order."""
        dev = self.env.device
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_upd_stream", None) is None:
            self._upd_stream = torch.cuda.Stream(device=dev)
        upd = self._upd_stream
        self._state.copy_(self.env.obs.view(self.E, self.S * 11))      # obs is overwritten by the env step
        action, h_new = self.agent.select_action_batch(self._state, self.hidden)
        action = action.contiguous()
        upd.wait_stream(cur)
        with torch.cuda.stream(upd):
            batch = self.replay.sample(self.agent.batch_size)
            gathered = torch.cuda.Event()
            gathered.record(upd)
            losses = None
            for _ in range(updates):
                losses = self.agent.update_parameters(1, batch=batch, sync_stats=False)
                if updates > 1:
                    batch = self.replay.sample(self.agent.batch_size)
        obs, rew, done = self.env.step(action)
        cur.wait_event(gathered)
        self.replay.push_batch(self._state, action, rew, obs.view(self.E, self.S * 11), done, self.hidden[0])
        self.hidden = h_new
        cur.wait_stream(upd)
        return (obs, rew, done, action), losses

    def capture(self, updates=1, warmup=2, overlap=False):
        """Capture `step(); update(updates)` into a CUDA graph.  overlap=True: the update runs BESIDE the env step as a
        parallel branch of the graph (see _step_overlapped: same work per step, the batch is drawn from the ring as it
        stood before this step's push).  Under data parallelism the NCCL all-reduces of
        the four gradient buckets are captured too (side-stream fork / join, policy/nn.py:Adam.reduce_async): every
        rank replays its graph once per step, in lock-step, so the collectives match up.  Needs a FULL replay
        ring (the sampling range is fixed at capture) -- run at least capacity / num_envs eager steps first (they
        also create the NCCL communicator, which cannot happen during capture).  Gaussian and index draws come
        from torch's default CUDA generator, which is graph-safe."""
        if len(self.replay) < self.replay.capacity:
            raise RuntimeError("fill the replay ring before capturing (sampling range is fixed in the graph)")
        dev = self.env.device
        self._h_static = self.hidden.clone()

        def body():
            self.hidden = self._h_static
            if overlap:
                out, losses = self._step_overlapped(updates)
                self._h_static.copy_(self.hidden)
                self.hidden = self._h_static
                return out, losses
            out = self.step()
            self._h_static.copy_(self.hidden)
            self.hidden = self._h_static
            losses = self.update(updates, sync_stats=False)
            return out, losses

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from .policy import ops as _ops
        n0 = self.env.launch_count + _ops.LAUNCHES
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_out = body()
        self.graph_launches = self.env.launch_count + _ops.LAUNCHES - n0
        self._graph_updates = updates
        return self

    def step_graph(self):
        """Replay the captured rollout step + update.  Returns ((obs, reward, done, action), losses)."""
        self._graph.replay()
        for opt in (self.agent.q1_optimizer, self.agent.q2_optimizer, self.agent.policy_optimizer):
            opt.t += self._graph_updates            # host mirrors of the device-side step counters
        if self.agent.auto_entropy_tuning:
            self.agent.alpha_optimizer.t += self._graph_updates
        self.agent.total_steps += self._graph_updates
        return self._graph_out
