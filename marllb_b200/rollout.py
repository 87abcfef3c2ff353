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
