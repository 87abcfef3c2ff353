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

    def step(self, epsilon: float = 0.0, u=None, rnd=None):
        """One rollout step for every env.  u [E, A] uniforms / rnd [E, A] int32 random actions
        pre-drawn by the caller (None = greedy).  Returns (obs, reward, done, actions [E, A])."""
        obs = self.env.obs.view(self.E, self.A, self.Sa * 11)
        act, self.hidden, _ = self.agent.select_actions_batch(obs, self.hidden, epsilon, u, rnd)
        ops.onehot_action(act, self.Sa, self.hot, self.cold, out=self._env_action)
        o, r, d = self.env.step(self._env_action)
        return o, r, d, act
