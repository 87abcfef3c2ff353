"""Drop-in `MultiAgentLoadBalanceEnv` (problem-05-qmix/src/multi_agent_env.py:22-290).

One env of A agents x Sa servers; `reset()` -> list of A observations,
`step(actions)` -> (list obs, list rewards, done, info), `get_state()`.

`strict_reference=True` (default) reproduces the reference's observation
slicing bit-for-bit, including its dimension quirk: the
flattened (S_tot, 11) array is cut with stride 4, and everything past index
4*S_tot is appended, so the per-agent vector has 4*Sa + 7*S_tot entries while
`obs_dim` advertises 4*Sa + 4; `get_state()` is zeros(4*S_tot) + 10 globals.
`strict_reference=False` gives the clean layout: agent obs = its own servers'
(Sa, 11) rows flattened, state = the full (S_tot*11) observation + the 10 globals.
"""
from __future__ import annotations

import numpy as np

from .env import LoadBalanceEnv


class MultiAgentLoadBalanceEnv:
    def __init__(self, num_agents: int = 4, servers_per_agent: int = 4,
                 action_type: str = 'continuous', reward_metric: str = 'jain',
                 max_steps: int = 100, use_shm: bool = False, global_reward: bool = True,
                 strict_reference: bool = True, **env_kwargs):
        self.num_agents = num_agents
        self.servers_per_agent = servers_per_agent
        self.total_servers = num_agents * servers_per_agent                  # multi_agent_env.py:59
        self.global_reward = global_reward
        self.strict_reference = strict_reference
        # flow-level simulation (mode='flow' or trace= / arrivals= / arrival_rate=): the wrapped env runs one
        # VecLoadBalanceEnv(1, num_agents=A) -- every LB agent assigns its own arrival stream over its own servers --
        # like LoadBalanceEnv does for one agent (env.py:132-137 here); `arrivals` is then a list of A streams
        if env_kwargs.get('mode') == 'flow' or any(k in env_kwargs for k in ('trace', 'arrivals', 'arrival_rate')):
            env_kwargs = dict(env_kwargs, num_lb_agents=num_agents)
        self.env = LoadBalanceEnv(num_servers=self.total_servers, action_type=action_type,
                                  reward_metric=reward_metric, max_steps=max_steps,
                                  use_shm=use_shm, **env_kwargs)             # multi_agent_env.py:63-69
        self.agent_servers = {i: list(range(i * servers_per_agent, (i + 1) * servers_per_agent))
                              for i in range(num_agents)}                    # :72-76
        self.observation_space = self.env.observation_space
        self.action_space = self.env.action_space
        self.obs_dim = self._get_obs_dim()
        self.state_dim = self._get_state_dim()
        self._last_global_obs = None

    def _get_obs_dim(self):
        if self.strict_reference:
            return self.servers_per_agent * 4 + 4                            # :86-93 (advertised)
        return self.servers_per_agent * 11

    def _get_state_dim(self):
        if self.strict_reference:
            return self.total_servers * 4 + 10                               # :95-98
        return self.total_servers * 11 + 10

    def reset(self):
        global_obs = self.env.reset()
        self._last_global_obs = global_obs
        return [self._get_agent_observation(global_obs, i) for i in range(self.num_agents)]

    def step(self, actions):
        global_action = self._combine_actions(actions)
        global_obs, global_reward, done, info = self.env.step(global_action)
        self._last_global_obs = global_obs
        observations = [self._get_agent_observation(global_obs, i) for i in range(self.num_agents)]
        if self.global_reward:
            rewards = [global_reward] * self.num_agents                      # :143-145
        else:
            rewards = self._compute_local_rewards(info)
        return observations, rewards, done, info

    def _get_agent_observation(self, global_obs, agent_id):
        if not self.strict_reference:
            lo = agent_id * self.servers_per_agent
            return np.asarray(global_obs[lo:lo + self.servers_per_agent]).reshape(-1)
        flat = np.asarray(global_obs).flatten()                              # :164-165
        per = 4                                                              # :170
        server_obs_dim = self.total_servers * per
        own = []
        for server_idx in self.agent_servers[agent_id]:                      # :177-180
            own.extend(flat[server_idx * per:(server_idx + 1) * per].tolist())
        return np.concatenate([np.array(own).flatten(), flat[server_obs_dim:].flatten()])   # :183-186

    def _combine_actions(self, actions):
        global_action = np.zeros(self.total_servers)
        for agent_id, action in enumerate(actions):
            action = np.atleast_1d(action)      # integer actions (QMIX) made the reference raise; accept both
            for i, server_idx in enumerate(self.agent_servers[agent_id]):
                if i < len(action):
                    global_action[server_idx] = action[i]                    # :202-206
        return global_action

    def _compute_local_rewards(self, info):
        """multi_agent_env.py:210-239; needs info['server_loads'], which the reference env never
        provides either (KeyError there as well)."""
        rewards = []
        for agent_id in range(self.num_agents):
            loads = [info['server_loads'][idx] for idx in self.agent_servers[agent_id]]
            if sum(loads) == 0:
                rewards.append(0.0)
            else:
                s, s2 = sum(loads), sum(x ** 2 for x in loads)
                rewards.append((s ** 2) / (self.servers_per_agent * s2 + 1e-8))
        return rewards

    def get_state(self):
        if self.strict_reference:
            # env.last_observation is never set in simulation mode -> zeros (multi_agent_env.py:249-254)
            obs = np.zeros(self.total_servers * 4)
        else:
            obs = (np.asarray(self._last_global_obs, dtype=np.float64).reshape(-1)
                   if self._last_global_obs is not None else np.zeros(self.total_servers * 11))
        loads = [0] * self.total_servers
        global_metrics = [0, 0, 0, 0, 0, np.std(loads), np.max(loads), np.min(loads),
                          self.env.current_step / self.env.max_steps, self.num_agents]   # :267-278
        return np.concatenate([obs, global_metrics])

    def render(self, mode='human'):
        return self.env.render(mode)

    def close(self):
        self.env.close()
