"""Drop-in `LoadBalanceEnv` with the reference's Python API.

Mirror of simulation-mode/problem-03-rl-environment/src/env.py:41-481: same
constructor keywords (env.py:71-87), attributes, `reset()`, `step(action)` ->
`(obs (S,11) float32, reward float, done bool, info dict)`, `seed`, `render`,
`close` and the private helpers the reference's tests call
(`_action_to_weights`, `_dict_to_array`, `_array_to_dict`,
`_simulate_observation`, `_normalize_observation`).

Two modes, both computed on the GPU through the C ABI:

* ``mode="legacy"`` -- what the reference's simulation mode actually does: the
  observation is a draw from ``np.random.RandomState(seed)`` (env.py:425-448)
  and only the reward is real.  The MT19937 stream is replayed on the device
  (`mlb_legacy_obs`), bit-for-bit equal to the reference for the same seed.
* ``mode="flow"`` -- the flow-level simulation of SURVEY App. B (arrivals ->
  weighted assignment -> FIFO queues -> reservoir features -> fairness reward)
  through the fused step kernel; `step_interval` is the simulated window and
  is never slept.

``mode=None`` (default) picks "flow" when arrivals are configured (`trace`,
`arrivals` or `arrival_rate`) and "legacy" otherwise, so a reference call site
that passes only reference keywords gets reference results.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .rewards import RewardFunction
from .spaces import Box, MultiDiscrete
from .vec_env import VecLoadBalanceEnv

FEATURE_NAMES = _lib.FEATURE_NAMES


class LoadBalanceEnv:
    """Gym-style load-balancing env (env.py:41-66) backed by CUDA kernels."""

    DEFAULT_DISCRETE_WEIGHTS = [1.0, 1.5, 2.0]            # env.py:69

    def __init__(
        self,
        num_servers: int = 4,
        action_type: str = 'discrete',
        discrete_weights: Optional[List[float]] = None,
        max_weight: float = 10.0,
        min_weight: float = 0.1,
        reward_metric: str = 'jain',
        reward_field: str = 'flow_duration_avg_decay',
        step_interval: float = 0.25,
        max_steps: int = 10000,
        use_shm: bool = False,
        shm_name: Optional[str] = None,
        use_ground_truth: bool = False,
        normalize_obs: bool = False,
        seed: Optional[int] = None,
        # ---- extensions (keyword-only in spirit; all default to reference behaviour)
        mode: Optional[str] = None,
        trace: Optional[str] = None,
        arrivals: Optional[dict] = None,
        arrival_rate: Optional[float] = None,
        mean_work: float = 1.0,
        horizon: Optional[float] = None,
        server_speeds=None,
        policy: str = 'sed',
        reservoir_capacity: int = 128,
        queue_capacity: int = 160,
        realtime: bool = False,
        device: int = 0,
        num_lb_agents: int = 1,
    ):
        self.num_servers = num_servers
        self.action_type = action_type
        self.discrete_weights = discrete_weights or self.DEFAULT_DISCRETE_WEIGHTS
        self.max_weight = max_weight
        self.min_weight = min_weight
        self.step_interval = step_interval
        self.max_steps = max_steps
        self.use_shm = use_shm
        self.shm_name = shm_name
        self.use_ground_truth = use_ground_truth
        self.normalize_obs = normalize_obs
        self.realtime = realtime
        self._device = device

        self.reward_fn = RewardFunction(metric=reward_metric, reward_field=reward_field)   # env.py:124
        self._seed = seed
        self._setup_spaces()                                                                 # env.py:130

        self.shm = None
        if self.use_shm:
            if self.shm_name is None:
                raise ValueError("shm_name required when use_shm=True")                      # env.py:136
            # The VPP shared-memory transport (problem-02) is out of scope here; like the
            # reference when attach fails (env.py:140-143) fall back to simulation mode.
            print("Warning: shared-memory transport is not part of marllb_b200")
            print("Falling back to simulation mode")
            self.use_shm = False

        if mode is None:
            mode = 'flow' if (trace is not None or arrivals is not None or arrival_rate is not None) else 'legacy'
        if mode not in ('legacy', 'flow'):
            raise ValueError(f"Unknown mode: {mode}")
        self.mode = mode

        self.current_step = 0
        self.last_observation = None
        self.episode_rewards = []
        self.episode_return = 0.0
        self.obs_mean = np.zeros((num_servers, 11))
        self.obs_std = np.ones((num_servers, 11))
        self.obs_count = 0

        self._L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("marllb_b200 needs a CUDA device (there is no CPU fallback)")
        self._dev = torch.device('cuda', device)
        if mode == 'legacy':
            self._mt = torch.empty(625, dtype=torch.int32, device=self._dev)
            self._obs_dev = torch.empty((num_servers, 11), dtype=torch.float32, device=self._dev)
            self._reseed(seed)
            self._vec = None
        else:
            if reward_field not in FEATURE_NAMES:
                raise ValueError(f"Unknown reward_field: {reward_field}")
            # num_lb_agents > 1: the A LB agents of MultiAgentLoadBalanceEnv (multi_agent_env.py:57-76) each assign
            # their own arrival stream to their own block of num_servers / A servers
            if num_servers % num_lb_agents:
                raise ValueError("num_servers must be a multiple of num_lb_agents")
            A = num_lb_agents
            self._vec = VecLoadBalanceEnv(
                1, num_servers=num_servers // A, num_agents=A, action_type=action_type,
                discrete_weights=self.discrete_weights, max_weight=max_weight, min_weight=min_weight,
                reward_metric=reward_metric, reward_field=reward_field, step_interval=step_interval,
                max_steps=max_steps, policy=policy, reservoir_capacity=reservoir_capacity,
                queue_capacity=queue_capacity, record_assign=True, device=device)
            if server_speeds is not None:
                self._vec.set_speeds(server_speeds)
            hz = horizon if horizon is not None else float(step_interval) * max_steps
            if arrivals is not None:
                streams = [arrivals] if isinstance(arrivals, dict) else list(arrivals)
                if len(streams) != A:
                    raise ValueError(f"arrivals: expected {A} streams (one per LB agent), got {len(streams)}")
                self._vec.load_arrivals(streams)
            elif trace is not None:
                from .traces import load_trace, split_round_robin
                tr = load_trace(trace, horizon=hz)
                self._vec.load_arrivals([tr] if A == 1 else split_round_robin(tr, A))   # row r -> agent r mod A
            else:
                rate = arrival_rate if arrival_rate is not None else 8.0 * num_servers / A   # per LB agent
                self._vec.gen_poisson(rate, mean_work, hz, seed=0 if seed is None else int(seed))

    # ------------------------------------------------------------------ spaces
    def _setup_spaces(self):
        num_features = 11
        if self.use_ground_truth:
            num_features += 3                                                               # env.py:160-161
        self.observation_space = Box(low=0, high=np.inf, shape=(self.num_servers, num_features),
                                     dtype=np.float32)
        if self.action_type == 'discrete':
            self.action_space = MultiDiscrete([len(self.discrete_weights)] * self.num_servers)
        elif self.action_type == 'continuous':
            self.action_space = Box(low=self.min_weight, high=self.max_weight,
                                    shape=(self.num_servers,), dtype=np.float32)
        else:
            raise ValueError(f"Unknown action_type: {self.action_type}")                     # env.py:184

    # ------------------------------------------------------------------ legacy RNG
    def _reseed(self, seed):
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        s = torch.tensor(np.array([int(seed) & 0xffffffff], np.uint32).view(np.int32)).to(self._dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._L.mlb_legacy_seed(C.c_void_p(self._mt.data_ptr()), C.c_void_p(s.data_ptr()), 1, st))
        torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ API
    def reset(self) -> np.ndarray:
        """env.py:186-213."""
        self.current_step = 0
        self.episode_rewards = []
        self.episode_return = 0.0
        if self.mode == 'legacy':
            obs = self._simulate_observation()
        else:
            obs = self._vec.reset()[0].cpu().numpy().copy()
        if self.normalize_obs:
            obs = self._normalize_observation(obs)
        return obs

    def step(self, action: np.ndarray) -> Tuple[np.ndarray, float, bool, Dict[str, Any]]:
        """env.py:215-286."""
        self.current_step += 1
        weights = self._action_to_weights(action)
        if self.mode == 'legacy':
            if self.realtime:
                time.sleep(self.step_interval)                                              # env.py:257
            next_obs = self._simulate_observation()
            obs_dict = self._array_to_dict(next_obs)
            reward = self.reward_fn.compute(obs_dict)                                       # env.py:262
        else:
            a = np.asarray(action)
            a = a.astype(np.float32) if self.action_type == 'continuous' else a.astype(np.int32)
            obs_t, rew_t, _ = self._vec.step(a.reshape(1, -1))
            next_obs = obs_t[0].cpu().numpy().copy()
            reward = float(rew_t[0].item())
            self._vec.check_status()
            obs_dict = self._array_to_dict(next_obs)
        self.episode_rewards.append(reward)
        self.episode_return += reward
        done = self.current_step >= self.max_steps                                          # env.py:267
        info = {
            'step': self.current_step,
            'weights': weights.tolist(),
            'active_servers': obs_dict.get('active_servers', list(range(self.num_servers))),
            'episode_return': self.episode_return,
        }
        if done:
            info['episode'] = {'r': self.episode_return, 'l': self.current_step}
        if self.normalize_obs:
            next_obs = self._normalize_observation(next_obs)
        return next_obs, reward, done, info

    def render(self, mode: str = 'human'):
        if mode == 'human':
            print(f"\n{'=' * 60}")
            print(f"Step: {self.current_step}/{self.max_steps}")
            print(f"Episode Return: {self.episode_return:.4f}")
            print("=" * 60)

    def close(self):
        if self._vec is not None:
            self._vec.close()

    def seed(self, seed: Optional[int] = None):
        """env.py:327-330."""
        self._seed = seed
        if self.mode == 'legacy':
            self._reseed(seed)
        return [seed]

    # ------------------------------------------------------------------ helpers
    def _action_to_weights(self, action: np.ndarray) -> np.ndarray:
        """env.py:334-353."""
        if self.action_type == 'discrete':
            weights = np.array([self.discrete_weights[int(a)] for a in action], dtype=np.float32)
        else:
            weights = np.asarray(action, dtype=np.float32)
            weights = np.clip(weights, self.min_weight, self.max_weight)
        return weights

    def _dict_to_array(self, obs_dict: dict) -> np.ndarray:
        """env.py:355-389."""
        obs = np.zeros((self.num_servers, 11), dtype=np.float32)
        active_servers = obs_dict.get('active_servers', [])
        server_stats = obs_dict.get('server_stats', {})
        for sid in active_servers:
            if sid < self.num_servers and sid in server_stats:
                stats = server_stats[sid]
                for i, fname in enumerate(FEATURE_NAMES):
                    obs[sid, i] = stats.get(fname, 0.0)
        return obs

    def _array_to_dict(self, obs: np.ndarray) -> dict:
        """env.py:391-423."""
        server_stats = {}
        active_servers = []
        for sid in range(self.num_servers):
            if np.any(obs[sid] > 0):
                active_servers.append(sid)
                server_stats[sid] = {fname: float(obs[sid, i]) for i, fname in enumerate(FEATURE_NAMES)}
        return {'active_servers': active_servers, 'server_stats': server_stats,
                'sequence_id': self.current_step}

    def _simulate_observation(self) -> np.ndarray:
        """env.py:425-448, replayed on the device (legacy mode only)."""
        if self.mode != 'legacy':
            raise RuntimeError("_simulate_observation is the legacy-mode observation source")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._L.mlb_legacy_obs(C.c_void_p(self._mt.data_ptr()), 1, self.num_servers,
                                     C.c_void_p(self._obs_dev.data_ptr()), st))
        return self._obs_dev.cpu().numpy().copy()

    def _normalize_observation(self, obs: np.ndarray) -> np.ndarray:
        """env.py:450-470 (host bookkeeping, identical arithmetic)."""
        self.obs_count += 1
        delta = obs - self.obs_mean
        self.obs_mean += delta / self.obs_count
        delta2 = obs - self.obs_mean
        self.obs_std = np.sqrt(np.maximum(
            (self.obs_std ** 2 * (self.obs_count - 1) + delta * delta2) / self.obs_count, 1e-8))
        return (obs - self.obs_mean) / (self.obs_std + 1e-8)

    # flow-mode extras ----------------------------------------------------------
    def assignments(self) -> np.ndarray:
        """Server id chosen for every flow processed so far (flow mode)."""
        n = int(self._vec.get_state("arr_cursor")[0, 0])
        return self._vec.get_assignments(n)


class LoadBalanceEnvGym(LoadBalanceEnv):
    """Alias kept for call sites that used the gym.Env subclass (env.py:474-481)."""
    metadata = {'render.modes': ['human']}

    def __init__(self, **kwargs):
        LoadBalanceEnv.__init__(self, **kwargs)
