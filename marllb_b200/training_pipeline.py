"""Offline training driver on the batched env.

Mirrors `TrainingPipeline` of the reference
(simulation-mode/problem-06-vpp-integration/src/training_pipeline.py:35-421): same constructor
arguments, CLI flags, epsilon schedule `max(0.01, 0.1 - episode/5000)` (:320), 200-step episode
cap (:322), evaluation every `eval_interval` episodes with the best model saved as
`<agent>_best.pth` (:263-272), checkpoints `<agent>_ep<N>.pth` and `training_stats.json` with the
reference's keys every `save_interval` (:275-292).

What differs, on purpose:
  * traces are actually used: `data/trace/**/*.csv` is parsed as the TSV it is and replayed
    through the flow-level env (the reference reads it with the wrong separator and then ignores
    it); without trace files synthetic Poisson arrivals are generated like
    `_generate_synthetic_traces` (:141-155);
  * `num_envs` episodes run at once on the GPU (VecLoadBalanceEnv + batched action selection);
    every env contributes one episode per round, so `num_episodes` counts the same thing;
  * it calls agent methods that exist (`store_episode`, `update`, `replay_buffer.push`,
    `update_parameters`), SURVEY App. C #4.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import time

import numpy as np
import torch

from . import traces as _traces
from .policy import QMIXAgent, SAC_GRU_Agent, ops
from .vec_env import VecLoadBalanceEnv

MAX_EPISODE_STEPS = 200   # training_pipeline.py:322
DT = 0.25                 # env.py:80


def epsilon_schedule(episode_num: int) -> float:
    """training_pipeline.py:320: epsilon = max(0.01, 0.1 - episode / 5000)."""
    return max(0.01, 0.1 - episode_num / 5000.0)


class TrainingPipeline:
    def __init__(self, agent_type='qmix', num_servers=16, num_agents=4, trace_dir='data/trace',
                 checkpoint_dir='checkpoints', config=None, num_envs=32, device=None, verbose=True):
        if agent_type not in ('qmix', 'sac-gru'):
            raise ValueError(f"Unknown agent type: {agent_type}")
        self.agent_type, self.num_servers = agent_type, num_servers
        self.num_agents = num_agents if agent_type == 'qmix' else 1
        if num_servers % self.num_agents:
            raise ValueError("num_servers must be a multiple of num_agents")
        self.trace_dir, self.checkpoint_dir = trace_dir, checkpoint_dir
        self.config = config or {}
        self.num_envs, self.verbose = num_envs, verbose
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        os.makedirs(checkpoint_dir, exist_ok=True)
        self._rng = np.random.RandomState(self.config.get('seed', 0))
        self.traces = self._load_traces()
        self.env = self._init_env()
        self.agent = self._init_agent()
        self.episode_rewards, self.episode_lengths, self.losses = [], [], []
        self._say(f"[TrainingPipeline] {agent_type}: {num_servers} servers, {self.num_agents} agents, "
                  f"{len(self.traces)} traces, {num_envs} envs in parallel")

    def _say(self, msg):
        if self.verbose:
            print(msg)

    # ------------------------------------------------------------------ traces (:98-155)
    def _load_traces(self):
        horizon = MAX_EPISODE_STEPS * DT
        out = []
        for f in sorted(glob.glob(os.path.join(self.trace_dir, '**', '*.csv'), recursive=True)):
            try:
                tr = _traces.load_trace(f, horizon=horizon, work_scale=self.config.get('work_scale', 1e-6))
            except Exception as e:  # like the reference: report and go on
                self._say(f"  Warning: Failed to load {f}: {e}")
                continue
            if len(tr['time']):
                out.append(tr)
        if not out:
            self._say("  Warning: No traces found, using synthetic Poisson")
            speeds = self.config.get('server_speeds')
            speeds_sum = 1.5 * self.num_servers if speeds is None else float(np.sum(speeds))
            for rate in self.config.get('rates', [100, 200, 500]):       # :143-145
                mean_work = 0.8 * speeds_sum / rate
                out.append(_traces.poisson_trace(rate, horizon, mean_work, rng=self._rng))
        return out

    # ------------------------------------------------------------------ env / agent (:156-199)
    def _init_env(self):
        Sa = self.num_servers // self.num_agents
        env = VecLoadBalanceEnv(self.num_envs, num_servers=Sa, num_agents=self.num_agents,
                                action_type='discrete' if self.agent_type == 'qmix' else 'continuous',
                                action_dtype='uint8', max_steps=MAX_EPISODE_STEPS, device=self.device.index or 0,
                                reward_metric=self.config.get('reward_metric', 'jain'))
        speeds = self.config.get('server_speeds')
        if speeds is None:
            speeds = np.where(np.arange(self.num_servers) % 2 == 0, 1.0, 2.0)
        self.server_speeds = np.asarray(speeds, np.float32).reshape(self.num_servers)
        env.set_speeds(self.server_speeds)
        return env

    def _init_agent(self):
        S, A = self.num_servers, self.num_agents
        if self.agent_type == 'qmix':
            Sa = S // A
            agent = QMIXAgent(num_agents=A, state_dim=4 * S + 10, obs_dim=Sa * 11, action_dim=Sa, hidden_dim=64,
                              mixing_embed_dim=32, lr=self.config.get('learning_rate', 0.0005),
                              gamma=self.config.get('gamma', 0.99), batch_size=self.config.get('batch_size', 32),
                              max_seq_len=self.config.get('max_seq_len', 50), device=self.device)
            agent.graph_updates = bool(self.config.get('graph_updates', True))   # update() replayed as one CUDA graph
            return agent
        lr = self.config.get('learning_rate', 0.0003)
        return SAC_GRU_Agent(state_dim=S * 11, action_dim=S, hidden_dim=256, lr_policy=lr, lr_q=lr, lr_alpha=lr,
                             gamma=self.config.get('gamma', 0.99), batch_size=self.config.get('batch_size', 256),
                             device=self.device)

    # ------------------------------------------------------------------ one round of num_envs episodes
    def _load_round(self):
        """Every env replays a randomly chosen trace (:215-217); A agents share it round-robin."""
        streams = []
        for _ in range(self.num_envs):
            tr = self.traces[self._rng.randint(0, len(self.traces))]
            streams += _traces.split_round_robin(tr, self.num_agents)
        self.env.load_arrivals(streams)
        return self.env.reset()

    def _global_state(self, obs, t):
        """[E, 4*S + 10]: the first four columns of every server + ten global metrics, the shape the
        reference declares (multi_agent_env.py:86-98, 241-282)."""
        E, S = obs.shape[0], obs.shape[1]
        load = obs[:, :, 0]
        g = torch.stack([load.sum(1), obs[:, :, 1].mean(1), obs[:, :, 6].mean(1), self.env.reward.float(),
                         (load > 0).float().mean(1), load.std(1, unbiased=False), load.max(1).values,
                         load.min(1).values, torch.full((E,), t / MAX_EPISODE_STEPS, device=obs.device),
                         torch.full((E,), float(self.num_agents), device=obs.device)], 1)
        return torch.cat([obs[:, :, :4].reshape(E, 4 * S), g], 1)

    def _run_round(self, episode_num, explore=True, learn=True):
        E, A, S = self.num_envs, self.num_agents, self.num_servers
        Sa = S // A
        obs = self._load_round()
        ret = torch.zeros(E, dtype=torch.float64, device=self.device)
        epsilon = epsilon_schedule(episode_num) if explore else 0.0
        loss = None
        if self.agent_type == 'qmix':
            hid = None
            T = MAX_EPISODE_STEPS
            ep_obs = torch.empty((T, E, A, Sa * 11), device=self.device)
            ep_act = torch.empty((T, E, A), dtype=torch.int32, device=self.device)
            ep_rew = torch.empty((T, E), dtype=torch.float64, device=self.device)
            ep_state = torch.empty((T, E, 4 * S + 10), device=self.device)
            env_action = torch.empty((E, S), dtype=torch.uint8, device=self.device)
            for t in range(T):
                o = obs.view(E, A, Sa * 11)
                u = torch.as_tensor(self._rng.random_sample((E, A)).astype(np.float32)).to(self.device)
                rnd = torch.as_tensor(self._rng.randint(0, Sa, (E, A)).astype(np.int32)).to(self.device)
                ep_obs[t].copy_(o)
                ep_state[t].copy_(self._global_state(obs, t))
                act, hid, _ = self.agent.select_actions_batch(o, hid, epsilon, u, rnd)
                ops.onehot_action(act, Sa, 2, 0, out=env_action)
                obs, rew, done = self.env.step(env_action)
                ep_act[t].copy_(act)
                ep_rew[t].copy_(rew)
                ret += rew * A                                                    # sum(rewards), :341
            if learn:
                h_obs, h_act = ep_obs.cpu().numpy(), ep_act.cpu().numpy()
                h_rew, h_state = ep_rew.cpu().numpy(), ep_state.cpu().numpy()
                for e in range(E):
                    self.agent.store_episode({
                        'observations': [list(h_obs[t, e]) for t in range(T)],
                        'actions': [list(h_act[t, e]) for t in range(T)],
                        'rewards': [[float(h_rew[t, e])] * A for t in range(T)],
                        'states': [h_state[t, e] for t in range(T)],
                        'dones': [t == T - 1 for t in range(T)]})
                if self.agent.episode_buffer.is_ready(self.agent.batch_size):
                    for _ in range(self.config.get('updates_per_round', 1)):
                        out = self.agent.update()
                    loss = None if out is None else out['loss']
        else:
            hid = self.agent.policy.init_hidden(E)                                # [1, E, gru] zeros (:311)
            for t in range(MAX_EPISODE_STEPS):
                s = obs.reshape(E, S * 11).clone()
                action, hid_new = self.agent.select_action_batch(s, hid, evaluate=not explore)
                obs, rew, done = self.env.step(action.contiguous())
                if learn:
                    hs, ha = s.cpu().numpy(), action.cpu().numpy()
                    hn, hr = obs.reshape(E, S * 11).cpu().numpy(), rew.cpu().numpy()
                    hh = hid[0].cpu().numpy()
                    for e in range(E):
                        self.agent.replay_buffer.push(hs[e], ha[e], float(hr[e]), hn[e], t == MAX_EPISODE_STEPS - 1,
                                                      hh[e][None, None])
                hid = hid_new
                ret += rew
            if learn and self.agent.replay_buffer.is_ready(self.agent.batch_size):
                out = self.agent.update_parameters(updates=self.config.get('updates_per_round', 1))
                loss = None if out is None else out['q1']
        self.env.check_status()
        return ret.cpu().numpy(), loss

    # ------------------------------------------------------------------ train / evaluate (:201-296, :371-421)
    def train(self, num_episodes=10000, save_interval=100, eval_interval=100):
        start, best_reward, done_eps = time.time(), -np.inf, 0
        checkpoint_path = None
        next_eval, next_save = eval_interval, save_interval
        while done_eps < num_episodes:
            t0 = time.time()
            rets, loss = self._run_round(done_eps)
            for r in rets:
                self.episode_rewards.append(float(r))
                self.episode_lengths.append(MAX_EPISODE_STEPS)
            if loss is not None:
                self.losses.append(float(loss))
            done_eps += len(rets)
            avg_loss = float(np.mean(self.losses[-10:])) if self.losses else 0.0
            self._say(f"[{done_eps:05d}/{num_episodes}] Reward avg: {np.mean(rets):7.2f} Loss: {avg_loss:7.4f} "
                      f"Time: {time.time() - t0:5.2f}s")
            if done_eps >= next_eval:
                next_eval += eval_interval
                eval_reward = self._evaluate()
                self._say(f"  [Eval] Average reward: {eval_reward:.2f}")
                if eval_reward > best_reward:
                    best_reward = eval_reward
                    self.agent.save(os.path.join(self.checkpoint_dir, f'{self.agent_type}_best.pth'))
            if done_eps >= next_save or done_eps >= num_episodes:
                next_save += save_interval
                checkpoint_path = os.path.join(self.checkpoint_dir, f'{self.agent_type}_ep{done_eps}.pth')
                self.agent.save(checkpoint_path)
                with open(os.path.join(self.checkpoint_dir, 'training_stats.json'), 'w') as f:
                    json.dump({'episode_rewards': self.episode_rewards, 'episode_lengths': self.episode_lengths,
                               'losses': self.losses, 'best_reward': float(best_reward),
                               'total_episodes': done_eps, 'total_time': time.time() - start}, f, indent=2)
        return {'best_reward': float(best_reward), 'final_checkpoint': checkpoint_path,
                'total_episodes': done_eps, 'total_time': time.time() - start}

    def _evaluate(self, num_episodes=10):
        rets, _ = self._run_round(0, explore=False, learn=False)
        return float(np.mean(rets[:max(1, min(num_episodes, len(rets)))]))


def main(argv=None):
    p = argparse.ArgumentParser(description='Training Pipeline for VPP Load Balancer (batched, B200)')
    p.add_argument('--agent', type=str, default='qmix', choices=['sac-gru', 'qmix'])
    p.add_argument('--servers', type=int, default=16)
    p.add_argument('--agents', type=int, default=4)
    p.add_argument('--episodes', type=int, default=10000)
    p.add_argument('--trace-dir', type=str, default='data/trace')
    p.add_argument('--checkpoint-dir', type=str, default='checkpoints')
    p.add_argument('--save-interval', type=int, default=100)
    p.add_argument('--eval-interval', type=int, default=100)
    p.add_argument('--config', type=str, default=None)
    p.add_argument('--num-envs', type=int, default=32, help='episodes simulated in parallel on the GPU')
    a = p.parse_args(argv)
    config = {}
    if a.config and os.path.exists(a.config):
        with open(a.config) as f:
            config = json.load(f)
    TrainingPipeline(a.agent, a.servers, a.agents, a.trace_dir, a.checkpoint_dir, config, a.num_envs).train(
        a.episodes, a.save_interval, a.eval_interval)


if __name__ == '__main__':
    main()
