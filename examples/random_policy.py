"""Random-policy baseline on the drop-in env -- the reference example
simulation-mode/problem-03-rl-environment/examples/random_policy.py with one import changed, followed by the same
loop on the flow-level simulation (a trace replayed through SED assignment, queues and reservoirs) and on 4096
batched envs.  Needs a CUDA device.

    python examples/random_policy.py [path/to/trace.csv]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout
from marllb_b200 import LoadBalanceEnv, VecLoadBalanceEnv       # reference: from env import LoadBalanceEnv


def run(env, episodes=2, max_steps=20):
    returns = []
    for ep in range(episodes):
        env.reset()
        total = 0.0
        for step in range(max_steps):
            action = env.action_space.sample()
            obs, reward, done, info = env.step(action)
            total += reward
            if step % 5 == 0:
                print(f"  ep {ep + 1} step {step + 1:3d}: weights={[f'{w:.1f}' for w in info['weights']]} reward={reward:.4f} "
                      f"n_flow_on={obs[:, 0].astype(int).tolist()}")
            if done:
                break
        returns.append(total)
    print(f"  mean return {np.mean(returns):.4f} +- {np.std(returns):.4f}")


if __name__ == "__main__":
    print("== legacy mode: the reference's own observation stream, bit for bit (seed 42)")
    run(LoadBalanceEnv(num_servers=4, action_type="discrete", reward_metric="jain",
                       reward_field="flow_duration_avg_decay", max_steps=20, use_shm=False, seed=42))
    print("== flow mode: Poisson flows at 40/s over servers of speed 1, 1, 2, 2 (rho = 0.8)")
    run(LoadBalanceEnv(num_servers=4, action_type="discrete", max_steps=20, arrival_rate=40.0, mean_work=0.12,
                       server_speeds=[1, 1, 2, 2], policy="sed", seed=42))
    if len(sys.argv) > 1:
        print(f"== flow mode: trace {sys.argv[1]}")
        run(LoadBalanceEnv(num_servers=4, action_type="discrete", max_steps=20, trace=sys.argv[1],
                           server_speeds=[373, 373, 746, 746], policy="sed"))
    print("== 4096 envs x 16 servers in one launch per step")
    env = VecLoadBalanceEnv(4096, num_servers=16, max_steps=100)
    env.set_speeds([1, 2] * 8)
    env.gen_poisson(rate=32.0, mean_work=0.6, horizon=26.0, seed=1)
    env.reset()
    rng = np.random.RandomState(0)
    for step in range(100):
        obs, reward, done = env.step(rng.randint(0, 3, (4096, 16)).astype(np.int32))
    print(f"  after 100 steps: mean jain {float(reward.mean()):.4f}, flows in system {int(obs[..., 0].sum())}")
