"""marllb_b200 -- B200-native implementation of MARLLB's simulation-mode hot path.

Public surface (reference API names kept):
    LoadBalanceEnv, LoadBalanceEnvGym      problem-03-rl-environment/src/env.py
    MultiAgentLoadBalanceEnv              problem-05-qmix/src/multi_agent_env.py
    RewardFunction + metric functions     problem-03-rl-environment/src/rewards.py
    ReservoirSampler, MultiMetricReservoir, PerServerFeatures
                                          problem-01-reservoir-sampling/src/{reservoir,features}.py
    VecLoadBalanceEnv, BatchedReservoirs  batched forms (new)
    policy.{QMIXAgent, AgentQNetwork, QMixingNetwork, EpisodeBuffer, SAC_GRU_Agent, PolicyNetwork, QNetwork, ...}
                                          problem-05-qmix/src, problem-04-sac-gru/src
    QMIXRollout, SACRollout, DeviceReplay env step fused with batched policy inference (CUDA graph)
    training_pipeline.TrainingPipeline    problem-06-vpp-integration/src/training_pipeline.py
    wire                                  problem-02-shared-memory-ipc/src/shm_layout.py (msg_out / msg_in)
All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
include/marllb_b200.h; there is no CPU fallback.
"""
from ._build import build  # noqa: F401

__all__ = ["build", "LoadBalanceEnv", "LoadBalanceEnvGym", "MultiAgentLoadBalanceEnv",
           "VecLoadBalanceEnv", "VecLegacyEnv", "RewardFunction", "ReservoirSampler", "MultiMetricReservoir",
           "PerServerFeatures", "BatchedReservoirs", "QMIXRollout", "SACRollout", "DeviceReplay"]

_LAZY = {
    "LoadBalanceEnv": "env", "LoadBalanceEnvGym": "env",
    "MultiAgentLoadBalanceEnv": "multi_agent_env", "VecLoadBalanceEnv": "vec_env", "VecLegacyEnv": "legacy_vec",
    "RewardFunction": "rewards", "ReservoirSampler": "reservoir",
    "MultiMetricReservoir": "reservoir", "PerServerFeatures": "reservoir",
    "BatchedReservoirs": "reservoir",
    "QMIXRollout": "rollout", "SACRollout": "rollout", "DeviceReplay": "rollout",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(f".{_LAZY[name]}", __name__)
        return getattr(mod, name)
    raise AttributeError(name)
