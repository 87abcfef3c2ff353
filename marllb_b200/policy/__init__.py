"""Batched policy inference / update on hand-written CUDA kernels (QMIX, SAC-GRU).

API mirrors of simulation-mode/problem-05-qmix/src and problem-04-sac-gru/src; all arithmetic
runs in marllb_b200/csrc/mlb_policy.cu through include/marllb_b200_policy.h.
"""
from .qmix import (AgentQNetwork, EpisodeBuffer, QMixingNetwork, QMIXAgent, VDNMixingNetwork,  # noqa: F401
                   WeightedQMixingNetwork)
from .sac import PolicyNetwork, PrioritizedReplayBuffer, QNetwork, ReplayBuffer, SAC_GRU_Agent  # noqa: F401
from . import paper  # noqa: F401  (original-paper agents: RNNAgent, QMix, QMix_Trainer, discrete SAC)
