"""Deterministic inputs for the full-size policy parity fixtures (config C3 QMIX update, config C4 SAC update).

Shared by tests/golden/make_fullsize_golden.py (runs the UNMODIFIED reference agents on torch CPU, in the build
container) and tests/test_gpu_policy_fullsize.py (runs this package's agents on the GPU box, where /root/reference does
not exist).  Everything here comes from numpy's RandomState, whose streams are identical on both sides, so the
fixture only has to carry the reference's *outputs* (a few sampled entries per tensor), not megabytes of inputs.
"""
import numpy as np

# ---- config C4: SAC-GRU learner of the 256-server env (bench.py --workload c4; BASELINE.json config 4)
SAC = dict(state_dim=256 * 11, action_dim=256, hidden_dim=256, gru_dim=128, batch_size=256)
# ---- config C3: QMIX learner of 2 agents x 32 servers (bench.py --workload c3; BASELINE.json config 3)
QMIX = dict(num_agents=2, state_dim=4 * 64 + 10, obs_dim=32 * 11, action_dim=32, hidden_dim=64, gru_dim=64,
            mixing_embed_dim=32, hypernet_embed_dim=64, batch_size=32, max_seq_len=50)
N_UPDATES = 3
N_SAMPLES = 48       # sampled entries per tensor kept in the fixture


def synth_state_dict(shapes, seed):
    """{name: float32 array}: weights ~ N(0, 1/fan_in), biases ~ N(0, 0.05^2); one RandomState stream per tensor so
    that the values do not depend on dict order."""
    out = {}
    for k, (name, shape) in enumerate(sorted(shapes.items())):
        rng = np.random.RandomState(seed * 1000 + k)
        x = rng.standard_normal(shape)
        scale = 0.05 if len(shape) == 1 else 1.0 / np.sqrt(shape[-1])
        out[name] = (x * scale).astype(np.float32)
    return out


def sample_index(name, numel, salt):
    """Flat indices of the entries of tensor `name` that the fixture keeps."""
    h = sum((i + 1) * ord(c) for i, c in enumerate(name)) % 100003
    rng = np.random.RandomState(salt * 100003 + h)
    return rng.randint(0, numel, size=min(N_SAMPLES, numel))


def sac_batch(u):
    """Batch + the two noise tensors (next-state sample first, then the new-action sample: sac_agent.py:176,211) of update u."""
    c = SAC
    rng = np.random.RandomState(7000 + u)
    B, S, A, G = c["batch_size"], c["state_dim"], c["action_dim"], c["gru_dim"]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    batch = (f(rng.standard_normal((B, S))), f(np.tanh(rng.standard_normal((B, A)))), f(rng.random_sample((B, 1))),
             f(rng.standard_normal((B, S))), f(rng.random_sample((B, 1)) < 0.1), f(rng.standard_normal((1, B, G)) * 0.2))
    eps_next, eps_new = f(rng.standard_normal((B, A))), f(rng.standard_normal((B, A)))
    return batch, eps_next, eps_new


def qmix_batch(u):
    """Episode batch of update u in EpisodeBuffer.sample_batch's layout, ragged lengths, zero padding (episode_buffer.py)."""
    c = QMIX
    rng = np.random.RandomState(8000 + u)
    B, T, A = c["batch_size"], c["max_seq_len"], c["num_agents"]
    batch = {'observations': rng.standard_normal((B, T, A, c["obs_dim"])),
             'actions': rng.randint(0, c["action_dim"], (B, T, A, 1)).astype(np.float64),
             'rewards': rng.random_sample((B, T, A)), 'states': rng.standard_normal((B, T, c["state_dim"])),
             'dones': np.zeros((B, T)), 'seq_lengths': rng.randint(T // 2, T + 1, size=B).astype(np.int32)}
    batch['seq_lengths'][:4] = T
    for b, L in enumerate(batch['seq_lengths']):
        for k in ('observations', 'actions', 'rewards', 'states'):
            batch[k][b, L:] = 0
        batch['dones'][b, L - 1] = 1.0
    return batch
