"""Wire format between the RL side and the VPP load-balancer plugin.

Byte-for-byte the layouts of the reference
(simulation-mode/problem-02-shared-memory-ipc/src/shm_layout.py:26-279; C side src/vpp/lb/shm.h:14-91):

  msg_out (VPP -> RL, observations)   '=QQQIIx' + 4 pad bytes header (37 B), then 64 x { u32 n_flow_on, f32[10] }
  msg_in  (RL -> VPP, actions)        '=QQII' header (24 B), f32 weights[64], 64 x { f32 prob, u32 alias }

so that a policy trained on the batched GPU env can drive the real data plane: `obs_to_msg_out`
turns (E, S, 11) observations into msg_out records (what the plugin would have sent),
`msg_out_to_obs` is the way back, `actions_to_msg_in` emits the weights together with the alias
table of `rl_controller._build_alias_table` (problem-06-vpp-integration/src/rl_controller.py:359-405).
Vectorised with numpy structured dtypes: E records at once, no per-field Python loops.
Host-side only (the shared-memory segment lives on the host).
"""
from __future__ import annotations

import time

import numpy as np

MAX_AS = 64                 # shm_layout.py:20
RESERVOIR_CAPACITY = 128
NUM_FEATURES = 5
RING_BUFFER_SIZE = 4

_SERVER = np.dtype([("n_flow_on", "<u4"), ("reservoir_features", "<f4", (10,))])
MSG_OUT = np.dtype([("sequence_id", "<u8"), ("timestamp_us", "<u8"), ("active_as_bitmap", "<u8"),
                    ("num_active_as", "<u4"), ("reserved", "<u4"), ("pad", "u1", (5,)),
                    ("as_stats", _SERVER, (MAX_AS,))])            # HEADER_FORMAT '=QQQIIx' + 'xxxx' = 37 bytes
_ALIAS = np.dtype([("prob", "<f4"), ("alias", "<u4")])
MSG_IN = np.dtype([("sequence_id", "<u8"), ("timestamp_us", "<u8"), ("num_servers", "<u4"), ("reserved", "<u4"),
                   ("weights", "<f4", (MAX_AS,)), ("alias_table", _ALIAS, (MAX_AS,))])
MSG_OUT_SIZE = MSG_OUT.itemsize     # 37 + 44 * 64 = 2853  (MessageOutLayout.MESSAGE_SIZE)
MSG_IN_SIZE = MSG_IN.itemsize       # 24 + 256 + 512 = 792 (MessageInLayout.MESSAGE_SIZE)


def obs_to_msg_out(obs, sequence_id=0, timestamp_us=None):
    """(E, S, 11) or (S, 11) float32 observations -> E msg_out records (numpy structured array;
    `.tobytes()` of one record is what MessageOutLayout.pack produces).  A server is active when any
    of its values is > 0 (env.py:410-413)."""
    obs = np.asarray(obs, np.float32)
    if obs.ndim == 2:
        obs = obs[None]
    E, S, C = obs.shape
    if S > MAX_AS or C != 11:
        raise ValueError(f"at most {MAX_AS} servers and exactly 11 columns")
    out = np.zeros(E, MSG_OUT)
    out["sequence_id"] = sequence_id
    out["timestamp_us"] = int(time.time() * 1e6) if timestamp_us is None else timestamp_us
    active = (obs > 0).any(axis=2)
    out["active_as_bitmap"] = (active.astype(np.uint64) << np.arange(S, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
    out["num_active_as"] = active.sum(axis=1)
    out["as_stats"]["n_flow_on"][:, :S] = obs[:, :, 0].astype(np.uint32)
    out["as_stats"]["reservoir_features"][:, :S] = obs[:, :, 1:]
    return out


def msg_out_to_obs(msgs, num_servers):
    """msg_out records (structured array or bytes) -> (E, num_servers, 11) float32; inactive servers
    are zero rows like `_dict_to_array` makes them (env.py:355-389)."""
    if isinstance(msgs, (bytes, bytearray, memoryview)):
        msgs = np.frombuffer(msgs, MSG_OUT)
    msgs = np.atleast_1d(msgs)
    E = len(msgs)
    obs = np.zeros((E, num_servers, 11), np.float32)
    obs[:, :, 0] = msgs["as_stats"]["n_flow_on"][:, :num_servers]
    obs[:, :, 1:] = msgs["as_stats"]["reservoir_features"][:, :num_servers]
    active = (msgs["active_as_bitmap"][:, None] >> np.arange(num_servers, dtype=np.uint64)) & np.uint64(1)
    return obs * active[:, :, None].astype(np.float32)


def build_alias_table(weights):
    """rl_controller._build_alias_table (rl_controller.py:359-405): prob = w * n in float64, small /
    large stacks popped from the end.  -> (prob float64 [n], alias int32 [n])."""
    w = np.array(weights, dtype=np.float64)
    n = len(w)
    prob = w * n
    alias = np.arange(n, dtype=np.int32)
    small = [i for i in range(n) if prob[i] < 1.0]
    large = [i for i in range(n) if not prob[i] < 1.0]
    while small and large:
        l, g = small.pop(), large.pop()
        alias[l] = g
        prob[g] = prob[g] + prob[l] - 1.0
        (small if prob[g] < 1.0 else large).append(g)
    return prob, alias


def weights_from_server_action(action, discrete_weights=(1.0, 1.5, 2.0)):
    """Per-server discrete action indices -> normalised server weights (env.py:334-353, then the
    normalisation of rl_controller.py:322-324)."""
    w = np.asarray(discrete_weights, np.float64)[np.asarray(action)]
    return w / w.sum(axis=-1, keepdims=True)


def actions_to_msg_in(weights, sequence_id=0, timestamp_us=None, with_alias=True):
    """(E, S) or (S,) server weights -> E msg_in records (`.tobytes()` of one = MessageInLayout.pack)."""
    weights = np.asarray(weights, np.float64)
    if weights.ndim == 1:
        weights = weights[None]
    E, S = weights.shape
    if S > MAX_AS:
        raise ValueError(f"at most {MAX_AS} servers")
    out = np.zeros(E, MSG_IN)
    out["sequence_id"] = sequence_id
    out["timestamp_us"] = int(time.time() * 1e6) if timestamp_us is None else timestamp_us
    out["num_servers"] = S
    out["weights"][:, :S] = weights.astype(np.float32)
    if with_alias:
        for e in range(E):
            prob, alias = build_alias_table(weights[e])
            out["alias_table"]["prob"][e, :S] = prob.astype(np.float32)
            out["alias_table"]["alias"][e, :S] = alias
    return out


def msg_in_to_actions(msgs):
    """-> list of dicts like MessageInLayout.unpack (weights and alias table cut to num_servers)."""
    if isinstance(msgs, (bytes, bytearray, memoryview)):
        msgs = np.frombuffer(msgs, MSG_IN)
    out = []
    for m in np.atleast_1d(msgs):
        n = int(m["num_servers"])
        out.append({"sequence_id": int(m["sequence_id"]), "timestamp_us": int(m["timestamp_us"]),
                    "timestamp": int(m["timestamp_us"]) / 1e6, "num_servers": n,
                    "weights": m["weights"][:n].tolist(),
                    "alias_table": [(float(p), int(a)) for p, a in m["alias_table"][:n]]})
    return out
