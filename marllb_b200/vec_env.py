"""Batched flow-level load-balancer environment on one B200.

`VecLoadBalanceEnv` is the batched counterpart of the reference's
`LoadBalanceEnv` (simulation-mode/problem-03-rl-environment/src/env.py:41-470)
and `MultiAgentLoadBalanceEnv` (problem-05-qmix/src/multi_agent_env.py:22-290):
the same constructor vocabulary, `reset()` / `step(action)`, but for E
independent env instances stepped by one fused CUDA kernel through the C ABI
(include/marllb_b200.h).  Observations, rewards and done flags are returned as
torch CUDA tensors that alias extension-owned buffers (valid until the next
step); `step_host` is the end-to-end variant with pinned host buffers.

PyTorch is used only for device/pinned memory and stream handles.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Config, check

_TORCH_TYPESTR = {torch.float32: "<f4", torch.float64: "<f8", torch.uint8: "|u1",
                  torch.int32: "<i4", torch.int64: "<i8"}


class _DevView:
    """Expose a raw device pointer through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, shape, dtype, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": _TORCH_TYPESTR[dtype],
                                         "data": (int(ptr), False), "version": 2, "strides": None}
        self._owner = owner  # keep the handle alive while views exist


def _dptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _nptr(a) -> C.c_void_p:
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class VecLoadBalanceEnv:
    """E independent LB envs (A agents x Sa servers each) stepped on one GPU.

    Args mirror env.py:71-87 (`num_servers` = servers per agent, `action_type`,
    `discrete_weights`, `max_weight`, `min_weight`, `reward_metric`,
    `reward_field`, `step_interval`, `max_steps`) plus `num_envs`, `num_agents`
    (multi_agent_env.py:44-53) and the flow-level knobs of SURVEY App. B.
    `rng_mode`: "replay" (default) -- reservoirs of server j replay np.random.RandomState(seed_base + j), the
    reference's stream (reservoir.py:45,76), the same in every env; "philox" -- a counter-based stream per
    (global env, server) with the same masked-rejection rule (include/marllb_b200.h), for large batches whose
    envs should not share sampling decisions.
    """

    def __init__(self, num_envs: int, num_servers: int = 4, num_agents: int = 1,
                 action_type: str = "discrete", discrete_weights: Optional[Sequence[float]] = None,
                 max_weight: float = 10.0, min_weight: float = 0.1, reward_metric: str = "jain",
                 reward_field="flow_duration_avg_decay", step_interval: float = 0.25,
                 max_steps: int = 10000, policy: str = "sed", reservoir_capacity: int = 128,
                 queue_capacity: int = 160, decay: float = 0.9, seed_base: int = 0,
                 rng_table_len: int = 65536, feature_cache: bool = True,
                 record_assign: bool = False, action_dtype: str = "int32", env_id_base: int = 0,
                 device: int = 0, normalize_obs: bool = False, rng_mode: str = "replay"):
        if action_type not in ("discrete", "continuous"):
            raise ValueError(f"Unknown action_type: {action_type}")          # env.py:184
        if reward_metric not in _lib.METRICS:
            raise ValueError(f"Unsupported metric: {reward_metric}. "
                             f"Supported: {list(_lib.METRICS.keys())}")      # rewards.py:321-323
        if policy not in _lib.POLICIES:
            raise ValueError(f"Unknown policy: {policy}")
        if rng_mode not in _lib.RNG_MODES:
            raise ValueError(f"Unknown rng_mode: {rng_mode} (replay: the reference's RandomState(seed_base + server) "
                             f"stream shared by all envs; philox: counter-based, independent per env)")
        if isinstance(reward_field, str):
            if reward_field not in _lib.FEATURE_NAMES:
                raise ValueError(f"Unknown reward_field: {reward_field}")
            reward_field = _lib.FEATURE_NAMES.index(reward_field)
        self._L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("marllb_b200 needs a CUDA device (there is no CPU fallback)")
        cfg = Config()
        check(self._L.mlb_config_default(C.byref(cfg)))
        cfg.device = device
        cfg.num_envs, cfg.num_agents, cfg.servers_per_agent = num_envs, num_agents, num_servers
        cfg.reservoir_k, cfg.queue_cap = reservoir_capacity, queue_capacity
        cfg.policy = _lib.POLICIES[policy]
        if action_type == "continuous":
            cfg.action_kind = _lib.ACTION_CONTINUOUS_F32
        else:
            cfg.action_kind = _lib.ACTION_DISCRETE_U8 if action_dtype == "uint8" else _lib.ACTION_DISCRETE_I32
        dw = list(discrete_weights) if discrete_weights else [1.0, 1.5, 2.0]   # env.py:69
        if len(dw) > 8:
            raise ValueError("at most 8 discrete weight levels")
        cfg.n_discrete = len(dw)
        for i, w in enumerate(dw):
            cfg.discrete_weights[i] = w
        cfg.min_weight, cfg.max_weight = min_weight, max_weight
        cfg.dt, cfg.decay = step_interval, decay
        cfg.reward_metric, cfg.reward_field = _lib.METRICS[reward_metric], int(reward_field)
        cfg.max_steps = max_steps
        cfg.rng_seed_base, cfg.rng_table_len = seed_base, rng_table_len
        cfg.feature_cache, cfg.record_assign = int(feature_cache), int(record_assign)
        cfg.env_id_base = env_id_base
        cfg.rng_mode = _lib.RNG_MODES[rng_mode]
        self.rng_mode = rng_mode
        self.cfg = cfg
        self.num_envs, self.num_agents, self.servers_per_agent = num_envs, num_agents, num_servers
        self.total_servers = num_agents * num_servers
        self.action_type, self.policy = action_type, policy
        self.discrete_weights = dw
        self.max_steps = max_steps
        self.reservoir_capacity = reservoir_capacity
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        check(self._L.mlb_create(C.byref(cfg), C.byref(self._h)))
        E, S = num_envs, self.total_servers
        self.obs = self._view(_lib.PTR_OBS, (E, S, 11), torch.float32)
        self.reward = self._view(_lib.PTR_REWARD, (E,), torch.float64)
        self.done = self._view(_lib.PTR_DONE, (E,), torch.uint8)
        self._adtype = {_lib.ACTION_DISCRETE_I32: torch.int32, _lib.ACTION_CONTINUOUS_F32: torch.float32,
                        _lib.ACTION_DISCRETE_U8: torch.uint8}[cfg.action_kind]
        self._np_adtype = {torch.int32: np.int32, torch.float32: np.float32, torch.uint8: np.uint8}[self._adtype]
        # running observation statistics (env.py:152-154): per env, float64, device-resident
        self.normalize_obs = normalize_obs
        self.obs_count = 0
        self.obs_mean = self.obs_std = self.normalized_obs = None
        # streamed trace (attach_stream): chunk bookkeeping
        self._trace = None
        self._steps_done = 0
        # pinned host buffers for the end-to-end path
        self._h_action = self._h_obs = self._h_reward = self._h_done = None
        self._h_obs_valid = False
        self.last_d2h_bytes = 0
        import os as _os
        try:
            ncpu = len(_os.sched_getaffinity(0))
        except AttributeError:
            ncpu = _os.cpu_count() or 1
        self.host_threads = max(1, min(16, ncpu))    # threads that apply step_host(obs="changed") records

    # ------------------------------------------------------------------ utils
    def _view(self, what, shape, dtype):
        p, n = C.c_void_p(), C.c_size_t()
        check(self._L.mlb_device_ptr(self._h, what, C.byref(p), C.byref(n)), self._h)
        return torch.as_tensor(_DevView(p.value, shape, dtype, self), device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        t = getattr(self, "_prefetch", None)
        if t is not None:                                   # a chunk of a streamed trace may still be staging
            t.join()
            self._prefetch = None
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize(self.device)
            self._L.mlb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ----------------------------------------------------------------- inputs
    def set_speeds(self, speeds):
        """Per-server processing speeds: shape (S,) broadcast to all envs or (E, S)."""
        a = np.ascontiguousarray(speeds, np.float32).reshape(-1)
        check(self._L.mlb_set_speeds(self._h, _nptr(a), a.size, _lib.HOST, self._stream()), self._h)

    def load_arrivals(self, streams):
        """streams: list over (env, agent) -- env-major -- of dicts with float32 arrays
        'time', 'work' and, for the alias policy, 'bucket' (int32) and 'u' (float32); the power-of-two
        policies 'sed2' / 'lsq2' need 'bucket' only."""
        EA = self.num_envs * self.num_agents
        if len(streams) != EA:
            raise ValueError(f"expected {EA} arrival streams (num_envs*num_agents), got {len(streams)}")
        off = np.zeros(EA + 1, np.int64)
        off[1:] = np.cumsum([len(s["time"]) for s in streams])
        cat = lambda key, dt: np.ascontiguousarray(np.concatenate([np.asarray(s[key], dt) for s in streams])
                                                   if off[-1] else np.zeros(0, dt))
        time, work = cat("time", np.float32), cat("work", np.float32)
        bucket = u = None
        if self.policy == "alias":
            bucket, u = cat("bucket", np.int32), cat("u", np.float32)
        elif self.policy in ("sed2", "lsq2"):
            bucket = cat("bucket", np.int32)
        self.load_arrivals_csr(time, work, off, bucket, u)

    def load_arrivals_csr(self, time, work, offsets, bucket=None, u=None):
        time = np.ascontiguousarray(time, np.float32)
        work = np.ascontiguousarray(work, np.float32)
        offsets = np.ascontiguousarray(offsets, np.int64)
        if bucket is not None:
            bucket = np.ascontiguousarray(bucket, np.int32)
        if u is not None:
            u = np.ascontiguousarray(u, np.float32)
        check(self._L.mlb_load_arrivals(self._h, _nptr(time), _nptr(work), _nptr(bucket), _nptr(u),
                                        _nptr(offsets), _lib.HOST, self._stream()), self._h)

    # ------------------------------------------------------------- streamed traces
    def attach_stream(self, stream):
        """Replay a `traces.TraceStream`: only the chunk being stepped through (and the next one) live in
        HBM.  While the envs step through chunk c, a background thread reads and parses chunk c+1 into pinned
        memory and `mlb_stage_arrivals` copies it in on a side CUDA stream; at the chunk boundary
        `mlb_commit_arrivals` swaps the buffer sets on the stepping stream.  reset() restarts the trace."""
        import threading
        self._trace = stream
        self._threading = threading
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._restart_stream()

    def _restart_stream(self):
        tr = self._trace
        if getattr(self, "_prefetch", None) is not None:
            self._prefetch.join()                 # an error of the abandoned chunk is irrelevant after a restart
            self._prefetch = None
        tr.restart()
        self._steps_done = 0
        self._chunk_end, first = tr.next_chunk()
        self.load_arrivals_csr(first["time"], first["work"], first["offsets"], first.get("bucket"), first.get("u"))
        self._kick_prefetch()

    def _kick_prefetch(self):
        self._staged = None
        self._prefetch_error = None

        def work():
            try:
                k_end, ch = self._trace.next_chunk()
                pin = {}
                for key in ("time", "work", "bucket", "u"):
                    if key in ch:
                        t = torch.from_numpy(np.ascontiguousarray(ch[key]))
                        pin[key] = t.pin_memory() if t.numel() else t
                off = np.ascontiguousarray(ch["offsets"], np.int64)
                with torch.cuda.device(self.device):
                    check(self._L.mlb_stage_arrivals(self._h, _dptr(pin["time"]), _dptr(pin["work"]), _dptr(pin.get("bucket")),
                                                     _dptr(pin.get("u")), _nptr(off),
                                                     C.c_void_p(self._copy_stream.cuda_stream)), self._h)
                self._staged = (k_end, pin, off)          # keeps the pinned buffers alive until the next swap
            except BaseException as exc:                  # surfaced by _join_prefetch on the stepping thread
                self._prefetch_error = exc
        self._prefetch = self._threading.Thread(target=work, daemon=True)
        self._prefetch.start()

    def _join_prefetch(self):
        """Wait for the staging thread; a failure there (parse error, MlbError of mlb_stage_arrivals) is
        re-raised here instead of leaving a stale chunk behind."""
        t = getattr(self, "_prefetch", None)
        if t is not None:
            t.join()
            self._prefetch = None
        err, self._prefetch_error = getattr(self, "_prefetch_error", None), None
        if err is not None:
            raise RuntimeError(f"streamed-trace prefetch failed: {err}") from err

    def _advance_stream(self):
        """Called before a step: swap in the next chunk when the current one is used up."""
        if self._steps_done < self._chunk_end:
            return
        self._join_prefetch()
        if self._staged is None:
            raise RuntimeError("streamed-trace prefetch produced no chunk")
        self._live = self._staged
        check(self._L.mlb_commit_arrivals(self._h, self._stream()), self._h)
        self._chunk_end = self._staged[0]
        self._kick_prefetch()

    def gen_poisson(self, rate: float, mean_work: float, horizon: float, seed: int = 0, t_start: float = 0.0,
                    window: int = 0):
        """Synthetic Poisson arrivals generated on the device (training_pipeline.py:141-155) for simulated time
        [t_start, horizon).  With t_start > 0 (between two steps of a running episode, t_start = steps done * dt) the
        resident arrivals are replaced by the next window's: only one window lives in HBM at a time."""
        check(self._L.mlb_gen_poisson_window(self._h, rate, mean_work, float(t_start), float(horizon), seed,
                                             int(window), self._stream()), self._h)

    def get_arrivals(self, env: int, agent: int = 0):
        n = C.c_int64()
        check(self._L.mlb_get_arrivals(self._h, env, agent, None, None, None, None, 0, C.byref(n)), self._h)
        t, w = np.empty(n.value, np.float32), np.empty(n.value, np.float32)
        b = np.empty(n.value, np.int32) if self.policy in ("alias", "sed2", "lsq2") else None
        u = np.empty(n.value, np.float32) if self.policy == "alias" else None
        check(self._L.mlb_get_arrivals(self._h, env, agent, _nptr(t), _nptr(w), _nptr(b), _nptr(u),
                                       n.value, C.byref(n)), self._h)
        out = {"time": t, "work": w}
        if b is not None:
            out.update(bucket=b, u=u)
        return out

    # ------------------------------------------------------------- reset/step
    def reset(self, mask=None):
        """env.py:186-213 for all (or the masked) envs; returns the (E,S,11) obs view (zeros)."""
        m = None
        if mask is not None:
            m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
            if m.shape != (self.num_envs,):
                raise ValueError("mask must have shape (num_envs,)")
        if self._trace is not None:
            if m is not None:
                raise ValueError("masked reset is not available while a trace stream is attached")
            self._restart_stream()
        check(self._L.mlb_reset(self._h, _nptr(m), self._stream()), self._h)
        self._h_obs_valid = False                    # the host mirror of step_host(obs="changed") is stale
        if m is not None:
            torch.cuda.current_stream().synchronize()  # pageable mask must outlive the async copy
        if self.normalize_obs:
            return self._normalize_observation(self.obs)                           # env.py:210-211
        return self.obs

    def step(self, action):
        """One env step for every env.  `action`: (E, S) CUDA tensor (or numpy array) of
        int32/uint8 indices (discrete) or float32 weights (continuous).
        Returns (obs (E,S,11) f32, reward (E,) f64, done (E,) u8) device tensors."""
        E, S = self.num_envs, self.total_servers
        if isinstance(action, torch.Tensor):
            a = action.to(device=self.device, dtype=self._adtype).contiguous()
        else:
            a = torch.as_tensor(np.ascontiguousarray(action, self._np_adtype)).to(self.device)
        if a.numel() != E * S:
            raise ValueError(f"action must have {E}x{S} entries, got shape {tuple(a.shape)}")
        if self._trace is not None:
            self._advance_stream()
            self._steps_done += 1
        check(self._L.mlb_step(self._h, _dptr(a), _lib.DEVICE, None, None, None, _lib.DEVICE,
                               self._stream()), self._h)
        self._last_action = a  # keep alive until the kernel has consumed it
        self._h_obs_valid = False
        if self.normalize_obs:
            return self._normalize_observation(self.obs), self.reward, self.done   # env.py:283-285
        return self.obs, self.reward, self.done

    def _normalize_observation(self, obs):
        """env.py:450-470 for every env at once: updates the running `obs_mean` / `obs_std` (float64
        (E,S,11) device tensors, one set of statistics per env like one reference env object each) and
        returns the normalised observation as a float64 (E,S,11) tensor (numpy promotes to float64 in the
        reference too).  Bit-exact with the reference's arithmetic."""
        E, S = self.num_envs, self.total_servers
        if self.obs_mean is None:
            self.obs_mean = torch.zeros((E, S, 11), dtype=torch.float64, device=self.device)
            self.obs_std = torch.ones((E, S, 11), dtype=torch.float64, device=self.device)
            self.normalized_obs = torch.empty((E, S, 11), dtype=torch.float64, device=self.device)
        o = obs.to(device=self.device, dtype=torch.float32).contiguous()
        self.obs_count += 1
        check(self._L.mlb_normalize_obs(_dptr(o), _dptr(self.obs_mean), _dptr(self.obs_std), self.obs_count,
                                        _dptr(self.normalized_obs), o.numel(), self._stream()), self._h)
        return self.normalized_obs

    # ------------------------------------------------------------- CUDA graph replay of the step
    def capture(self, warmup: int = 1):
        """Capture `step(graph_action)` (the four kernels of one env step) into a CUDA graph.  Small batches
        (config C2: 4096 envs) are bound by launch gaps, not by the kernels; a replay has none.  Write the next
        actions into the static buffer `graph_action` [E, S] and call `step_graph()`.  The `warmup` + 1 steps run
        here advance the envs (call reset() afterwards for a fresh episode)."""
        if self._trace is not None or self.normalize_obs:
            raise RuntimeError("capture() covers the plain device step (no trace stream, no normalisation)")
        dev = self.device
        self.graph_action = torch.zeros((self.num_envs, self.total_servers), dtype=self._adtype, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self.graph_action)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        n0 = self.launch_count
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.step(self.graph_action)
        self.graph_launches = self.launch_count - n0          # kernels inside one replay
        self.graph_replays = 0
        return self

    def step_graph(self):
        """Replay the captured step; returns (obs, reward, done) like step()."""
        self._graph.replay()
        self.graph_replays += 1
        self._h_obs_valid = False
        return self.obs, self.reward, self.done

    def pinned_actions(self) -> "torch.Tensor":
        """A pinned (E, S) host tensor of the action dtype; fill it and pass it to step_host to
        skip the staging copy."""
        return torch.empty((self.num_envs, self.total_servers), dtype=self._adtype, pin_memory=True)

    def step_host(self, action, obs: str = "full"):
        """End-to-end step with HOST buffers: H2D of the actions from pinned memory, the env kernels, pinned D2H
        of obs / reward / done, then a stream synchronise.  With many envs the library pipelines this in chunks of
        envs (copy of chunk c overlaps the kernels of chunk c+1).  `action`: numpy array (staged through a pinned
        buffer) or a pinned torch tensor from pinned_actions() (used in place).  Returns numpy views of the pinned
        output buffers (the observation array is the same persistent buffer on every call).

        obs="full": all E*S*11 floats cross PCIe every step.  obs="changed": only the n_flow_on column and the rows
        in which a reservoir slot was written do (`mlb_step_changed`); host threads apply them to the persistent
        array.  Same result bit for bit; `last_d2h_bytes` tells what moved.  Deep into an episode Algorithm R accepts
        few samples and the difference is large; early on most rows change and the library falls back to block copies."""
        if obs not in ("full", "changed"):
            raise ValueError("obs must be 'full' or 'changed'")
        E, S = self.num_envs, self.total_servers
        if self._h_obs is None:
            self._h_obs = torch.zeros((E, S, 11), dtype=torch.float32).pin_memory()
            self._h_reward = torch.empty((E,), dtype=torch.float64, pin_memory=True)
            self._h_done = torch.empty((E,), dtype=torch.uint8, pin_memory=True)
            self._h_obs_valid = False
        if isinstance(action, torch.Tensor) and action.is_pinned() and action.dtype == self._adtype \
                and action.is_contiguous() and action.numel() == E * S:
            src = action
        else:
            if self._h_action is None:
                self._h_action = self.pinned_actions()
            self._h_action.numpy()[...] = np.asarray(action).reshape(E, S)
            src = self._h_action
        if self._trace is not None:
            self._advance_stream()
            self._steps_done += 1
        if obs == "changed" and self._h_obs_valid and not self.normalize_obs:
            moved = C.c_int64()
            check(self._L.mlb_step_changed(self._h, _dptr(src), _dptr(self._h_obs), _dptr(self._h_reward),
                                           _dptr(self._h_done), self.host_threads, C.byref(moved), self._stream()), self._h)
            self.last_d2h_bytes = int(moved.value)
        else:
            # also the key frame of the changed-rows mode: the host mirror is (re)built by one full copy
            check(self._L.mlb_step(self._h, _dptr(src), _lib.HOST, _dptr(self._h_obs),
                                   _dptr(self._h_reward), _dptr(self._h_done), _lib.HOST, self._stream()), self._h)
            torch.cuda.current_stream(self.device).synchronize()
            self.last_d2h_bytes = self.e2e_bytes[1]
            self._h_obs_valid = True
        return self._h_obs.numpy(), self._h_reward.numpy(), self._h_done.numpy()

    @property
    def e2e_bytes(self):
        """(h2d, d2h) bytes moved per step_host call."""
        E, S = self.num_envs, self.total_servers
        return E * S * torch.empty(0, dtype=self._adtype).element_size(), E * S * 11 * 4 + E * 8 + E

    def check_status(self):
        """Raise if a kernel reported a sticky error (RNG table exhausted, bad action index)."""
        check(self._L.mlb_status(self._h, self._stream()), self._h)

    @property
    def launch_count(self) -> int:
        return int(self._L.mlb_launch_count(self._h))

    def profile_begin(self, max_steps: int):
        """Record CUDA events around the two kernels of the next `max_steps` step() calls."""
        check(self._L.mlb_profile_begin(self._h, int(max_steps)), self._h)

    def profile_end(self):
        """-> (event_kernel_ms, feature_kernel_ms, steps) summed over the profiled steps."""
        ev, ft, n = C.c_double(), C.c_double(), C.c_int32()
        check(self._L.mlb_profile_end(self._h, C.byref(ev), C.byref(ft), C.byref(n)), self._h)
        self.last_pair_ms = float(self._L.mlb_profile_pair_ms(self._h))   # pair_kernel part of the statistics pass
        return ev.value, ft.value, n.value

    # ------------------------------------------------------------ state dumps
    def _state_spec(self, field):
        E, S, K = self.num_envs, self.total_servers, self.reservoir_capacity
        KP = (K + 31) // 32 * 32
        return {"n_flow_on": (_lib.F_N_FLOW_ON, torch.int32, (E, S)),
                "res_values": (_lib.F_RES_VALUES, torch.float32, (E, S, 2, KP)),
                "res_ts": (_lib.F_RES_TS, torch.float32, (E, S, 2, KP)),
                "res_count": (_lib.F_RES_COUNT, torch.int32, (E, 2, S)),       # uint32 on the device
                "res_cursor": (_lib.F_RES_CURSOR, torch.int32, (E, 2, S)),
                "dropped": (_lib.F_DROPPED, torch.int32, (E, S)),
                "last_fin": (_lib.F_LAST_FIN, torch.float32, (E, S)),
                "head": (_lib.F_HEAD, torch.int32, (E, S)),
                "step": (_lib.F_STEP, torch.int32, (E,)),
                "obs": (_lib.F_OBS, torch.float32, (E, S, 11)),
                "arr_cursor": (_lib.F_ARR_CURSOR, torch.int32, (E, self.num_agents))}[field]

    def state_view(self, field: str) -> "torch.Tensor":
        """Zero-copy device view of a state field (leading dimension = env); unsigned counters appear as int32."""
        what, dt, shape = self._state_spec(field)
        return self._view(what, shape, dt)

    def get_state(self, field: str, envs=None) -> np.ndarray:
        """Host copy of a state field; `envs` (index list) restricts it to those envs -- at bench sizes the
        reservoir arrays are tens of GB, a parity check pulls a handful of envs."""
        K = self.reservoir_capacity
        unsigned = field in ("res_count", "res_cursor", "dropped", "head")
        if envs is not None:
            idx = torch.as_tensor(np.asarray(envs, np.int64), device=self.device)
            torch.cuda.current_stream(self.device).synchronize()
            out = self.state_view(field).index_select(0, idx).cpu().numpy()
            out = out.view(np.uint32) if unsigned else out
        else:
            what, dt, shape = self._state_spec(field)
            npdt = {torch.int32: np.uint32 if unsigned else np.int32, torch.float32: np.float32}[dt]
            out = np.empty(shape, npdt)
            check(self._L.mlb_get_state(self._h, what, _nptr(out), out.nbytes, _lib.HOST), self._h)
        if field in ("res_values", "res_ts"):
            out = out[..., :K]
        return out

    def get_assignments(self, n: int) -> np.ndarray:
        out = np.empty(n, np.int32)
        check(self._L.mlb_get_assignments(self._h, _nptr(out), n, _lib.HOST, self._stream()), self._h)
        return out
