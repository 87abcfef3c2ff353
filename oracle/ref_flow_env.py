"""Composed flow-level env built from the UNMODIFIED reference's own classes.

TEST INFRASTRUCTURE ONLY -- never imported by marllb_b200/ (the product).
Runs only where /root/reference exists (the build container).  Its job is to
pin oracle/flow_oracle.c (the travelling C restatement) against the reference's
real `ReservoirSampler`, `MultiMetricReservoir`, `PerServerFeatures`,
`LoadBalanceEnv._action_to_weights/_array_to_dict` and `RewardFunction`:
tests/golden/make_golden.py runs this class and freezes its outputs.

Why a composition: the reference's simulation-mode `step()` draws random
features and never simulates flows (env.py:215-286, 425-448); SURVEY.md App. B
defines the flow-level step from the pieces the reference does ship.  Rules
marked [R file:line] follow the reference, [B] are builder decisions, frozen
here before any kernel was written.

Per env: A agents, each owning S_a servers (partition as
problem-05-qmix/src/multi_agent_env.py:71-76); S = A*S_a rows in the obs.

step(action), window k -> [t0, t1) with t1 = f32(k+1)*f32(dt)       [R env.py:80]
  w = _action_to_weights(action)                                    [R env.py:334-353]
  for every agent, for every arrival (a, work) with a < t1, in order:
    retire(now=a): per own server, pop ring head while fin < now:
        n_flow_on -= 1                                              [R src/vpp/lb/lbhash.h:120]
        fct.add(value=f32(fin-arr), timestamp=fin)                  [R lbhash.h:122-124, reservoir.py:50]
    choose server (SED / LSQ / ALIAS)                               [R src/vpp/lb/node.c:393-460]
    queue full (tail-head >= Q) -> drop-and-count                   [B]
    else start=max(last_fin,a); fin=f32(start+f32(work/speed)); push [R paper Alg.1 l.19, event-time form B]
         n_flow_on += 1                                             [R lbhash.h:142,167]
  retire(now=t1)
  per server, per still-active flow in arrival order:
        flow_duration.add(value=f32(t1-arr), timestamp=t1)          [R lbhash.h:131-135; one per step B]
  obs[j] = [n_flow_on | fct feats | flow_duration feats]            [R features.py:256-286]
  reward = RewardFunction.compute(_array_to_dict(obs))              [R env.py:259-262]
  done = step >= max_steps                                          [R env.py:267]

All simulated times are float32 seconds since reset (the VPP layout stores
f32 time/value pairs, src/vpp/lb/shm.h:23-25); they are handed to the
reference's reservoir as exact Python floats, so its float64 timestamp array
holds exactly the float32 values the CUDA kernel stores.
"""
from __future__ import annotations

import numpy as np

import ref_import

f32 = np.float32

POLICIES = ("sed", "lsq", "alias", "sed2", "lsq2")


def build_alias_table_ref(p):
    """Reference alias builder, called unbound (rl_controller.py:359-405)."""
    try:
        rc = ref_import.load("rl_controller", "problem-06-vpp-integration")
        return rc.RLController._build_alias_table(None, p)
    except Exception:  # heavy imports unavailable -> restatement of the same 30 lines
        n = len(p)
        prob = np.array(p, dtype=np.float64) * n
        alias = np.arange(n, dtype=np.int32)
        small = [i for i, q in enumerate(prob) if q < 1.0]
        large = [i for i, q in enumerate(prob) if not q < 1.0]
        while small and large:
            l = small.pop()
            g = large.pop()
            alias[l] = g
            prob[g] = prob[g] + prob[l] - 1.0
            (small if prob[g] < 1.0 else large).append(g)
        return [(prob[i], alias[i]) for i in range(n)]


class RefFlowEnv:
    def __init__(self, num_agents, servers_per_agent, speeds, arrivals,
                 reservoir_k=128, queue_cap=160, dt=0.25, decay=0.9,
                 policy="sed", action_type="discrete", discrete_weights=None,
                 min_weight=0.1, max_weight=10.0, reward_metric="jain",
                 reward_field="flow_duration_avg_decay", max_steps=10000,
                 seed_base=0):
        res = ref_import.load("reservoir")
        feats = ref_import.load("features")
        env = ref_import.load("env")
        self.A = int(num_agents)
        self.Sa = int(servers_per_agent)
        self.S = self.A * self.Sa
        self.K = int(reservoir_k)
        self.Q = int(queue_cap)
        self.dt = f32(dt)
        self.decay = float(decay)
        assert policy in POLICIES
        self.policy = policy
        self.max_steps = int(max_steps)
        self.speeds = np.asarray(speeds, dtype=np.float32).reshape(self.S)
        # arrivals: list (per agent) of dicts with 'time','work' [, 'bucket','u']
        self.arrivals = arrivals
        self._MultiMetricReservoir = res.MultiMetricReservoir
        self._PerServerFeatures = feats.PerServerFeatures
        self.seed_base = int(seed_base)
        # a reference env instance supplies _action_to_weights/_array_to_dict/reward_fn
        self.ref_env = env.LoadBalanceEnv(
            num_servers=self.S, action_type=action_type,
            discrete_weights=discrete_weights, max_weight=max_weight,
            min_weight=min_weight, reward_metric=reward_metric,
            reward_field=reward_field, step_interval=0.0,
            max_steps=max_steps, use_shm=False)
        self.reset()

    # ------------------------------------------------------------------
    def reset(self):
        S = self.S
        self.cur_step = 0
        self.reservoirs = [
            self._MultiMetricReservoir(metrics=["fct", "flow_duration"],
                                       capacity=self.K, seed=self.seed_base + j)
            for j in range(S)]                       # [R basic_usage.py:157-163]
        self.psf = self._PerServerFeatures(S)
        self.last_fin = np.zeros(S, dtype=np.float32)
        self.rings = [[] for _ in range(S)]          # FIFO of (arr f32, fin f32)
        self.dropped = np.zeros(S, dtype=np.int64)
        self.cursor = [0] * self.A
        self.ref_env.current_step = 0
        self.ref_env.episode_return = 0.0
        return np.zeros((S, 11), dtype=np.float32)

    # ------------------------------------------------------------------
    def _retire(self, j, now):
        ring = self.rings[j]
        while ring and ring[0][1] < now:
            arr, fin = ring.pop(0)
            self.psf.n_flow_on[j] -= 1
            self.reservoirs[j].add("fct", float(f32(fin - arr)), timestamp=float(fin))

    def _choose(self, agent, w, k_flow):
        lo = agent * self.Sa
        n_on = self.psf.n_flow_on
        if self.policy == "sed":
            best, best_score = lo, None
            for j in range(lo, lo + self.Sa):
                score = f32((int(n_on[j]) + 1) / (1e-9 + float(w[j])))   # [R node.c:395-404]
                if best_score is None or score < best_score:
                    best, best_score = j, score
            return best
        if self.policy == "lsq":
            best, best_score = lo, None
            for j in range(lo, lo + self.Sa):
                score = f32(int(n_on[j]))                                # [R node.c:419-431]
                if best_score is None or score < best_score:
                    best, best_score = j, score
            return best
        b = int(self.arrivals[agent]["bucket"][k_flow])
        if self.policy in ("sed2", "lsq2"):
            # power of two choices [R node.c:408-417, 433-441]: asindex0 = new_flow_table[hash & mask],
            # asindex1 = new_flow_table[(hash + 1) & mask]; asindex1 wins on a strictly lower score.
            # [B] the pre-drawn bucket plays the hash, the flow table is the identity modulo Sa.
            c0, c1 = lo + b, lo + (b + 1) % self.Sa
            if self.policy == "sed2":
                s0 = f32((int(n_on[c0]) + 1) / (1e-9 + float(w[c0])))
                s1 = f32((int(n_on[c1]) + 1) / (1e-9 + float(w[c1])))
            else:
                s0, s1 = f32(int(n_on[c0])), f32(int(n_on[c1]))
            return c1 if s1 < s0 else c0
        u = float(f32(self.arrivals[agent]["u"][k_flow]))
        prob, alias = self._alias[agent][b]
        return lo + (b if u < prob else int(alias))                      # [R test_integration.py:57-63]

    def step(self, action):
        env = self.ref_env
        env.current_step += 1
        self.cur_step += 1
        t1 = f32(f32(self.cur_step) * self.dt)
        w = env._action_to_weights(np.asarray(action))
        if self.policy == "alias":
            self._alias = []
            for i in range(self.A):
                wi = w[i * self.Sa:(i + 1) * self.Sa].astype(np.float64)
                self._alias.append(build_alias_table_ref(wi / wi.sum()))
        assign = [[] for _ in range(self.A)]
        for i in range(self.A):
            times = self.arrivals[i]["time"]
            works = self.arrivals[i]["work"]
            c = self.cursor[i]
            while c < len(times) and f32(times[c]) < t1:
                a = f32(times[c])
                wk = f32(works[c])
                for j in range(i * self.Sa, (i + 1) * self.Sa):
                    self._retire(j, a)
                j = self._choose(i, w, c)
                assign[i].append(j)
                if len(self.rings[j]) >= self.Q:
                    self.dropped[j] += 1
                else:
                    start = self.last_fin[j] if self.last_fin[j] > a else a
                    fin = f32(start + f32(wk / self.speeds[j]))
                    self.rings[j].append((a, fin))
                    self.last_fin[j] = fin
                    self.psf.n_flow_on[j] += 1
                c += 1
            self.cursor[i] = c
        feats = []
        for j in range(self.S):
            self._retire(j, t1)
            for arr, _fin in self.rings[j]:
                self.reservoirs[j].add("flow_duration", float(f32(t1 - arr)), timestamp=float(t1))
            feats.append(self.reservoirs[j].get_feature_vector(self.decay, current_time=float(t1)))
        obs = self.psf.get_state_vector(feats)
        obs_dict = env._array_to_dict(obs)
        reward = float(env.reward_fn.compute(obs_dict))
        done = env.current_step >= env.max_steps
        info = {"step": env.current_step, "weights": w.tolist(),
                "active_servers": obs_dict["active_servers"], "assign": assign,
                "n_dropped": int(self.dropped.sum())}
        return obs, reward, done, info

    # state dumps for bit-exact comparison -------------------------------
    def dump(self):
        S, K = self.S, self.K
        vals = np.zeros((S, 2, K), np.float32)
        ts = np.zeros((S, 2, K), np.float32)
        cnt = np.zeros((S, 2), np.int64)
        for j in range(S):
            for m, name in enumerate(("fct", "flow_duration")):
                r = self.reservoirs[j].reservoirs[name]
                vals[j, m] = r.values
                ts[j, m] = r.timestamps.astype(np.float32)   # exact: they are f32 values
                assert np.array_equal(ts[j, m].astype(np.float64), r.timestamps)
                cnt[j, m] = r.count
        return {"n_flow_on": self.psf.n_flow_on.astype(np.int32).copy(),
                "res_values": vals, "res_ts": ts, "res_count": cnt,
                "dropped": self.dropped.copy(),
                "last_fin": self.last_fin.copy()}
