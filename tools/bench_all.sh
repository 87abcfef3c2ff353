#!/bin/bash
# one short bench line per workload: tools/bench_all.sh [workloads...]  -> gpurun_out/wl_<name>.json
for w in ${@:-c2 c3 c4}; do
  python bench.py --workload $w --no-cpu --e2e-steps 3 > gpurun_out/wl_$w.json 2> gpurun_out/wl_$w.err
  python - $w <<'PY'
import json, sys
w = sys.argv[1]
j = json.loads(open(f"gpurun_out/wl_{w}.json").read().strip().splitlines()[-1])
r = j["roofline"]
print(w, "ms/step %.3f" % j["ms_per_step"], "value %.3e" % j["value"], "step frac", r.get("step", {}).get("frac"),
      "policy ms", r.get("policy_ms_per_step"), "e2e %.3e" % j["e2e"]["value"])
PY
done
