#!/bin/bash
# quick GPU check: parity tests of the env path + short bench line (ms/step, agent-steps/s, roofline frac, e2e)
python -m pytest tests -m gpu -x -q ${KB_TESTS:-} 2>&1 | tail -${KB_TAIL:-4}
python bench.py --steps 50 --warmup 5 --no-cpu --e2e-steps 3 $@ 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('ms/step %.3f  value %.3fM  frac %.3f  e2e %.3fM' % (j['ms_per_step'], j['value']/1e6, j['roofline']['frac'], j['e2e']['value']/1e6))
"
