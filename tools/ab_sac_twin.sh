# A/B: SAC update with the twin critics on two streams (default) vs one stream (MLB_SAC_TWIN=0)
for knob in 0 1; do
  echo "== MLB_SAC_TWIN=$knob"; MLB_SAC_TWIN=$knob python tools/policy_bench.py 2>&1 | grep -i "sac update"
  MLB_SAC_TWIN=$knob python bench.py --workload c4 --no-cpu --no-configs --late-burnin 0 --e2e-steps 2 --steps 50 2>/dev/null | python -c "
import sys,json
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c4 ms/step %.4f value %.3e' % (j['ms_per_step'], j['value']))"
done
