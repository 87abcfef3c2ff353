# A/B of the small-launch statistics-pass variants on c2 / c4 (env vars read at env creation)
for wl in c2 c4; do for knob in "" "MLB_PAIR_SERIAL=1" "MLB_PAIR_WPE1=1"; do
  env $knob python bench.py --workload $wl --no-cpu --no-configs --late-burnin 0 --e2e-steps 2 --steps 50 2>/dev/null | python -c "
import sys,json
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); ks=j['roofline']['kernels']
print('$wl [$knob]', 'ms/step %.4f value %.3e' % (j['ms_per_step'], j['value']), [ (k['kernel'][:22], round(k['ms'],4)) for k in ks])"
done; done
