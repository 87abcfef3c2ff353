"""One-off scale check: the bench-size parity case of tests/test_gpu_bench_parity.py over a whole long window
(default 2000 device steps + 100 host-buffer steps at 131072 envs x 64 servers): eight envs spread over the index range
against the C oracle at every step, reservoir dumps at the end.  python tools/deep_parity.py [workload dev_steps host_steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import test_gpu_bench_parity as t
wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
dev, host = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2000, 100)
t0 = time.time()
t._run_case(wl, dev, host, continuous=(wl == "c4"))
print(f"{wl}: {dev} device steps + {host} host-buffer steps, 8 sampled envs: integers bit-exact, obs 1e-5, rewards 1e-9, "
      f"reservoir dumps equal ({time.time() - t0:.0f} s)")
