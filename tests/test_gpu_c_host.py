"""GPU: the C ABI driven from a plain C program (examples/c_host_step.c: no CUDA headers, no Python, host buffers on
both sides) gives the same env trajectory as the Python classes -- the drop-in boundary is the shared library, not
the Python package."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_host_program_matches_python_path(tmp_path):
    from marllb_b200 import VecLoadBalanceEnv, _build
    lib = _build.build()
    exe = str(tmp_path / "c_host_step")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host_step.c"),
                    "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib), "-lm"], check=True)
    E, S, steps = 64, 16, 20
    out = subprocess.run([exe, str(E), str(S), str(steps)], check=True, capture_output=True, text=True, timeout=300).stdout
    line = next(l for l in out.splitlines() if l.startswith("checksum"))
    flows_c, reward_c = int(line.split()[1]), float(line.split()[2])
    assert "kernels launched" in out

    env = VecLoadBalanceEnv(E, num_servers=S, max_steps=steps)
    env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0).astype(np.float32))
    rate = 2.0 * S
    env.gen_poisson(rate, 0.8 * 1.5 * S / rate, steps * 0.25 + 1.0, seed=1234)
    env.reset()
    lcg, total = np.uint32(12345), 0.0
    with np.errstate(over="ignore"):
        for _ in range(steps):
            act = np.empty(E * S, np.int32)
            for i in range(E * S):
                lcg = np.uint32(lcg * np.uint32(1664525) + np.uint32(1013904223))
                act[i] = (int(lcg) >> 16) % 3
            _, rew, done = env.step(act.reshape(E, S))
            total += float(rew.sum().item())
    env.check_status()
    assert bool(done.all())
    assert int(env.get_state("n_flow_on").sum()) == flows_c
    assert total == pytest.approx(reward_c, rel=1e-12)
    env.close()
