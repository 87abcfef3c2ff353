"""GPU: the shipped examples run (examples/random_policy.py: legacy mode, flow mode, batched env)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_policy_example_runs():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "random_policy.py")], check=True,
                         capture_output=True, text=True, timeout=600, env=env).stdout
    assert "legacy mode" in out and "flow mode" in out and "after 100 steps" in out
    # the first reward of LoadBalanceEnv(num_servers=4, seed=42) after reset is reproducible
    assert out.count("mean return") == 2
