"""Import harness for the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Only usable in the build container, where /root/reference exists; it cannot
travel to the GPU box.  Used by oracle/ref_flow_env.py and
tests/golden/make_golden.py to pin the C/numpy restatements against the
reference's own code.  Nothing in marllb_b200/ may import this module.
"""
import importlib
import os
import sys

REF_ROOT = os.environ.get("MARLLB_REFERENCE", "/root/reference")
_SIM = os.path.join(REF_ROOT, "simulation-mode")
_PATHS = [
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym_stub"),
    os.path.join(_SIM, "problem-01-reservoir-sampling", "src"),
    os.path.join(_SIM, "problem-03-rl-environment", "src"),
]


def available() -> bool:
    return os.path.isdir(_SIM)


def _prep():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for p in _PATHS:
        if p not in sys.path:
            sys.path.insert(0, p)


def load(module: str, problem_src: str = None):
    """Import reference module `module` (e.g. 'reservoir', 'env', 'rewards')."""
    _prep()
    if problem_src is not None:
        p = os.path.join(_SIM, problem_src, "src")
        if p not in sys.path:
            sys.path.insert(0, p)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):  # env.py prints import warnings
        return importlib.import_module(module)
