"""The scripts under examples/ run as they are (each is the numeric core of one of the reference's scripts)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,expect", [
    ("critical_sudakov.py", "max |MC cdf - closed form|"),
    ("softdrop_ecf_fused.py", "fixed coupling: cdf(first edge)"),
    ("shower_correlations.py", "Soft Drop (z_cut = .1) leading ECF"),
    ("reference_imports_unchanged.py", "integrator class: jetmontecarlo_b200.montecarlo.integrator"),
])
def test_example_runs(cuda, tmp_path, script, expect):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert expect in res.stdout, res.stdout[-2000:]
