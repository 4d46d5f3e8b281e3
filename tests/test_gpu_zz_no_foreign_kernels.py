"""(Named to run last: it needs the CUDA profiling interface, which an outer profiler may hold.)
No kernel but this library's runs under the mirrored API: the flows of
scripts/foreign_kernels.py under torch.profiler -- a path left on an array library's kernels is not a kernel of
this library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_mirrored_flow_launches_only_library_kernels(cuda):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "foreign_kernels.py")], capture_output=True,
                         text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    if "profiler unavailable" in res.stdout:
        pytest.skip("the CUDA profiling interface reports no kernels on this box (another profiler attached?)")
    lines = [l for l in res.stdout.splitlines() if "jmc kernels" in l or "ERROR" in l]
    assert len(lines) >= 10 and not any("ERROR" in l for l in lines), res.stdout[-3000:]
    seen = sum(int(l.split("jmc kernels")[1].split()[0]) for l in lines)
    if seen == 0:       # the flows ran (no ERROR) but the profiler reported no kernel at all: CUPTI is taken or absent
        pytest.skip("torch.profiler saw no CUDA kernels on this box (another profiler attached?)")
    for line in lines:
        assert "foreign 0" in line, line
        assert int(line.split("jmc kernels")[1].split()[0]) > 0, line
    assert "total foreign kernel launches: 0" in res.stdout
