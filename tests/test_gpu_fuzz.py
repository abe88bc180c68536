"""Randomised sweeps (tools/fuzz_oracle.py, tools/fuzz_paths.py) as part of the GPU suite: random shapes around the
library's switch points, every scorer, sparse / dense / signed / tied inputs — against the CPU oracle and, on larger
shapes, fixed-point path against fp64 path.  The sweeps found the flushed-to-zero bug of wide-dynamic-range columns."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("tool,seed,cases", [("fuzz_oracle.py", 101, 60), ("fuzz_paths.py", 102, 12), ("fuzz_ranks.py", 103, 80)])
def test_randomised_sweep(tool, seed, cases):
    env = dict(os.environ, SEED=str(seed), CASES=str(cases))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
