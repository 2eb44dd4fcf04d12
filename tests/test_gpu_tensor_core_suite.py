"""The GPU parity suites read TWISTERL_B200_PRECISION (default fp32).  A plain `pytest -m gpu` must also cover the
tensor-core (f16x2, tcgen05) path that bench.py measures, so this test re-runs them in a child process with the
variable set -- the fused persistent pair kernel, its balanced schedule and the MCTS leaf evaluation included."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_gpu_suites_on_the_tensor_core_path():
    if os.environ.get("TWISTERL_B200_PRECISION", "fp32") == "f16x2":
        pytest.skip("already running on the f16x2 path")
    env = dict(os.environ, TWISTERL_B200_PRECISION="f16x2")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_parity.py",
                        "tests/test_gpu_eval.py", "tests/test_gpu_az.py", "tests/test_cpp_host.py"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, (r.stdout[-4000:] + r.stderr[-2000:])
    assert " passed" in r.stdout
