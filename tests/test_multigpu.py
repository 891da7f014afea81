"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): torchrun launches tests/mgpu_worker.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from algebraicmultigrid_jl_b200 import _devlib

    return _devlib.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_fine_level_matches_oracle(amg, world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stderr[-4000:]
    assert "[mgpu] ALL OK" in out.stdout
