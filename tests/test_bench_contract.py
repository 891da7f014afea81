"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the driver reads,
on the same workload naming as our arm (N = 1: Gauss-Seidel, N > 1: the Jacobi workload of config C4), and our arm
refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300, env=e)


def _line(out):
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout + out.stderr
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _line(_run("--impl", "reference", "--size", "12", "--steps", "3", "--warmup", "1"))
    assert d["impl"] == "reference" and d["metric"] == "V-cycle iterations/s" and d["unit"] == "V-cycles/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["value"] > 0 and d["steps"] >= 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "SymmetricGaussSeidel" in d["config"]["workload"]


def test_reference_arm_uses_the_partitioned_workload_for_n_gt_1():
    d = _line(_run("--impl", "reference", "--gpus", "4", "--size", "12", "--steps", "2", "--warmup", "1"))
    assert d["n_gpus"] == 4 and "Jacobi" in d["config"]["workload"]
    # under torchrun only rank 0 runs it
    quiet = _run("--impl", "reference", "--gpus", "4", "--size", "12", "--steps", "2", "--warmup", "1", env={"RANK": "1"})
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    out = _run("--size", "12", "--steps", "2", "--warmup", "1")
    assert out.returncode != 0 and "no CPU fallback" in (out.stdout + out.stderr)
