"""bench.py's reference arm runs on host cores only, so its JSON contract is checked here without a GPU: single process and
under torchrun with two ranks (rank 0 alone prints the line, the other rank exits 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def check(line, n):
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["n_gpus"] == n and line["steps"] == 1 and line["value"] > 0
    assert line["metric"] == "megapixels/sec" and line["unit"] == "MP/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == "c2_blend"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_single_process():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = json_lines(out.stdout)
    assert len(lines) == 1
    check(lines[0], 1)


def test_reference_arm_under_torchrun():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = json_lines(out.stdout)
    assert len(lines) == 1  # rank 0 only
    check(lines[0], 2)
