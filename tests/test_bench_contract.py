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


def test_gpu_arm_parameters_come_from_the_product_host_layer():
    """The GPU arm must not execute anything under oracle/: its matrices come from libzosimos_cuda.so's host layer
    (bench.HostParams) and agree with the oracle's to f32 rounding; only the cpu_baseline / reference legs import the oracle."""
    import inspect

    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    assert "oracle" not in inspect.getsource(bench.make_gpu_workload).replace("the oracle is not touched", "")
    assert "oracle" not in inspect.getsource(bench.HostParams).replace("under oracle/", "")
    H = bench.HostParams
    a = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
    b = H.mul3(H.inv3(H.to_xyz("bt709", "D65")), H.to_xyz("bt2020", "D65"))
    assert np.abs(a - b).max() < 1e-6
    W, Hh, ang = 7680, 4320, float(np.deg2rad(30.0))
    m = (O.shift(W / 2, Hh / 2) @ O.rotate(ang) @ O.shift(-W / 2, -Hh / 2))
    assert np.abs(m - H.rotation_about(W / 2, Hh / 2, ang)).max() < 1e-3  # f32 left multiplications vs one f64 product


def test_every_workload_has_one_config_for_both_arms_and_a_cpu_leg():
    """`config` of the JSON line is arm independent (the driver compares the two arms' dicts), every workload of the default run has one,
    and the CPU side (cpu_baseline leg, --impl reference) can run each of them: one bounded oracle sample per workload."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.ALL_WORKLOADS[0] == bench.HEADLINE == "c2_blend"
    for name in bench.ALL_WORKLOADS + ["band_affine"]:
        cfg = bench.workload_config(name)
        assert cfg == bench.workload_config(name) and cfg["workload"] == name and cfg["what"] and cfg["l2"] and cfg["parity"]
        assert name in bench.DEFAULT_FRAMES
    for name in bench.ALL_WORKLOADS:
        r = bench.cpu_baseline_child(name, budget_s=0.01)
        assert r["value"] > 0 and r["kind"] == "port" and r["cores"] >= 1 and r["unit"] == ("fps" if name == "loop_rs" else "MP/s"), (name, r)


def test_reference_arm_honours_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c5_rgb10a2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json_lines(out.stdout)[0]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config("c5_rgb10a2") and line["impl"] == "reference" and line["value"] > 0
