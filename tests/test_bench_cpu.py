"""bench.py's CPU arm (`--impl reference`: the oracle's C port, no GPU needed) and the keys both arms must share."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = "0.01"


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def _check_line(line, world):
    sys.path.insert(0, ROOT)
    import bench
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["n_gpus"] == world and line["higher_is_better"] is True
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["vs_baseline"] is None
    # the GPU arm prints workload_config(world, scale) too: the driver compares the two `config` objects key by key
    assert line["config"] == json.loads(json.dumps(bench.workload_config(world, float(SCALE))))
    assert "model" not in line["config"] and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "write_depth" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["ms_per_step"] > 0


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scale", SCALE], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_line(lines[0], 1)


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: rank 0 alone runs and prints the CPU arm, the other ranks leave with exit code 0 and no output line"""
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--scale", SCALE],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert r.returncode == 0, r.stdout + r.stderr
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_line(lines[0], 2)
