"""bench.py's reference arm (the oracle port on host cores) runs without a GPU: check its JSON line against the
driver's contract, single process and under torchrun (rank 0 prints, the other ranks exit 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check(line, n_gpus):
    assert REQUIRED <= set(line), sorted(REQUIRED - set(line))
    assert line["impl"] == "reference" and line["metric"] == "tps_pp_rectified_img_per_s" and line["unit"] == "img/s"
    assert line["n_gpus"] == n_gpus and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "32"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check(lines[0], 1)
    # --steps / --warmup / --batch are honoured exactly (the driver compares them with our arm: same_config)
    assert lines[0]["steps"] == 2 and lines[0]["warmup"] == 1 and lines[0]["config"]["global_batch"] == 32


def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "1", "--batch", "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check(lines[0], 2)
