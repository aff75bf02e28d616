"""bench.py's reference arm on CPU: ONE JSON line on stdout with the keys the driver reads; under torchrun only rank 0 prints.
(The GPU arm's line is exercised on the B200 box; this guards the part of the contract that runs anywhere.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    return lines


def check_line(d, n_gpus):
    assert d["impl"] == "reference" and d["metric"] == "so3_reparam_wignerD_fwd_bwd_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"


def test_reference_arm_single_process():
    lines = run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--samples", str(1 << 24)])
    assert len(lines) == 1, lines
    check_line(json.loads(lines[0]), 1)


def test_reference_arm_under_torchrun_prints_once():
    lines = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                 "--master-port", "29577", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"])
    js = [ln for ln in lines if ln.lstrip().startswith("{")]
    assert len(js) == 1, lines
    check_line(json.loads(js[0]), 2)
