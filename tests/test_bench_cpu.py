"""bench.py's CPU legs (the `--impl reference` arm and `cpu_baseline`) on the tiny config: they must survive any step
count the driver passes (the oracle's page pool is sized from the request) and print the contract's JSON keys."""
import json
import os
import subprocess
import sys

import bench
from pegainfer_b200.config import PRESETS
from pegainfer_b200.synthetic import random_weights, to_numpy_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpu_decode_rate_at_the_gpu_arms_context():
    """The CPU arm decodes at the GPU arm's context length (pre-filled KV, no CPU prefill), with the OpenMP team size
    set explicitly and read back, three timed repeats."""
    cfg = PRESETS["qwen3-tiny"]
    w = to_numpy_bits(random_weights(cfg, seed=0, device="cpu"))
    rate, runs, n, threads, ctx_end = bench.cpu_decode_rate(cfg, w, 2048, budget_s=3.0)
    assert rate > 0 and len(runs) == 3 and n >= 2 and threads >= 1
    # first-touch step + the untimed steps until the step time is steady (1..64) + the timed steps, all past the filled context
    assert 2048 + 2 + 3 * n <= ctx_end <= 2048 + 1 + 64 + 3 * n


def test_oracle_thread_control_and_row_spreading():
    import numpy as np
    from oracle import qwen3_oracle as O
    assert O.set_num_threads(2) == 2 and O.get_max_threads() == 2
    O.set_num_threads(os.cpu_count() or 1)
    a = np.arange(7 * 33, dtype=np.uint16).reshape(7, 33)
    b = O.spread_rows(a)
    assert b is not a and (a == b).all()


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "3", "--model", "qwen3-tiny"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] >= 1 and len(line["cpu_baseline"]["runs"]) == 3
    assert "OpenMP" in line["cpu_baseline"]["sample"] and "ctx" in line["cpu_baseline"]["sample"]
    # both placements of the timed loop are tried (a child process without torch, and this process) and named
    assert "child process" in line["cpu_baseline"]["sample"] and "this process" in line["cpu_baseline"]["sample"]


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: the driver launches the reference arm under torchrun like the GPU arm; rank 0 alone runs and prints the line,
    the other ranks exit 0 without work (no process group is created)."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_PROC_BIND", "OMP_PLACES")}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "4",
                        "--warmup", "3", "--model", "qwen3-tiny"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{") and '"impl": "reference"' in ln]
    assert len(lines) == 1, r.stdout[-1000:]
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
