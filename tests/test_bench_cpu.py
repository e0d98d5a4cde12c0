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


def test_cpu_decode_rate_outlives_one_page_pool():
    cfg = PRESETS["qwen3-tiny"]
    w = to_numpy_bits(random_weights(cfg, seed=0, device="cpu"))
    rate, n, dt = bench.cpu_decode_rate(cfg, w, 300, budget_s=30.0)  # 300 steps >> the old fixed 8-page pool
    assert n == 300 and rate > 0 and dt > 0


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "3", "--model", "qwen3-tiny"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
