#!/bin/bash
# GPU session 4 (1 GPU): one-GPU checks of everything new since session 2, each step bounded and logged separately.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PK_SKIP_FULLSIZE=1
step() { echo "== $1"; shift; timeout 300 "$@" > $O/c4_$STEPNAME.log 2>&1; echo "rc=$? $(tail -2 $O/c4_$STEPNAME.log | tr '\n' ' ')"; }
STEPNAME=attn step "attention v3.1" python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attention"
STEPNAME=tworank step "two-rank emulation (LL all-reduce, top-1 exchange)" python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "two_ranks"
STEPNAME=gemm step "pair GEMM + SwiGLU" python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm"
STEPNAME=model step "model tests (prefill through the pair GEMM)" python -m pytest tests/test_model_gpu.py -m gpu -q -x
STEPNAME=prefill_ops step "prefill op timings" python tools/bench_prefill_ops.py
for slots in 10 20; do
  PK_ATTN_SLOTS=$slots PK_GEMV_STAGES=6 timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK
done | tee $O/c4_ab.log
PK_ATTN_SLOTS=20 timeout 100 python tools/attn_sweep.py 1 2304 2>&1 | grep ATTN | tee -a $O/c4_ab.log
timeout 100 python tools/attn_sweep.py 1 2304 2>&1 | grep ATTN | tee -a $O/c4_ab.log
echo "== TTFT"
timeout 300 python - <<'PY' 2>&1 | tail -3 | tee $O/c4_ttft.log
import os, sys, statistics
sys.path.insert(0, os.getcwd())
from pegainfer_b200.config import QWEN3_4B
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model
from pegainfer_b200.synthetic import iter_random_weights, synthetic_prompt
m = Qwen3Model(QWEN3_4B, iter_random_weights(QWEN3_4B, seed=0, device="cuda"), ModelRuntimeConfig(num_pages=400, max_batch=1))
for n in (2048, 128):
    p = synthetic_prompt(n)
    m.generate(p, 2)
    print("TTFT", n, round(statistics.median([m.generate(p, 1)[1] for _ in range(5)]), 3), "ms")
PY
echo done
