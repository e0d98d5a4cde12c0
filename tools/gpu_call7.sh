#!/bin/bash
# GPU session 7 (1 GPU): Qwen3.5 host + HD-256 prefill attention, decode buckets > 4, unified step, then config 4 numbers.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PK_SKIP_FULLSIZE=1
echo "== Qwen3.5 ops + model"
timeout 600 python -m pytest tests/test_qwen35_ops_gpu.py tests/test_qwen35_model_gpu.py -m gpu -q -s > $O/c7_q35.log 2>&1; grep -E "qwen3.5|passed|failed|Error|error" $O/c7_q35.log | tail -12
echo "== model tests (buckets > 4, unified step)"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q > $O/c7_model.log 2>&1; tail -4 $O/c7_model.log
echo "== config 4"
timeout 600 python - <<'PY' 2>&1 | tail -5 | tee $O/c7_config4.log
import json, os, sys
sys.path.insert(0, os.getcwd())
import bench
print(json.dumps(bench.config4_leg(0)))
PY
echo done
