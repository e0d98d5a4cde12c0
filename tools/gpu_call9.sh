#!/bin/bash
# GPU session 9 (1 GPU): Qwen3.5 fused decode + delta-rule sequence kernel v2 + GEMV x_mode 3 / epi 4, the emulated
# two-rank collectives after the top-1 line fix, config-4 numbers fused vs unfused and seq v1 vs v2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_qwen35_ops_gpu.py tests/test_qwen35_model_gpu.py -x -q -s 2>&1 | tail -15 | tee $O/c9_q35_tests.log
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "two_ranks or fused_prologue" 2>&1 | tail -3
for v in "PK_Q35_FUSED=1 PK_GDR_SEQ=0" "PK_Q35_FUSED=0 PK_GDR_SEQ=0" "PK_Q35_FUSED=1 PK_GDR_SEQ=1"; do
  echo "== config 4: $v"
  env $v timeout 400 python -c "import bench, json; print(json.dumps(bench.config4_leg(0)))" 2>&1 | tail -1 | tee -a $O/c9_config4.jsonl
done
echo done
