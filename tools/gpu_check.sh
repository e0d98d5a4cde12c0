#!/bin/bash
# Run GPU test groups one by one, each under its own timeout, logging progress to gpurun_out/
# (a hung kernel then costs one group, not the whole call).  Usage: tools/gpu_check.sh [pytest -k expr ...]
mkdir -p gpurun_out
LOG=gpurun_out/gpu_check.log
: > $LOG
groups=("$@")
if [ ${#groups[@]} -eq 0 ]; then
  groups=("embedding or add or silu" "rms_norm" "gemv" "test_gemm" "qk_norm_rope or scatter" "paged_attention_decode" "batch_prefill or planner" "argmax or top1")
fi
for g in "${groups[@]}"; do
  echo "=== ops: $g" | tee -a $LOG
  timeout -k 5 150 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "$g" 2>&1 | tail -25 | tee -a $LOG
  echo "rc=$?" | tee -a $LOG
done
if [ -z "$SKIP_MODEL" ]; then
for t in test_tiny_prefill_decode_parity test_small_config_parity test_long_context_split_kv_path test_batch_matches_sequential test_determinism_and_page_lifecycle test_generate_matches_oracle_free_running; do
  echo "=== model: $t" | tee -a $LOG
  timeout -k 5 200 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "$t" 2>&1 | tail -30 | tee -a $LOG
done
fi
