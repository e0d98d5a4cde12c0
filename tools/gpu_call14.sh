#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "prefill_attention_tc and tc2" 2>&1 | tail -3
timeout 300 python tools/bench_prefill_attn.py tc2 2>&1 | tail -4
for d in ${TR:-32}; do
  echo "== PK_FA2_DBG=$d"
  PK_FA2_DBG=$d timeout 120 python tools/fa2_trace.py 2>&1 | tail -9 | tee -a gpurun_out/c14_trace_$d.log
done
