#!/bin/bash
# GPU session (1 GPU): prefill attention kernels: parity tests, timing per implementation, timing ablations of tc2
# (PK_FA2_DBG bits: 1 no P V MMAs, 2 no S MMAs, 4 no exponentials, 8 no O rescale -- results wrong, timing only).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "prefill_attention_tc" 2>&1 | tail -5
timeout 300 python tools/bench_prefill_attn.py ${IMPLS:-tc2 tc} 2>&1 | tail -12 | tee $O/c13_attn.log
for d in ${ABL:-}; do
  echo "== PK_FA2_DBG=$d"
  PK_T=2048,8192 PK_FA2_DBG=$d timeout 120 python tools/bench_prefill_attn.py tc2 2>&1 | tail -2 | cut -c1-90 | tee -a $O/c13_ablate.log
done
echo done
