#!/bin/bash
# GPU session 11 (1 GPU): the two-tile tcgen05 prefill attention (tc2) against the oracle, then its timing against tc.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "prefill_attention_tc or batch_prefill_paged or top1" 2>&1 | tail -25 | tee $O/c11_tests.log
timeout 300 python tools/bench_prefill_attn.py 2>&1 | tail -14 | tee $O/c11_attn.log
echo done
