#!/bin/bash
# GPU session 12 (1 GPU): ncu --set full of one T=2048 launch of the two-tile prefill attention kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attention_tc2 --launch-skip 50 -c 1 -f -o $O/c12_fa2 python tools/bench_prefill_attn.py tc2 > $O/c12_ncu.log 2>&1
tail -3 $O/c12_ncu.log
ls -la $O/c12_fa2.ncu-rep
echo done
