#!/bin/bash
# GPU session 2 (1 GPU): where do the ~10 us of the fused decode attention go?  cluster-size / ring sweep in the train
# harness, ncu --set full of the kernel, decode-step A/B.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for cs in 1 2 4 8 16; do
  PK_ATTN_CLUSTER=$cs timeout 120 python tools/attn_sweep.py 2>&1 | grep ATTN
done | tee $O/c2_attn_sweep.log
PK_ATTN_CLUSTER=4 PK_ATTN_SLOTS=20 timeout 120 python tools/attn_sweep.py 2304 4096 2>&1 | grep ATTN | tee -a $O/c2_attn_sweep.log
PK_ATTN_CLUSTER=8 PK_ATTN_SLOTS=16 timeout 120 python tools/attn_sweep.py 2304 4096 2>&1 | grep ATTN | tee -a $O/c2_attn_sweep.log
PK_ATTN=cluster timeout 120 python tools/attn_sweep.py 2>&1 | grep ATTN | tee -a $O/c2_attn_sweep.log
echo "== decode step A/B"
for cs in 8 4 2; do
  PK_ATTN_CLUSTER=$cs PK_GEMV_STAGES=6 timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK
done | tee $O/c2_ab.log
PK_ATTN_CLUSTER=4 PK_ATTN_SLOTS=20 PK_GEMV_STAGES=6 timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK | tee -a $O/c2_ab.log
echo "== ncu full: tma attention, cluster 16 and 8"
for cs in 16 8; do
  PK_ATTN_CLUSTER=$cs timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attention_tma -s 150 -c 2 \
    -o $O/c2_attn_cs$cs python tools/quick_decode.py --steps 8 --reps 1 > $O/c2_ncu_cs$cs.log 2>&1
  tail -2 $O/c2_ncu_cs$cs.log
done
echo done
