#!/bin/bash
# GPU session 3 (2 GPUs): tensor-parallel correctness (LL GEMV-fused all-reduce, vocab-sharded lm_head) + TP2 A/B,
# attention v3.1 re-check.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== one-GPU: attention v3.1 + two-rank emulation"
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or two_ranks" > $O/c3_ops.log 2>&1; tail -4 $O/c3_ops.log
echo "== TP2 small model + collectives (default ll mode), then kernel mode"
timeout 600 $TR --master-port 29611 tests/tools/tp_check.py > $O/c3_tp2_small_ll.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH|Error|error" $O/c3_tp2_small_ll.log | tail -5
PK_TP_MODE=kernel timeout 600 $TR --master-port 29612 tests/tools/tp_check.py > $O/c3_tp2_small_kernel.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH" $O/c3_tp2_small_kernel.log | tail -3
echo "== TP2 Qwen3-8B full size"
timeout 1500 $TR --master-port 29613 tests/tools/tp_check.py --model qwen3-8b --prompt 128 --steps 8 > $O/c3_tp2_8b.log 2>&1; grep -E "TP_CHECK|worst|step" $O/c3_tp2_8b.log | tail -12
echo "== TP2 bench A/B (CUDA-generated weights, tuning only)"
for mode in ll kernel; do
  PK_TP_MODE=$mode timeout 600 $TR --master-port 2962$((RANDOM % 10)) bench.py --gpus 2 --steps 128 --warmup 8 --quick --weights cuda 2>$O/c3_bench_tp2_$mode.err |
    tee $O/c3_bench_tp2_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'ttft', round(d['ttft_ms'],2))"
done
echo "== one-GPU decode A/B: attention v3.1"
for slots in 10 20; do
  PK_ATTN_SLOTS=$slots PK_GEMV_STAGES=6 timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK
done | tee $O/c3_ab.log
PK_ATTN_SLOTS=20 timeout 100 python tools/attn_sweep.py 1 2304 2>&1 | grep ATTN | tee -a $O/c3_ab.log
echo done
