#!/bin/bash
# GPU session 5 (2 GPUs): tensor-parallel correctness (LL GEMV-fused all-reduce, vocab-sharded lm_head) + TP2 A/B.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/c5_topo.txt 2>&1
echo "== TP2 small model + collectives, kernel mode first (round-1 path), then ll (default)"
PK_TP_MODE=kernel timeout 300 $TR --master-port 29612 tests/tools/tp_check.py > $O/c5_tp2_small_kernel.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH" $O/c5_tp2_small_kernel.log | tail -3
timeout 300 $TR --master-port 29611 tests/tools/tp_check.py > $O/c5_tp2_small_ll.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH|Error|error" $O/c5_tp2_small_ll.log | tail -5
echo "== TP2 bench A/B (CUDA-generated weights, tuning only)"
for mode in ll kernel; do
  PK_TP_MODE=$mode timeout 400 $TR --master-port 2962$((RANDOM % 10)) bench.py --gpus 2 --steps 128 --warmup 8 --quick --weights cuda 2>$O/c5_bench_tp2_$mode.err |
    tee $O/c5_bench_tp2_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'ttft', round(d['ttft_ms'],2))"
done
echo "== TP2 Qwen3-8B full size vs the live TP-2 oracle"
timeout 900 $TR --master-port 29613 tests/tools/tp_check.py --model qwen3-8b --prompt 128 --steps 8 > $O/c5_tp2_8b.log 2>&1; grep -E "TP_CHECK|worst" $O/c5_tp2_8b.log | tail -4
echo done
