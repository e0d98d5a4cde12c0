#!/bin/bash
# Experiments queued for the next GPU session (each line prints one result; run under gpurun).
#   bash tools/next_round_sweeps.sh decode     # 1 GPU: GEMV ring / segment sweep on the full decode step
#   bash tools/next_round_sweeps.sh tp 2       # N GPUs: all-reduce protocol (flag vs LL) x variant (kernel vs GEMV-fused)
#   bash tools/next_round_sweeps.sh tp1        # 1 GPU: Qwen3-8B at TP1 (the N=1 point of BASELINE config 3)
#   bash tools/next_round_sweeps.sh qwen35     # 1 GPU: first hardware run of the Qwen3.5 ops + bring-up harness
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
case "${1:-decode}" in
decode)
  for st in 4 5 6; do for kc in 1024 1536 2048; do
    PK_GEMV_STAGES=$st PK_GEMV_KC=$kc timeout 120 python tools/quick_decode.py 2>&1 | grep QUICK
  done; done | tee gpurun_out/sweep_decode.log ;;
tp)
  n=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
  port=29600
  for proto in flag ll; do for fused in 0 1; do
    port=$((port + 1))
    PK_TP_PROTO=$proto PK_TP_FUSED=$fused timeout 150 $TR --master-port $port tests/tools/tp_check.py 2>&1 | grep -E "TP_CHECK|MISMATCH" | sed "s/^/proto=$proto fused=$fused /"
    port=$((port + 1))
    PK_TP_PROTO=$proto PK_TP_FUSED=$fused timeout 200 $TR --master-port $port bench.py --gpus $n --steps 128 --warmup 8 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('proto=$proto fused=$fused tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3))"
  done; done | tee gpurun_out/sweep_tp$n.log ;;
qwen35)
  PK_TEST_QWEN35=1 timeout 300 python -m pytest tests/test_qwen35_ops_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/qwen35_ops.log ;;
tp1)
  timeout 300 python bench.py --model qwen3-8b 2>/dev/null | tee gpurun_out/bench_8b_tp1.json | cut -c1-300 ;;
esac
