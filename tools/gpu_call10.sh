#!/bin/bash
# GPU session 10 (8 GPUs): TP8 small-model + collective check, the TP8 bench line on the seed-0 CPU checkpoint (parity key
# against tests/golden/parity_qwen3-8b_p128_tp8.npz), then TP4 and TP2 bench lines on 4 / 2 of the same box's GPUs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
show() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', 'tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'ttft', round(d['ttft_ms'],2), 'parity', d.get('parity',{}).get('ok'), d.get('parity',{}).get('worst_err_ulp_rowmax'), 'launches', d.get('gpu_launches'))" 2>&1 | tail -1; }
echo "== TP8 small model + collectives"
timeout 300 bash -c "$(declare -f run); run 8 29811 tests/tools/tp_check.py" > $O/c10_tp8_small.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH|sampled|Error" $O/c10_tp8_small.log | tail -6
echo "== TP8 bench (CPU checkpoint, parity key)"
timeout 600 bash -c "$(declare -f run); run 8 29812 bench.py --gpus 8 --steps 256 --warmup 8 --no-cpu-baseline --no-gpu-reference --no-tp-base --no-config4" > $O/c10_bench_tp8.json 2> $O/c10_bench_tp8.err; show $O/c10_bench_tp8.json
echo "== TP4 bench (tuning checkpoint)"
timeout 300 bash -c "$(declare -f run); run 4 29813 bench.py --gpus 4 --steps 256 --warmup 8 --quick --weights cuda" > $O/c10_bench_tp4.json 2> $O/c10_bench_tp4.err; show $O/c10_bench_tp4.json
echo "== TP4 small model"
timeout 200 bash -c "$(declare -f run); run 4 29814 tests/tools/tp_check.py" > $O/c10_tp4_small.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH|sampled|Error" $O/c10_tp4_small.log | tail -4
echo "== TP2 small model (top-1 exchange fix)"
timeout 200 bash -c "$(declare -f run); run 2 29815 tests/tools/tp_check.py" > $O/c10_tp2_small.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH|sampled|Error" $O/c10_tp2_small.log | tail -4
echo done
