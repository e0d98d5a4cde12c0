#!/bin/bash
# GPU session 8 (N GPUs, N = $1): tensor-parallel check (small model, all-reduce kernels, vocab-sharded greedy token) and a
# bench A/B of the collective modes on the tuning checkpoint; with FULL=1 also the full-size Qwen3-8B check and the real bench line.
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== TP$N small model + collectives (ll default)"
timeout 300 $TR --master-port 29711 tests/tools/tp_check.py > $O/c8_tp${N}_small_ll.log 2>&1; grep -E "TP_CHECK|worst|MISMATCH|rank|Error" $O/c8_tp${N}_small_ll.log | tail -8
echo "== TP$N bench A/B (tuning checkpoint)"
for mode in ll kernel; do
  PK_TP_MODE=$mode timeout 400 $TR --master-port 2972$((RANDOM % 10)) bench.py --gpus $N --steps 128 --warmup 8 --quick --weights cuda 2>$O/c8_bench_tp${N}_$mode.err |
    tee $O/c8_bench_tp${N}_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'ttft', round(d['ttft_ms'],2))"
done
if [ "${FULL:-0}" = "1" ]; then
  echo "== TP$N Qwen3-8B full size vs the live TP-$N oracle"
  timeout 900 $TR --master-port 29713 tests/tools/tp_check.py --model qwen3-8b --prompt 128 --steps 8 > $O/c8_tp${N}_8b.log 2>&1; grep -E "TP_CHECK|worst|rank" $O/c8_tp${N}_8b.log | tail -4
  echo "== TP$N bench line (CPU checkpoint, parity key, tp1 leg)"
  timeout 900 $TR --master-port 29714 bench.py --gpus $N --steps 256 --warmup 8 > $O/c8_bench_tp${N}_full.json 2> $O/c8_bench_tp${N}_full.err
  python -c "import json; d=json.loads(open('$O/c8_bench_tp${N}_full.json').read().strip().splitlines()[-1]); print('full', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ttft', round(d['ttft_ms'],2), 'parity', d['parity'].get('ok'), d['parity'].get('worst_err_ulp_rowmax'), 'tp1', d.get('tp1',{}).get('value'), 'speedup', d.get('speedup_vs_tp1'))"
fi
echo done
