#!/bin/bash
# GPU session (1 GPU): model-level validation after the prefill attention rewrite + measured op errors + TTFT
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -s -k "prefill_attention_tc" 2>&1 | grep -E "max err|passed|failed" | tee $O/c15_attn_err.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 128 --warmup 8 --quick > $O/c15_bench_quick.json 2> $O/c15_bench_quick.err
python -c "import json; d=json.loads(open('$O/c15_bench_quick.json').read().strip().splitlines()[-1]); print('tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ttft', round(d['ttft_ms'],2), 'parity', d.get('parity',{}).get('ok'), d.get('parity',{}).get('worst_err_ulp_rowmax'))"
echo done
