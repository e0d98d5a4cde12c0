#!/bin/bash
# GPU session (1 GPU): pair-vs-single GEMM choice after the cost-model calibration: GEMM parity tests, per-op timing, TTFT
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm" 2>&1 | tail -3
timeout 300 python tools/bench_prefill_ops.py 2>&1 | grep -E "T =|gemm|SwiGLU" | tee $O/c18_prefill_ops.log
timeout 600 python bench.py --steps 64 --warmup 8 --quick --weights cuda > $O/c18_bench_quick.json 2> $O/c18_bench_quick.err
python -c "import json; d=json.loads(open('$O/c18_bench_quick.json').read().strip().splitlines()[-1]); print('tok/s', round(d['value'],1), 'ttft', round(d['ttft_ms'],2))"
echo done
