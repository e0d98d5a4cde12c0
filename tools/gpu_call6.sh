#!/bin/bash
# GPU session 6 (1 GPU): GEMM (split-K, 320-wide pair tiles, SwiGLU), full suite on the regenerated fixtures, full bench
# line, ncu: decode launch list, GEMV dram bytes per shape, pair-GEMM full sections.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== gemm tests"
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "gemm" > $O/c6_gemm.log 2>&1; tail -3 $O/c6_gemm.log
echo "== prefill ops"
timeout 300 python tools/bench_prefill_ops.py > $O/c6_prefill_ops.log 2>&1; cat $O/c6_prefill_ops.log
echo "== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -s > $O/c6_pytest.log 2>&1
grep -E "fullsize|passed|failed|error" $O/c6_pytest.log | tail -12
echo "== bench (full line)"
timeout 1500 python bench.py > $O/c6_bench.json 2> $O/c6_bench.err
cut -c1-600 $O/c6_bench.json; tail -3 $O/c6_bench.err
echo "== ncu: decode launch list + GEMV dram bytes"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 400 --csv \
  --log-file $O/c6_decode_launches.csv python tools/quick_decode.py --steps 8 --reps 1 > $O/c6_ncu_list.log 2>&1
tail -1 $O/c6_ncu_list.log
echo "== ncu full: pair GEMM (o_proj 320-wide tile, gate_up 256-wide)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 47 -c 1 -o $O/c6_gemm2_o python tools/bench_prefill_ops.py > $O/c6_ncu_g1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 70 -c 1 -o $O/c6_gemm2_gu python tools/bench_prefill_ops.py > $O/c6_ncu_g2.log 2>&1
ls -la $O/*.ncu-rep | tail -3
echo done
