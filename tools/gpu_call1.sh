#!/bin/bash
# Round-2 GPU session 1 (1 GPU): new decode-attention kernel, full-size parity, first hardware run of the Qwen3.5 ops,
# decode A/B + GEMV ring sweep, the full bench line, launch list.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/c1_gpu.txt 2>&1
nproc >> $O/c1_gpu.txt; free -g | head -2 >> $O/c1_gpu.txt

echo "== attention op tests (new TMA kernel)"
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" > $O/c1_attn_tests.log 2>&1
tail -5 $O/c1_attn_tests.log
if ! grep -q " passed" $O/c1_attn_tests.log || grep -q "failed" $O/c1_attn_tests.log; then
  echo "!! TMA attention failing: the rest of the session runs with PK_ATTN=cluster"
  export PK_ATTN=cluster
fi

echo "== Qwen3.5 ops (first hardware run)"
PK_TEST_QWEN35=1 timeout 600 python -m pytest tests/test_qwen35_ops_gpu.py -m gpu -q > $O/c1_q35_tests.log 2>&1
tail -15 $O/c1_q35_tests.log
PK_TEST_QWEN35=1 timeout 600 python tests/tools/qwen35_bringup.py > $O/c1_q35_bringup.log 2>&1
tail -8 $O/c1_q35_bringup.log

echo "== full GPU suite"
timeout 1800 python -m pytest tests -m gpu -q -s > $O/c1_pytest.log 2>&1
grep -E "fullsize|passed|failed|error" $O/c1_pytest.log | tail -20

echo "== decode A/B"
for attn in tma cluster; do
  PK_ATTN=$attn timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK
done | tee $O/c1_ab_attn.log
for slots in 8 16; do
  PK_ATTN_SLOTS=$slots timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK
done | tee -a $O/c1_ab_attn.log
PK_ATTN_CLUSTER=8 timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK | tee -a $O/c1_ab_attn.log
for st in 5 6; do
  PK_GEMV_STAGES=$st timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK
done | tee $O/c1_sweep_gemv.log
PK_GEMV_STAGES=6 PK_GEMV_KC=1536 timeout 300 python tools/quick_decode.py 2>&1 | grep QUICK | tee -a $O/c1_sweep_gemv.log

echo "== bench (full line)"
timeout 1500 python bench.py > $O/c1_bench.json 2> $O/c1_bench.err
cut -c1-1500 $O/c1_bench.json; tail -5 $O/c1_bench.err

echo "== launch list (one decode step region)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file $O/c1_decode_launches.csv \
  python tools/quick_decode.py --steps 8 --reps 1 > $O/c1_ncu_list.log 2>&1
tail -2 $O/c1_ncu_list.log
echo "== micro (config 5)"
timeout 900 python tests/tools/bench_decode_micro.py > $O/c1_micro.json 2> $O/c1_micro.err
tail -3 $O/c1_micro.err
echo done
