mkdir -p gpurun_out
timeout 150 python tools/check_prefill_tc.py > gpurun_out/tc0.log 2>&1; echo "tc rc=$?"; tail -8 gpurun_out/tc0.log
if grep -q "TC_ATTN_ALL PASS" gpurun_out/tc0.log; then
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -x -q -k "prefill or parity or generate" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:prefill_attention_tc -s 1 -c 1 -o gpurun_out/fa_tc_long python tools/check_prefill_tc.py long > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc.log
fi
