mkdir -p gpurun_out
timeout 150 python tools/check_prefill_tc.py 0 > gpurun_out/tc0.log 2>&1; echo "tc0 rc=$?"; tail -12 gpurun_out/tc0.log
timeout 150 python tools/check_prefill_tc.py 1 small long > gpurun_out/tc1.log 2>&1; echo "tc1 rc=$?"; tail -5 gpurun_out/tc1.log
timeout 200 python tools/bench_decode_micro.py > gpurun_out/micro.json 2> gpurun_out/micro.err; echo "micro rc=$?"; tail -3 gpurun_out/micro.err
