mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | cut -c1-400
timeout 200 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-300
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/r1_v6_prefill_decode_launches.csv python tools/profile_decode.py 2048 3 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
timeout 200 python tools/bench_decode_micro.py > gpurun_out/micro.json 2> gpurun_out/micro.err; echo "micro rc=$?"
timeout 100 python tools/bench_prefill_ops.py > gpurun_out/prefill_ops.log 2>&1; tail -8 gpurun_out/prefill_ops.log
