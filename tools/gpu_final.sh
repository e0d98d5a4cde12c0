#!/bin/bash
# Final 1-GPU validation of the round: smoke, the whole GPU test suite, the bench line and the reference arm, as the driver runs them.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee $O/final_pytest.log
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc $?"
python -c "import json; d=json.loads(open('$O/final_bench.json').read().strip().splitlines()[-1]); print('tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ttft', round(d['ttft_ms'],2), 'roofline', round(d['roofline']['frac'],3), 'parity', d['parity'].get('ok'), d['parity'].get('worst_err_ulp_rowmax'), 'clocks', d.get('clocks'), 'cfg4', (d.get('config4') or {}).get('decode_tok_s'))"
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err; tail -c 600 $O/final_bench_ref.json
echo done
