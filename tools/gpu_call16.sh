#!/bin/bash
# GPU session (1 GPU): ncu launch list of one 2048-token prefill + 2 decode steps (no CUDA graph), and CUDA-event TTFT split
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/c16_prefill_launches.csv python tools/profile_decode.py 2048 2 1 > $O/c16_ncu.log 2>&1
tail -2 $O/c16_ncu.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/c16_prefill_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0][-60:]
    ns = float(r[-1])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(v[1] for v in agg.values())
print("total us", tot / 1e3, "launches", len(rows))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:62s} n={v[0]:4d} total {v[1]/1e3:9.1f} us  avg {v[1]/v[0]/1e3:8.2f} us")
PY
echo done
