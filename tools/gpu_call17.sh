#!/bin/bash
# GPU session (1 GPU): ncu launch list of one 2048-token prefill + 2 decode steps (no CUDA graph, PDL on), and
# ncu --set full of one T=2048 launch of the final prefill attention kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/c17_prefill_decode_launches.csv python tools/profile_decode.py 2048 2 1 > $O/c17_ncu.log 2>&1
tail -1 $O/c17_ncu.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/c17_prefill_decode_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0][-56:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
print("total us", round(tot / 1e3, 1), "launches", len(rows))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{k:58s} n={v[0]:4d} total {v[1]/1e3:9.1f} us  avg {v[1]/v[0]/1e3:8.2f} us  {100*v[1]/tot:5.1f} %")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attention_tc2 --launch-skip 50 -c 1 -f -o $O/c17_fa2_final python tools/bench_prefill_attn.py tc2 > $O/c17_ncu2.log 2>&1
ncu -i $O/c17_fa2_final.ncu-rep --page details > $O/c17_fa2_final_details.txt 2>/dev/null
grep -E "Duration|Elapsed Cycles|Registers Per|Executed Ipc Active|Issue Slots Busy|No Eligible" $O/c17_fa2_final_details.txt | head -8
echo done
