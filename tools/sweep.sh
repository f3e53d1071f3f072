#!/bin/bash
# tools/sweep.sh -- BASELINE.json configs 3 and 5 on one GPU: feature-count sweep at 1241x376 and the 1920x1080 / 4000-feature case.
mkdir -p gpurun_out
for n in 500 1000 2000 4000 8000; do
  timeout 300 python bench.py --features $n --steps 40 --warmup 3 --cpu-sample 2 --sequence 0 2>/dev/null | tail -1 > gpurun_out/sweep_n$n.json
done
timeout 400 python bench.py --width 1920 --height 1080 --calib zed --features 4000 --units 4 --steps 20 --warmup 3 --cpu-sample 2 --sequence 0 2>/dev/null | tail -1 > gpurun_out/sweep_1080p_n4000.json
python - <<'PY'
import json,glob
rows=[]
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f,'unreadable',e); continue
    r=d['roofline']
    rows.append((f.split('/')[-1], d['config']['features'], d['config']['units_per_gpu'], round(d['value'],1), round(d['e2e']['value'],1), round(r['avg_launch_ms'],3), round(r['achieved'],1), round(r['frac'],4), round(d['cpu_baseline']['value'],1)))
print("file features units value_fps e2e_fps lk_ms lk_GBps frac cpu_fps")
for r in rows: print(*r)
PY
