cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --sweep 0 --cpu-seconds 0.3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('value %.0f e2e %.0f lk_ms %.3f single %.3f'%(d['value'],d['e2e']['value'],d['roofline']['avg_launch_ms'],d['roofline']['single_stream_ms_per_step']))"
VO_SM_PARTITION=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.3 --sweep 0 > gpurun_out/launches_bench_r02.out 2>&1
