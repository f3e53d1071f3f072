cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for cfg in "1 0" "1 1" "0 1" "0 0"; do
  set -- $cfg
  VO_OPT_PRIORITIES=$1 VO_OPT_BATCH_GRAPHS=$2 timeout 300 python bench.py --steps 20 --warmup 3 --sweep 0 --cpu-seconds 0.2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('prio/graphs $cfg','value %.0f e2e %.0f summary %.0f lk_ms %.3f single %.3f launches %d'%(d['value'],d['e2e']['value'],d['e2e']['summary_only']['value'],d['roofline']['avg_launch_ms'],d['roofline']['single_stream_ms_per_step'],d['gpu_launches']))"
done 2>&1 | tee gpurun_out/value_ab_r2c.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pre_r02.csv python tools/run_batch.py 8 2 > /dev/null 2>&1
