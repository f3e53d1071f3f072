cd $GRAFT_REPO_ROOT
for cfg in "8 2 0" "6 2 0" "5 2 0" "7 2 0"; do
  set -- $cfg
  VO_OPT_LK_CTAS_PER_SM=$1 VO_OPT_LK_SPAN=$2 VO_OPT_LK_QUOTA=$3 timeout 300 python bench.py --steps 20 --warmup 3 --sweep 0 --cpu-seconds 0.2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cps/span/quota $cfg','value %.0f e2e %.0f summary %.0f lk_ms %.3f single %.3f'%(d['value'],d['e2e']['value'],d['e2e']['summary_only']['value'],d['roofline']['avg_launch_ms'],d['roofline']['single_stream_ms_per_step']))"
done 2>&1 | tee gpurun_out/value_ab_r2b.txt
bash tools/_run_prof.sh
