cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for part in -4 -8 -12; do
  VO_OPT_SM_PARTITION=$part timeout 300 python bench.py --steps 20 --warmup 3 --sweep 0 --cpu-seconds 0.2 2>gpurun_out/part_$part.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('partition $part','value %.0f e2e %.0f summary %.0f lk_ms %.3f single %.3f parity %s'%(d['value'],d['e2e']['value'],d['e2e']['summary_only']['value'],d['roofline']['avg_launch_ms'],d['roofline']['single_stream_ms_per_step'],d['parity']['vs_oracle']))"
  tail -2 gpurun_out/part_$part.err
done 2>&1 | tee gpurun_out/value_ab_r2f.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 400 gpurun_out/bench_r2b.err; head -c 300 gpurun_out/bench_r2b.json
