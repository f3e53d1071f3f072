cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err ) 2>&1 | grep real; tail -c 300 gpurun_out/bench_r2f.err; head -c 200 gpurun_out/bench_r2f.json; echo
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2f_ref.json 2> gpurun_out/bench_r2f_ref.err ) 2>&1 | grep real; head -c 200 gpurun_out/bench_r2f_ref.json; echo
VO_SM_PARTITION=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.3 --sweep 0 > gpurun_out/launches_bench_r02.out 2>&1
tail -c 200 gpurun_out/launches_bench_r02.out
