cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/_run_prof.sh 2>&1 | tail -12
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err ) 2>&1 | grep real; tail -c 300 gpurun_out/bench_r2d.err; head -c 200 gpurun_out/bench_r2d.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2d_ref.json 2> gpurun_out/bench_r2d_ref.err ) 2>&1 | grep real; head -c 200 gpurun_out/bench_r2d_ref.json
