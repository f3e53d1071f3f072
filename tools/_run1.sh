cd $GRAFT_REPO_ROOT
# the round-end sequence on one GPU: GPU tests, smoke, the bench line, the reference arm
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 ${BENCH_FLAGS:-} > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err ) 2>&1 | grep real; tail -c 300 gpurun_out/bench_check.err; head -c 200 gpurun_out/bench_check.json; echo
if [ -z "${SKIP_REF:-}" ]; then ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_check_ref.json 2> gpurun_out/bench_check_ref.err ) 2>&1 | grep real; head -c 200 gpurun_out/bench_check_ref.json; echo; fi
