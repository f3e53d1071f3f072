cd $GRAFT_REPO_ROOT
# 1. the LK kernel, full capture (4 units x 2000 features, default instantiation and work-item size)
LK_KERNELS=4:8:2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_lk_ring -s 3 -c 1 -f -o gpurun_out/lk_r02 python tools/lk_ab.py 8 2000 1 > gpurun_out/ncu_lk_r02.out 2>&1
tail -1 gpurun_out/ncu_lk_r02.out
# 2. launch list of the bench command (cold-cache serialised times: compare shares)
VO_SM_PARTITION=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.3 --sweep 0 > gpurun_out/launches_bench_r02.out 2>&1
tail -c 300 gpurun_out/launches_bench_r02.out
# 3. launch list of the sequence mode
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_seq_r02.csv python tools/run_seq.py 6 > gpurun_out/launches_seq_r02.out 2>&1
# 4. sanitizers on the new kernels (small cases)
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_lk.py -q -x -k "empty or truncation or random" > gpurun_out/memcheck_lk_r02.txt 2>&1; tail -3 gpurun_out/memcheck_lk_r02.txt
VO_LK_SPAN=1 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_lk.py -q -x -k "empty or truncation" > gpurun_out/racecheck_lk_r02.txt 2>&1; tail -3 gpurun_out/racecheck_lk_r02.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_stages.py -q -x -k "mono" > gpurun_out/memcheck_ess_r02.txt 2>&1; tail -3 gpurun_out/memcheck_ess_r02.txt
