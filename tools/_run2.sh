cd $GRAFT_REPO_ROOT
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2e_n8.json 2> gpurun_out/bench_r2e_n8.err ) 2>&1 | grep real
tail -c 300 gpurun_out/bench_r2e_n8.err; tail -1 gpurun_out/bench_r2e_n8.json | head -c 200; echo
( time VO_BENCH_GATHER=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2e_n8_nogather.json 2> gpurun_out/bench_r2e_n8_nogather.err ) 2>&1 | grep real
tail -1 gpurun_out/bench_r2e_n8_nogather.json | head -c 200; echo
