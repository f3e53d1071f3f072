cd $GRAFT_REPO_ROOT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r2f_n4.json 2> gpurun_out/bench_r2f_n4.err
tail -c 200 gpurun_out/bench_r2f_n4.err; tail -1 gpurun_out/bench_r2f_n4.json | head -c 300; echo
