cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
echo rc=$?; tail -c 1500 gpurun_out/bench_r2_n2.err; head -c 600 gpurun_out/bench_r2_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/bench_r2_ref_n2.json 2> gpurun_out/bench_r2_ref_n2.err
echo rc=$?; head -c 300 gpurun_out/bench_r2_ref_n2.json
