cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "gather or pipelined" 2>&1 | tail -2
run() { name=$1; n=$2; shift 2; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 20 --warmup 3 --sweep 0 --cpu-seconds 0.3 > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/diag_{n}.json').read().strip().splitlines()[-1]); e=d['e2e']
    print(n, 'value',round(d['value']),'e2e',round(e['value']),'summary',round(e['summary_only']['value']),'blocks',d['timed_blocks_ms']['e2e'], d['parity']['vs_oracle'], d['parity']['gathered_records_ok_rank0'])
except Exception as ex: print(n,'failed',ex); print(open(f'gpurun_out/diag_{n}.err').read()[-600:])
PY
}
run snap_n2 2 A=1
run snap_n2_b 2 A=1
