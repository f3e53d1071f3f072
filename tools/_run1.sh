cd $GRAFT_REPO_ROOT
VO_LK_SPAN=1 timeout 900 python -m pytest tests/test_gpu_lk.py -x -q 2>&1 | tail -5
VO_LK_SPAN=3 timeout 900 python -m pytest tests/test_gpu_lk.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_lk.py tests/test_gpu_path.py tests/test_gpu_seq.py -x -q 2>&1 | tail -3
timeout 300 python tools/lk_ab.py 8 2000 10 2>&1 | tail -1 | tee gpurun_out/lk_ab_r2b.json
