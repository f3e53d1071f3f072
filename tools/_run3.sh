cd $GRAFT_REPO_ROOT
VO_SM_PARTITION=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_fast_score|k_pyr_level|k_pad_level0|k_fast_nms_row" -s 21 -c 7 -f -o gpurun_out/prework_r02 python bench.py --steps 2 --warmup 3 --cpu-seconds 0.3 --sweep 0 > gpurun_out/ncu_prework_r02.out 2>&1
tail -2 gpurun_out/ncu_prework_r02.out | cut -c1-300
ls -la gpurun_out/prework_r02.ncu-rep
