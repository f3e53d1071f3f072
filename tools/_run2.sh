cd $GRAFT_REPO_ROOT
for v in 4 4:8; do
  tag=$(echo $v | tr ':' '_')
  LK_KERNELS=$v timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_lk_ring -s 3 -c 1 -f -o gpurun_out/lk_r2a_$tag python tools/lk_ab.py 4 2000 1 > gpurun_out/ncu_$tag.out 2>&1
  tail -2 gpurun_out/ncu_$tag.out
done
