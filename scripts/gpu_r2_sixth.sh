cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp localexpstereo_b200/liblexp_cuda.so /tmp/orig.so
cp variants/liblexp_cuda_tma4.so localexpstereo_b200/liblexp_cuda.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lexp_fused -s 30 -c 1 -f -o gpurun_out/prof_r2_tma4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_r2_tma4.log 2>&1; tail -2 gpurun_out/ncu_r2_tma4.log
cp /tmp/orig.so localexpstereo_b200/liblexp_cuda.so
# the unchanged reference loop through the adapter with the new combiner
bash scripts/gpu_dropin_speed.sh > gpurun_out/r2f_dropin.log 2>&1; grep -o '"under_test[^}]*' gpurun_out/dropin_speed.txt | cut -c1-400
