cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lexp_fused -s 30 -c 2 -f -o gpurun_out/prof_v1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_v1.log
ls -la gpurun_out
