cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; WL=$2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lexp_fused -s 30 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log | cut -c1-200
