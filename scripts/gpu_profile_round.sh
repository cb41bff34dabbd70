cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r1}
# launch list: every kernel of one bench step with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
tail -1 gpurun_out/launches_$TAG.log | cut -c1-150
# full capture of the dominant kernel: a layer-0 launch and a layer-2 launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lexp_fused -s 30 -c 1 -f -o gpurun_out/prof_${TAG}_L0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_L0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lexp_fused -s 230 -c 1 -f -o gpurun_out/prof_${TAG}_L2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_L2.log 2>&1
ls -la gpurun_out | tail -8
