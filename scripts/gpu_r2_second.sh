# round 2, second GPU call: the whole -m gpu suite (new size matrix, texture-less guides, strict naive), default bench,
# ncu --set full of one layer-0 launch, benches of the other BASELINE workloads
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -25 > gpurun_out/r2b_tests.log; tail -6 gpurun_out/r2b_tests.log
timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lexp_fused -s 30 -c 1 -f -o gpurun_out/prof_r2b_L0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_r2b.log 2>&1; tail -2 gpurun_out/ncu_r2b.log
for w in adirondack_shape_1436x992x290_r20 synthetic_4k_3840x2160x512_r32 middv2_cones_shape_450x375x64_naive; do
  extra=""; [ "$w" = synthetic_4k_3840x2160x512_r32 ] && extra="--no-cpu-baseline"
  timeout 900 python bench.py --workload $w --steps 5 $extra > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err; tail -c 600 gpurun_out/r2b_bench_$w.json; tail -2 gpurun_out/r2b_bench_$w.err
done
