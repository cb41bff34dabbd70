cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --workload tiny_450x375x64_r20 --steps 3 --warmup 3 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; tail -3 gpurun_out/bench_tiny.err; cat gpurun_out/bench_tiny.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_v1.json 2> gpurun_out/bench_v1.err; tail -5 gpurun_out/bench_v1.err; cat gpurun_out/bench_r1_v1.json
