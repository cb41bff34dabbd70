# round 2, first GPU call: trace, variant A/B, drop-in speed (tests come with the variants' parity runs)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash scripts/gpu_trace.sh trace > gpurun_out/r2_trace.log 2>&1
bash scripts/gpu_variants.sh occ3 pdl rowtab occ3pdl occ3rowtab linkstr kg4 crolling hreread > gpurun_out/r2_variants.log 2>&1
bash scripts/gpu_dropin_speed.sh > gpurun_out/r2_dropin.log 2>&1
tail -30 gpurun_out/r2_trace.log; cat gpurun_out/variants.txt; tail -40 gpurun_out/r2_dropin.log
