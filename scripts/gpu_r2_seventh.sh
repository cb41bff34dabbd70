cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash scripts/gpu_variants.sh tma4 tma3 > gpurun_out/r2g_variants.log 2>&1; cat gpurun_out/variants.txt; tail -3 gpurun_out/r2g_variants.log
T=$(nproc)
timeout 600 ./oracle/_ref/dropin_check --threads $T --W 1024 --H 768 --D 64 --K 3 2>&1 | tee gpurun_out/r2g_dropin.txt | cut -c1-600
