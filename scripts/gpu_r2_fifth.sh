cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash scripts/gpu_variants.sh tma4 tma3 tma5 > gpurun_out/r2e_variants.log 2>&1; cat gpurun_out/variants.txt; tail -5 gpurun_out/r2e_variants.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k volume_prep 2>&1 | tail -2
