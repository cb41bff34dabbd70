# On the GPU box (gpurun -- bash scripts/gpu_variants.sh tma4 r1): A/B the prepared kernel variants built by
# scripts/build_variants.py.  For each: swap the library in, run the fast GPU parity tests, then a short bench; results go to
# gpurun_out/variants.txt.  The shipped library is restored at the end.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp localexpstereo_b200/liblexp_cuda.so /tmp/orig.so
run_bench() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.2f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['roofline']['ms_by_layer'], 'e2e %.3g' % d['e2e']['value'])"
}
{
run_bench baseline
for x in "$@"; do
  cp variants/liblexp_cuda_$x.so localexpstereo_b200/liblexp_cuda.so
  if timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_naive.py -m gpu -x -q 2>&1 | tail -1 | grep -q passed; then
    run_bench "variant $x (parity ok)"
  else
    echo "variant $x: PARITY FAILED"
  fi
done
cp /tmp/orig.so localexpstereo_b200/liblexp_cuda.so
run_bench baseline-again
} 2>&1 | tee gpurun_out/variants.txt
