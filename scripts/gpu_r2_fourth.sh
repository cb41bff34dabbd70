# round 2, fourth GPU call (1 GPU): PatchMatch phase with the all-PDL launch chain
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pm.py tests/test_gpu_parity.py tests/test_gpu_naive.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2d_tests.log; cat gpurun_out/r2d_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print('value %.3e ms %.2f frac %.3f | unary sweep %.2f ms | e2e(pm) %.3e (%.2f ms) | e2e unary maps %.3e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('unary_sweep',{}).get('ms_per_step',0), d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d['e2e'].get('unary_maps', {}).get('value', 0)))
print(d['roofline']['ms_by_layer'], d['clocks'])
PY
