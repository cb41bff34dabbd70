cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3; done
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "textureless or single_cell or branches or filter_rect" 2>&1 | tail -1; done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/val2_bench.json 2> gpurun_out/val2_bench.err; tail -2 gpurun_out/val2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/val2_bench.json'))
print('value %.3e ms %.2f frac %.3f | unary sweep %.2f ms | e2e(pm) %.3e (%.2f ms) | e2e unary maps %.3e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('unary_sweep',{}).get('ms_per_step',0), d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d['e2e'].get('unary_maps', {}).get('value', 0)))
PY
