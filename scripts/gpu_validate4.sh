# flakiness check after the stream-ordering fix of the uploads: the non-PM suite 4 times without -x, the PM tests once, smoke, default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_pm.py -p no:cacheprovider 2>&1 | grep -E "FAILED|passed|failed|error" | head -8
done
timeout 600 python -m pytest tests/test_gpu_pm.py -q -m gpu 2>&1 | tail -2
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/val4_bench.json 2> gpurun_out/val4_bench.err; tail -2 gpurun_out/val4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/val4_bench.json'))
print('value %.3e ms %.2f frac %.3f | unary sweep %.2f ms | e2e(pm) %.3e (%.2f ms) | cpu %.3e (%d thr)' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('unary_sweep',{}).get('ms_per_step',0), d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
print(d['roofline']['ms_by_layer'], d['clocks'], d['gpu_launches'])
PY
