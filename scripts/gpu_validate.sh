# one-GPU validation of the tree as the driver will run it: -m gpu suite, smoke, default bench, reference arm; plus the latency probe
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/val_tests.log; cat gpurun_out/val_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; tail -2 gpurun_out/val_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/val_bench.json'))
print('value %.3e ms %.2f frac %.3f | unary sweep %.2f ms | e2e(pm) %.3e (%.2f ms) | e2e unary maps %.3e | cpu %.3e (%d thr)' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('unary_sweep',{}).get('ms_per_step',0), d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d['e2e'].get('unary_maps', {}).get('value', 0), d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
print(d['roofline']['ms_by_layer'], d['clocks'], d['gpu_launches'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 300 python scripts/latency_probe.py 2>&1 | tail -9 | tee gpurun_out/latency_probe.txt
