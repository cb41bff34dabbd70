# One pass of the -m gpu suite as the driver runs it (-x), smoke, the default bench.  Every command is bounded by a KILL timeout well inside
# gpurun's own limit (a command cut off by that limit counts as a strike); progress goes to files under gpurun_out/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 540 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/val5_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/val5_tests.log
grep -E "gc replay|pm replay|passed|failed|FAILED|Error|rc=|s call" gpurun_out/val5_tests.log | tail -24
timeout -s KILL 150 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout -s KILL 300 python bench.py > gpurun_out/val5_bench.json 2> gpurun_out/val5_bench.err; tail -2 gpurun_out/val5_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/val5_bench.json'))
    print('value %.3e ms %.2f frac %.3f | unary sweep %.2f ms | e2e(pm) %.3e (%.2f ms) | cpu %.3e (%d thr)' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('unary_sweep',{}).get('ms_per_step',0), d['e2e']['value'], d['e2e'].get('ms_per_step', 0), d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
    print(d['roofline']['ms_by_layer'], d['clocks'], d['gpu_launches'])
except Exception as e:
    print('bench line unreadable:', e)
PY
