cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; tail -5 gpurun_out/bench_cur.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cur.json'))
print('value %.3e evals/s  ms/step %.2f  frac %.3f  e2e %.3e  cpu %.3e (%d thr)'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d.get('cpu_baseline',{}).get('value',0),d.get('cpu_baseline',{}).get('cores',0)))
print(d['roofline']['ms_by_layer'], d['clocks'])
PY
