cd $GRAFT_REPO_ROOT
cp localexpstereo_b200/liblexp_cuda.so /tmp/orig.so
for x in "$@"; do
  cp gpurun_x$x.so localexpstereo_b200/liblexp_cuda.so
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $x', 'ms/step %.2f'%d['ms_per_step'], d['roofline']['ms_by_layer'])"
done
cp /tmp/orig.so localexpstereo_b200/liblexp_cuda.so
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('baseline', 'ms/step %.2f'%d['ms_per_step'], d['roofline']['ms_by_layer'])"
