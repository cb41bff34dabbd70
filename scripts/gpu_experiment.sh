cd $GRAFT_REPO_ROOT
for x in x1 x2; do
  cp localexpstereo_b200/liblexp_cuda.so /tmp/orig.so
  cp gpurun_$x.so localexpstereo_b200/liblexp_cuda.so
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$x', 'ms/step %.2f'%d['ms_per_step'], d['roofline']['ms_by_layer'])"
  cp /tmp/orig.so localexpstereo_b200/liblexp_cuda.so
done
