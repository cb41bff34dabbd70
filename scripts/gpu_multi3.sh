# gpurun --gpus N -- bash scripts/gpu_multi3.sh N : bench of the cell shard, then of the replicas mode
cd $GRAFT_REPO_ROOT
N=$1
mkdir -p gpurun_out
for mode in "" "--replicas"; do
  tag=n$N$(echo $mode | tr -d '-')
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 $mode > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  tail -2 gpurun_out/bench_$tag.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
    print('$tag value %.3e ms/step %.2f e2e %.3e (%.2f ms) frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e'].get('ms_per_step',0),d['roofline']['frac']))
except Exception as e: print('parse fail',e)
PY
done
