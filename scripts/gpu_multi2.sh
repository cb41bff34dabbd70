# gpurun --gpus N -- bash scripts/gpu_multi2.sh N : multi-GPU correctness (peer-memory cell shard == single-GPU sweep), then the bench
cd $GRAFT_REPO_ROOT
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_n$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py > gpurun_out/multi_check_n$N.log 2>&1; grep MULTI_GPU_CHECK gpurun_out/multi_check_n$N.log || tail -15 gpurun_out/multi_check_n$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -4 gpurun_out/bench_n$N.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N value %.3e ms/step %.2f e2e %.3e (%.2f ms) frac %.3f graph %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e'].get('ms_per_step',0),d['roofline']['frac'],d['config']['cuda_graph']))
except Exception as e: print('parse fail',e)
PY
