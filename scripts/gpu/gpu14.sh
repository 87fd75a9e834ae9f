N=$1
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/dist_check.py ldc3d-sv-k3-tiny 2>&1 | grep -E "^world|identical|rror" | head -5
if [ "$2" = "bench" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/bench_r1_p$N.json 2> gpurun_out/bench_r1_p$N.log; grep -E "Error|error|timed out" gpurun_out/bench_r1_p$N.log | tail -4; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r1_p$N.json') if l.startswith('{')][0]
print(d['n_gpus'], round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items() if v['ms_per_step']>0}, d['residual_reduction'], d['gpu_launches'])"
fi
