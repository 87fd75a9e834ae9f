N=$1
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/dist_check.py ldc3d-sv-k3-tiny > gpurun_out/dist_check_n$N.log 2>&1; grep -E "^world|identical" gpurun_out/dist_check_n$N.log; grep -i -E "nvls|Connected all" gpurun_out/dist_check_n$N.log | head -6
export NCCL_DEBUG=WARN
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.log; grep -E "Error|error" gpurun_out/bench_r1_n$N.log | tail -4; cat gpurun_out/bench_r1_n$N.json | cut -c1-400
