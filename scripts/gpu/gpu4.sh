set -x
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py ldc3d-sv-k3-tiny 2>&1 | grep -v Warning | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py ldc2d-pkp0-tiny 2>&1 | grep -v Warning | tail -4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.log; tail -4 gpurun_out/bench_r1_n2.log; cat gpurun_out/bench_r1_n2.json
