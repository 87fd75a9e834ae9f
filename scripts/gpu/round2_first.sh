# First GPU call of the next round: everything that was written after round 1's GPU budget was
# spent, plus the evidence that is missing.  1 GPU; ~8 min.   gpurun --timeout 900 -- 'bash scripts/gpu/round2_first.sh'
mkdir -p gpurun_out
t0=$(date +%s)
# 1. the whole GPU suite on the final round-1 tree (SV configs now sweep in the reference's "0+:1-" order)
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/pytest_r2_first.log 2>&1; echo "pytest rc=$? after $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/pytest_r2_first.log
# 2. the gated bfs2d (BASELINE configs[2]) GPU tests
ALFIB_GPU_PENDING=1 timeout 200 python -m pytest tests/test_gpu_bfs.py -q -m gpu 2>&1 | tail -3
# 3. full-size --set full capture of the five shared-block apply launches -> roofline.traffic
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tile_ops|sep_rhs|slot_sum" -s 5 -c 5 \
  -o gpurun_out/prof_shared_apply_full python scripts/profile_apply.py ldc3d-sv-k3 apply 3 2>&1 | tail -2
echo "ncu after $(( $(date +%s) - t0 ))s"
# 4. bench (default) for the record
timeout 400 python bench.py > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.log; echo "bench rc=$? after $(( $(date +%s) - t0 ))s"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_first.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['setup_s'], d['continuation']['time_s'], d['continuation']['iteration_parity'])"
# then, as a separate 2-GPU call:  gpurun --gpus 2 --timeout 600 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_2gpu.json; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py'
# after `git am scripts/r2_prep/*.patch` + build, 2 GPUs (distributed level vectors):
#   gpurun --gpus 2 --timeout 900 -- 'for c in ldc3d-sv-k3-tiny ldc2d-pkp0-tiny ldc3d-pkp0-tiny; do timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_check_halo.py $c; done; ALFIB_PEER=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/dist_check_halo.py ldc3d-sv-k3-tiny; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/dist_check_halo.py ldc3d-sv-k3 --time'
# weak scaling (r2_prep/0007), 2 GPUs first:  gpurun --gpus 2 --timeout 1200 -- 'timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 --scaling weak --no-cpu-baseline > gpurun_out/bench_r2_weak2.json 2> gpurun_out/bench_r2_weak2.log; tail -3 gpurun_out/bench_r2_weak2.log'
# rank-locally generated problem on 2 GPUs vs the serial oracle (r2_prep/0008), before the weak bench:
#   gpurun --gpus 2 --timeout 600 -- 'timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/dist_check_bricks.py; ALFIB_PEER=1 timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/dist_check_bricks.py'
