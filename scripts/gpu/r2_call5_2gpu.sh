# Round 2, call 5 (2 GPUs): where the multi-GPU time goes — per-operation timings at 1 and 2 ranks, three transports
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 1 --master-port 29801 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n1.log 2>&1; el dkb-n1 $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n1.log | cut -c1-400
timeout 400 $TR --nproc-per-node 2 --master-port 29802 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_nccl.log 2>&1; el dkb-n2-nccl $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_nccl.log | cut -c1-400
ALFIB_PEER=1 timeout 400 $TR --nproc-per-node 2 --master-port 29803 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_mbox.log 2>&1; el dkb-n2-mbox $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_mbox.log | cut -c1-400
ALFIB_PEER=1 ALFIB_MBOX_OFF=1 timeout 400 $TR --nproc-per-node 2 --master-port 29804 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_pull.log 2>&1; el dkb-n2-pull $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_pull.log | cut -c1-400
CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/apply_variants.py ldc3d-sv-k3 100 > gpurun_out/r2_apply_variants2.txt 2> gpurun_out/r2_apply_variants2.err; el variants $?; grep variant gpurun_out/r2_apply_variants2.txt | cut -c1-200
el done 0
