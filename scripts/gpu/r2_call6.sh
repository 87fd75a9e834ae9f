# Round 2, call 6 (2 GPUs): TMA tile-op versions; exchange cost experiments (skip exchange / skip kernel)
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
CUDA_VISIBLE_DEVICES=0 ALFIB_TILE_TMA=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_coarse_condensed.py -q -m gpu -x > gpurun_out/r2_t_tma2.log 2>&1; el tma-tests $?; tail -3 gpurun_out/r2_t_tma2.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/apply_variants.py ldc3d-sv-k3 100 > gpurun_out/r2_apply_variants3.txt 2> gpurun_out/r2_apply_variants3.err; el variants $?; grep variant gpurun_out/r2_apply_variants3.txt | cut -c1-220; tail -2 gpurun_out/r2_apply_variants3.err
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ALFIB_PEER=1 ALFIB_DEBUG_SKIP_EXCHANGE=1 timeout 400 $TR --nproc-per-node 2 --master-port 29811 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_noexch.log 2>&1; el dkb-n2-noexch $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_noexch.log | cut -c1-300
ALFIB_PEER=1 ALFIB_DEBUG_SKIP_SPMV=1 timeout 400 $TR --nproc-per-node 2 --master-port 29812 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_nospmv.log 2>&1; el dkb-n2-nospmv $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_nospmv.log | cut -c1-300
ALFIB_DEBUG_SKIP_SPMV=1 timeout 400 $TR --nproc-per-node 2 --master-port 29813 scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_nospmv_nccl.log 2>&1; el dkb-n2-nospmv-nccl $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_nospmv_nccl.log | cut -c1-300
el done 0
