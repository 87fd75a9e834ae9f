# Round 2, call 7 (1 GPU): TMA tile-op configurations (warps x stage size, RED scatter)
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 400 python scripts/apply_variants.py ldc3d-sv-k3 100 > gpurun_out/r2_apply_variants4.txt 2> gpurun_out/r2_apply_variants4.err; el variants $?; grep variant gpurun_out/r2_apply_variants4.txt | cut -c1-220; tail -2 gpurun_out/r2_apply_variants4.err
for cfg in 1 2 3; do ALFIB_TILE_TMA=1 ALFIB_TILE_TMA_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_coarse_condensed.py -q -m gpu -x > gpurun_out/r2_t_tma_cfg$cfg.log 2>&1; el tma-tests-cfg$cfg $?; tail -2 gpurun_out/r2_t_tma_cfg$cfg.log; done
el done 0
