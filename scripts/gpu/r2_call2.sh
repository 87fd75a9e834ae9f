# Round 2, call 2 (1 GPU): TMA tile ops (parity + timing), setup phases, gated bfs tests, ncu of the apply kernels
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
ALFIB_TILE_TMA=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_coarse_condensed.py -q -m gpu -x > gpurun_out/r2_t_tma.log 2>&1; el tma-tests $?; tail -5 gpurun_out/r2_t_tma.log
timeout 300 python scripts/apply_variants.py ldc3d-sv-k3 100 > gpurun_out/r2_apply_variants.txt 2> gpurun_out/r2_apply_variants.err; el variants $?; grep variant gpurun_out/r2_apply_variants.txt; tail -3 gpurun_out/r2_apply_variants.err
ALFIB_SCHUR_SETUP=1 timeout 300 python scripts/setup_bench.py ldc3d-sv-k3 3 > gpurun_out/r2_setup_schur.txt 2>&1; el setup-schur $?; tail -4 gpurun_out/r2_setup_schur.txt
ALFIB_GPU_PENDING=1 timeout 300 python -m pytest tests/test_gpu_bfs.py -q -m gpu > gpurun_out/r2_t_bfs.log 2>&1; el bfs $?; tail -4 gpurun_out/r2_t_bfs.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tile_ops|sep_rhs|slot_sum" -s 15 -c 5 -o gpurun_out/r2_prof_apply_v2 python scripts/profile_apply.py ldc3d-sv-k3 apply 4 > gpurun_out/r2_ncu_v2.log 2>&1; el ncu-v2 $?; tail -2 gpurun_out/r2_ncu_v2.log
ALFIB_TILE_TMA=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tile_ops|sep_rhs|slot_sum" -s 15 -c 5 -o gpurun_out/r2_prof_apply_tma python scripts/profile_apply.py ldc3d-sv-k3 apply 4 > gpurun_out/r2_ncu_tma.log 2>&1; el ncu-tma $?; tail -2 gpurun_out/r2_ncu_tma.log
el done 0
