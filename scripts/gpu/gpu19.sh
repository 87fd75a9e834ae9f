# condensed patch inverses: first GPU run — edge + parity tests on the small problems, then per-op timings
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python scripts/kernel_bench.py ldc3d-sv-k3-half 20 1 2>&1 | tail -12 | tee gpurun_out/kernel_bench_half_condensed.txt
timeout 300 python scripts/kernel_bench.py ldc3d-sv-k3-half 20 0 2>&1 | tail -12 | tee gpurun_out/kernel_bench_half_dense.txt
