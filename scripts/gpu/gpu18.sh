python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -x 2>&1 | tail -3
ALFIB_FACTOR_TIMING=1 python scripts/kernel_bench.py ldc3d-sv-k3-half 10 2>&1 | grep -E "timing|^factor" | tail -2
