python scripts/kernel_bench.py ldc3d-pkp0-mid 20 2>&1 | grep -v "^{" | tail -9
python scripts/kernel_bench.py ldc2d-pkp0 20 2>&1 | grep -v "^{" | tail -9
