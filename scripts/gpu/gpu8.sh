python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python scripts/kernel_bench.py ldc3d-sv-k3-half 10 2>&1 | grep -E "^setup|^factor|^apply|^cycle"
