python -m pytest tests -q -m gpu 2>&1 | tail -3
python scripts/bench_small.py ldc2d-sv-k2 ldc2d-pkp0 ldc3d-sv-k3-half 2>&1 | grep graph=
