python -m pytest tests -q -m gpu -k "not fullsize" 2>&1 | tail -5
python scripts/parity_table.py ldc2d-sv-k2-tiny ldc2d-pkp0-tiny ldc3d-sv-k3-tiny ldc3d-pkp0-tiny 2>&1 | grep -v Warn | tee gpurun_out/parity_table.txt | tail -12
python scripts/bench_small.py ldc3d-pkp0-small 2>&1 | grep graph=
