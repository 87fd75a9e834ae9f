python -m pytest tests -q -m gpu 2>&1 | tail -3
ALFIB_FACTOR_TIMING=1 python scripts/kernel_bench.py ldc3d-sv-k3-half 10 2>&1 | grep -E "timing|^factor" | tail -2
ALFIB_FACTOR_TIMING=1 timeout 1500 python bench.py > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.log; grep -E "timing|setup" gpurun_out/bench_r1_c.log | tail -8; cat gpurun_out/bench_r1_c.json
