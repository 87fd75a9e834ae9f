python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -15
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_first.json 2> gpurun_out/bench_first.log
tail -5 gpurun_out/bench_first.log; cat gpurun_out/bench_first.json
