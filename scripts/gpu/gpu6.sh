python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python scripts/kernel_bench.py ldc3d-sv-k3-half 20 2>&1 | grep -v "^{" | tail -9
timeout 1500 python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.log; tail -2 gpurun_out/bench_r1_b.log; cat gpurun_out/bench_r1_b.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>/dev/null; wc -l gpurun_out/launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:patch_apply -s 2 -c 2 -f -o gpurun_out/prof_apply_full python scripts/profile_apply.py ldc3d-sv-k3 apply 5 2>&1 | tail -2
