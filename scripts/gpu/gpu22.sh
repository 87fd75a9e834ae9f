# rest of the GPU suite + ncu evidence for the condensed apply
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_condensed.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/bench_under_ncu_condensed.json 2> gpurun_out/bench_under_ncu_condensed.log; tail -2 gpurun_out/bench_under_ncu_condensed.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile_ops|sep_rhs" -s 4 -c 4 -o gpurun_out/prof_condensed_apply_full python scripts/profile_apply.py ldc3d-sv-k3 apply 3 2>&1 | tail -3
