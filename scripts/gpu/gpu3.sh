set -x
python -m pytest tests -q -m gpu 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.log; tail -3 gpurun_out/bench_r1_a.log; cat gpurun_out/bench_r1_a.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:patch_apply -s 8 -c 3 -o gpurun_out/prof_apply_half python scripts/profile_apply.py ldc3d-sv-k3-half apply 6 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsr_spmv -s 2 -c 2 -o gpurun_out/prof_spmv_half python scripts/profile_apply.py ldc3d-sv-k3-half spmv 6 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_half_cycle.csv python scripts/profile_apply.py ldc3d-sv-k3-half cycle 2 2>&1 | tail -3
ls -la gpurun_out/
