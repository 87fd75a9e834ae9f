# Round 2, call 9 (1 GPU): the whole GPU suite on the new defaults (TMA tile ops, Schur setup), default bench,
# launch list of one cycle, full ncu capture of the finest-level apply (-> roofline.traffic), per-op timings
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; el pytest-gpu $?; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.log; el bench $?; tail -5 gpurun_out/r2_bench_n1.log
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][-1])
    print("ms/cycle %.2f  e2e %.2f  frac %.3f  setup %s  red %.3e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["setup_s"], d["residual_reduction"]))
    print("breakdown:", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    print("cpu:", d["cpu_baseline"] and {k: d["cpu_baseline"][k] for k in ("value", "cores")}, "continuation:", {k: d["continuation"].get(k) for k in ("time_s", "iteration_parity", "velocity_rel_diff_vs_cpu")})
except Exception as e:
    print("unreadable", e)
PY
timeout 300 python scripts/kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_kernel_bench.txt 2>&1; el kernel-bench $?; grep -v "^\[synth" gpurun_out/r2_kernel_bench.txt | head -12
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1400 --csv --log-file gpurun_out/r2_launches_cycle.csv python scripts/profile_apply.py ldc3d-sv-k3 cycle 4 > gpurun_out/r2_ncu_launches.log 2>&1; el ncu-launches $?; tail -1 gpurun_out/r2_ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tile_ops|sep_rhs|slot_sum" -s 15 -c 5 -o gpurun_out/r2_prof_apply_tma_final python scripts/profile_apply.py ldc3d-sv-k3 apply 4 > gpurun_out/r2_ncu_tma_final.log 2>&1; el ncu-apply $?; tail -1 gpurun_out/r2_ncu_tma_final.log
el done 0
