# Round 2, call 17 (1 GPU): the whole GPU suite on the final tree, the default bench (as the driver runs it), smoke()
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; el pytest-gpu $?; grep -v "Warning\|warn\|block_diag\|sparse\|^  *$\|Recover\|Avoid\|For more\|This function\|https\|^$" gpurun_out/r2_pytest_gpu_final.log | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; el smoke $?; tail -2 gpurun_out/r2_smoke.log
ALFIB_PROBLEM_CACHE= timeout 1500 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.log; el bench $?; grep -v "3-D continuation: Re" gpurun_out/r2_bench_n1_final.log | tail -8 | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1_final.json") if l.startswith("{")][-1])
    print("steps %d warmup %d ms/cycle %.2f  e2e %.2f  frac %.3f traffic %s red %.3e launches %d" % (d["steps"], d["warmup"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["residual_reduction"], d["gpu_launches"]))
    print("setup:", d["setup_s"])
    print("breakdown:", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    c = d["continuation"]
    print("cpu:", d["cpu_baseline"] and {k: d["cpu_baseline"][k] for k in ("value", "cores")}, "continuation:", {k: c.get(k) for k in ("time_s", "iteration_parity", "velocity_rel_diff_vs_cpu")})
    print("three_d:", {k: v for k, v in c.get("three_d", {}).items() if k not in ("nonlinear_iter", "linear_iter")})
    print("clocks:", d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
el done 0
