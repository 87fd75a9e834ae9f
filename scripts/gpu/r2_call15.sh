# Round 2, call 15 (1 GPU): Burman patch corrections, the continuation tests against the LU fixture, coarse-factor timing
# after the buffer change, the default bench (with continuation.three_d)
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 600 python -m pytest tests/test_burman.py tests/test_continuation.py tests/test_outer.py tests/test_pc_protocol.py -q -m gpu > gpurun_out/r2_t_burman.log 2>&1; el new-tests $?; grep -v Warning gpurun_out/r2_t_burman.log | tail -12
for mode in host device; do
  timeout 300 python scripts/cont_bench.py 2 $mode > gpurun_out/r2_cont_bench_${mode}_b.txt 2>&1; el cont-bench-$mode $?; tail -2 gpurun_out/r2_cont_bench_${mode}_b.txt | cut -c1-560
done
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.log; el bench $?; tail -4 gpurun_out/r2_bench_n1_b.log | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1_b.json") if l.startswith("{")][-1])
    print("ms/cycle %.2f  e2e %.2f  frac %.3f traffic %s setup %s  red %.3e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["setup_s"], d["residual_reduction"]))
    print("breakdown:", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    c = d["continuation"]
    print("cpu:", d["cpu_baseline"] and {k: d["cpu_baseline"][k] for k in ("value", "cores")}, "continuation:", {k: c.get(k) for k in ("time_s", "iteration_parity", "velocity_rel_diff_vs_cpu")})
    print("three_d:", {k: v for k, v in c.get("three_d", {}).items() if k not in ("nonlinear_iter", "linear_iter")})
except Exception as e:
    print("unreadable", e)
PY
el done 0
