# Round 2, call 24 (1 GPU): the bench line of the final tree (state verdict of the 3-D ladders with the CPU-vs-CPU floor of
# the fixtures) and the ncu launch list of one steady-state cycle after the FGMRES second passes moved into their consumers
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final3.json 2> gpurun_out/r2_bench_n1_final3.log; el bench $?
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1_final3.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction")}, "e2e", d["e2e"]["ms_per_step"],
          "roofline", {k: d["roofline"][k] for k in ("frac", "avg_ms")}, "setup", d["setup_s"]["per_newton_step"], "clocks", d["clocks"])
    print("   ", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    c = d["continuation"]
    print("    2-D parity", c.get("iteration_parity"), c.get("velocity_rel_diff_vs_cpu"), c.get("pressure_rel_diff_vs_cpu"))
    for k in ("three_d", "three_d_burman"):
        t = c.get(k) or {}
        print("   ", k, {a: b for a, b in t.items() if not isinstance(b, list)})
except Exception as e:
    print("bench line unreadable", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 858 --csv --log-file gpurun_out/r2_launches_cycle_b.csv python scripts/profile_apply.py ldc3d-sv-k3 cycle 5 > gpurun_out/r2_ncu_launches_b.log 2>&1; el ncu $?
tail -2 gpurun_out/r2_ncu_launches_b.log | cut -c1-200
el done 0
