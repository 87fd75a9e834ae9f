# Round 2, call 21a (1 GPU): five-level cycle test; cfg5 with Burman stabilisation (dense inverses + patch corrections: the
# reference's own ldc3d job configuration at baseN 4); configs[3] at BASELINE size over five levels (10.33 M dofs)
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
nproc; free -g | head -2 | tail -1
timeout 600 python -m pytest tests/test_gpu_edges.py -q -m gpu > gpurun_out/r2_t_edges.log 2>&1; el edges $?; grep -E "passed|failed|Error" gpurun_out/r2_t_edges.log | tail -3
timeout 1200 python bench.py --config ldc3d-sv-k3-burman --steps 5 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2_bench_burman.json 2> gpurun_out/r2_bench_burman.log; el bench-burman $?; tail -5 gpurun_out/r2_bench_burman.log | cut -c1-300
timeout 1500 python bench.py --config ldc3d-pkp0-l5 --steps 5 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2_bench_pkp0_l5.json 2> gpurun_out/r2_bench_pkp0_l5.log; el bench-l5 $?; tail -5 gpurun_out/r2_bench_pkp0_l5.log | cut -c1-300
python - <<'PY'
import json
for f in ("burman", "pkp0_l5"):
    try:
        d = json.loads([l for l in open("gpurun_out/r2_bench_%s.json" % f) if l.startswith("{")][-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction")}, "e2e", d["e2e"]["ms_per_step"], "roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "avg_ms", "algorithmic_bytes_per_launch")}, d["setup_s"]["per_newton_step"], d["config"]["velocity_dofs"])
        print("   ", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
el done 0
