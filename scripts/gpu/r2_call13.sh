# Round 2, call 13 (1 GPU): new GPU tests (multiplicative sweeps, fine-grained drop-in), the reference's baseN-6 size on one
# GPU, the 2-D [P2]^2-P0 configuration at BASELINE size, the largest feasible [P1+FB]^3 member, small-problem setup costs
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 400 python -m pytest tests/test_multiplicative.py tests/test_fieldsplit0_dropin.py tests/test_pc_protocol.py -q -m gpu > gpurun_out/r2_t_new.log 2>&1; el new-tests $?; tail -6 gpurun_out/r2_t_new.log
timeout 200 python scripts/cont_bench.py 2 > gpurun_out/r2_cont_bench.txt 2>&1; el cont-bench $?; tail -2 gpurun_out/r2_cont_bench.txt | cut -c1-900
timeout 300 python bench.py --config ldc2d-pkp0 --steps 20 --warmup 5 --no-continuation > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.log; el bench-cfg2 $?; cut -c1-400 gpurun_out/r2_bench_cfg2.json
timeout 400 python bench.py --config ldc3d-pkp0-mid --steps 10 --warmup 3 --no-continuation --no-cpu-baseline > gpurun_out/r2_bench_pkp0mid.json 2> gpurun_out/r2_bench_pkp0mid.log; el bench-pkp0-mid $?; cut -c1-400 gpurun_out/r2_bench_pkp0mid.json
ALFIB_PROBLEM_CACHE= timeout 1200 python bench.py --config ldc3d-sv-k3-n6 --steps 5 --warmup 3 --no-continuation --no-cpu-baseline > gpurun_out/r2_bench_n6.json 2> gpurun_out/r2_bench_n6.log; el bench-n6 $?; tail -4 gpurun_out/r2_bench_n6.log; cut -c1-400 gpurun_out/r2_bench_n6.json
python - <<'PY'
import json
for f in ("cfg2", "pkp0mid", "n6"):
    try:
        d = json.loads([l for l in open("gpurun_out/r2_bench_%s.json" % f) if l.startswith("{")][-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction")}, "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["setup_s"]["per_newton_step"], d["config"]["velocity_dofs"])
        print("   ", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
el done 0
