# Round 2, call 22 (1 GPU): the final tree — default bench (incl. both 3-D continuation ladders: unstabilised to Re 2900,
# Burman-stabilised to Re 5000), the GPU suite, A/B of the fused FGMRES dots, compute-sanitizer memcheck of the round-2 paths
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final2.json 2> gpurun_out/r2_bench_n1_final2.log; el bench $?
grep -v "3-D continuation: Re" gpurun_out/r2_bench_n1_final2.log | tail -4 | cut -c1-250
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1_final2.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction")}, "e2e", d["e2e"]["ms_per_step"],
          "roofline", {k: d["roofline"][k] for k in ("frac", "avg_ms")}, "setup", d["setup_s"]["per_newton_step"])
    print("   ", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    c = d["continuation"]
    print("    2-D parity", c.get("iteration_parity"), c.get("velocity_rel_diff_vs_cpu"))
    for k in ("three_d", "three_d_burman"):
        t = c.get(k) or {}
        print("   ", k, {a: b for a, b in t.items() if not isinstance(b, list)})
except Exception as e:
    print("bench line unreadable", e)
PY
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_final2.log 2>&1; el pytest $?; tail -3 gpurun_out/r2_pytest_gpu_final2.log
for f in 1 0; do
  ALFIB_FUSE_DOTS=$f timeout 300 python scripts/kernel_bench.py ldc3d-sv-k3 10 > gpurun_out/r2_kernel_bench_fuse$f.txt 2>&1; el kb-fuse$f $?
  grep -E "^(smooth|cycle|apply|spmv)" gpurun_out/r2_kernel_bench_fuse$f.txt
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; el memcheck $?
grep -E " ok |ERROR SUMMARY|Error|error:" gpurun_out/r2_sanitizer_memcheck.txt | head -14
el done 0
