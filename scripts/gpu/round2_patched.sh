# Round 2, after `git am scripts/r2_prep/*.patch && python -m alfi_b200.build --force` HERE (the built .so travels):
# 1 GPU, ~12 min.   gpurun --timeout 1100 -- 'bash scripts/gpu/round2_patched.sh'
# Every step has its own timeout (a hung step must not take the box down: strikes) and writes to gpurun_out/.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
# 1. the new GPU tests first (smallest first), then the whole suite
timeout 120 python -m pytest tests/test_gpu_distributed.py -q -m gpu -x > gpurun_out/r2_t_distributed.log 2>&1; el distributed-1rank $?; tail -3 gpurun_out/r2_t_distributed.log
timeout 200 python -m pytest tests/test_gpu_coarse_condensed.py -q -m gpu -x > gpurun_out/r2_t_coarse.log 2>&1; el coarse-condensed $?; tail -3 gpurun_out/r2_t_coarse.log
timeout 300 python -m pytest tests/test_gpu_schur_setup.py -q -m gpu -x -s > gpurun_out/r2_t_schur.log 2>&1; el schur-setup $?; grep -a "per-Newton-step setup\|passed\|failed\|Error" gpurun_out/r2_t_schur.log | tail -14
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_all.log 2>&1; el pytest-all $?; tail -3 gpurun_out/r2_pytest_all.log
# 2. bench: default, then with the Schur-complement setup (per-Newton-step setup and continuation are the numbers to read)
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.log; el bench-default $?
ALFIB_SCHUR_SETUP=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_schur.json 2> gpurun_out/r2_bench_schur.log; el bench-schur $?
python - <<'PY'
import json
for f in ("r2_bench_default", "r2_bench_schur"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/cycle %.2f  e2e %.2f  frac %.2f  setup/Newton %.2fs  continuation %.2fs parity %s  red %.3e" % (
            d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["setup_s"]["per_newton_step"],
            d["continuation"]["time_s"], d["continuation"].get("iteration_parity"), d["residual_reduction"]))
        print("   breakdown:", {k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 200 python -m pytest tests/test_gpu_fused_index.py -q -m gpu > gpurun_out/r2_t_fused.log 2>&1; el fused-index $?; tail -2 gpurun_out/r2_t_fused.log
ALFIB_FUSE_INDEX=1 timeout 200 python scripts/kernel_bench.py ldc3d-sv-k3 2>&1 | tail -2 | sed "s/^/[fused index] /"
# 3. sweep of the X_SS column-chunk width (patch 3) on the finest-level smoother application
for w in 256 96 64; do ALFIB_SPLIT_COLS=$w timeout 200 python scripts/kernel_bench.py ldc3d-sv-k3 2>&1 | tail -2 | sed "s/^/[split $w] /"; done
el done 0
# then, as a separate 2-GPU call (see round2_first.sh for the distributed-vector checks):
#   gpurun --gpus 2 --timeout 900 -- '<the dist_check_halo.py line of round2_first.sh>'
