# shared-block condensed form + tile op v2: validation, bench, variant timing, sanitizer, ncu evidence.
# Ordered by importance; every step has its own timeout so that a late step cannot eat the budget.
mkdir -p gpurun_out
CYCLE_K='regex:tile_ops|sep_rhs|slot_sum|bsr_spmv|csr_apply|dense_gemv|gemv_reduce|finalize_kernel|maxpy|multi_dot|set_rows|scale_kernel|hessenberg|axpby|sub_kernel|patch_apply'
t0=$(date +%s)
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/pytest_r1_shared.log 2>&1; rc=$?
tail -5 gpurun_out/pytest_r1_shared.log; echo "pytest rc=$rc after $(( $(date +%s) - t0 ))s"
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r1_shared.log | head -20
  # which half is at fault: the old path must still pass (it is byte-for-byte the validated code)
  ALFIB_CONDENSE_SHARED=0 ALFIB_TILE_V1=1 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -k "not variants" 2>&1 | tail -3
  ALFIB_CONDENSE_SHARED=0 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -k "not variants" 2>&1 | tail -3
  ALFIB_TILE_V1=1 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -m gpu -k "not variants" 2>&1 | tail -3
  export ALFIB_CONDENSE_SHARED=0 ALFIB_TILE_V1=1
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_r1_shared.json 2> gpurun_out/bench_r1_shared.log; echo "bench rc=$? after $(( $(date +%s) - t0 ))s"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_shared.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_ms'], d['roofline']['algorithmic_bytes_per_launch'], d['setup_s'], d['continuation']['iteration_parity']); print(d['breakdown_ms']); print(d['config']['patch_inverses'], d['gpu_launches'], d['clocks'])"
unset ALFIB_CONDENSE_SHARED ALFIB_TILE_V1
timeout 240 python scripts/variant_bench.py ldc3d-sv-k3 50 > gpurun_out/variant_bench_r1.txt 2>/dev/null; tail -1 gpurun_out/variant_bench_r1.txt | cut -c1-1500; echo "variants after $(( $(date +%s) - t0 ))s"
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize.py 2>&1 | grep -E "ok |ERROR SUMMARY|Error|error:" | head -8 | tee gpurun_out/sanitizer_r1_shared.txt; echo "memcheck after $(( $(date +%s) - t0 ))s"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"tile_ops|sep_rhs|slot_sum" -s 5 -c 5 -o gpurun_out/prof_shared_apply_half python scripts/profile_apply.py ldc3d-sv-k3-half apply 3 2>&1 | tail -2; echo "ncu full after $(( $(date +%s) - t0 ))s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k "$CYCLE_K" -s 2800 -c 1000 --log-file gpurun_out/launches_cycle_r1_shared.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/bench_under_ncu_shared.json 2> gpurun_out/bench_under_ncu_shared.log; echo "ncu launches after $(( $(date +%s) - t0 ))s"; wc -l gpurun_out/launches_cycle_r1_shared.csv
