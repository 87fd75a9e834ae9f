python -m pytest tests -q -m gpu -x 2>&1 | tail -4
echo "== block spmv"; python scripts/kernel_bench.py ldc3d-sv-k3-half 20 2>&1 | grep -v "^{" | tail -10
echo "== flat spmv"; ALFIB_SPMV_FLAT=1 python scripts/kernel_bench.py ldc3d-sv-k3-half 20 2>&1 | grep -E "^spmv|^smooth"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_apply -s 2 -c 2 -o gpurun_out/prof_apply_half python scripts/profile_apply.py ldc3d-sv-k3-half apply 6 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsr_spmv -s 2 -c 2 -f -o gpurun_out/prof_spmv_half python scripts/profile_apply.py ldc3d-sv-k3-half spmv 6 2>&1 | tail -2
