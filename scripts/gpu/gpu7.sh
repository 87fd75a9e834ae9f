timeout 900 ncu --set full --clock-control none --import-source on -k regex:patch_factor_kernel -s 1 -c 1 -f -o gpurun_out/prof_factor_small python scripts/profile_apply.py ldc3d-sv-k3-small factor 1 2>&1 | tail -2
python -m pytest tests/test_continuation.py tests/test_pc_protocol.py -q -m gpu 2>&1 | tail -5
