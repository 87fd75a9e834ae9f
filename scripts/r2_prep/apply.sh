#!/bin/bash
# Apply the round-2 preparation patches on top of main, rebuild libalfib.so for sm_100a and run the CPU suite
# (≈5 min).  Run from the repository root, in THIS container; the built .so then travels with gpurun.
set -e
git am scripts/r2_prep/000*.patch
python -m alfi_b200.build --force
python -m pytest tests -x -q -m "not gpu"
echo "next: gpurun --timeout 1100 -- 'bash scripts/gpu/round2_patched.sh'   (then the 2-GPU lines of scripts/gpu/round2_first.sh)"
