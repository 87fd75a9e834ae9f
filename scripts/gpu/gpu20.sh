timeout 300 python scripts/debug_condensed.py 3 6 2>&1 | tail -30
timeout 300 python scripts/debug_condensed.py 2 4 2>&1 | tail -8
timeout 600 compute-sanitizer --tool initcheck --print-limit 8 python scripts/debug_condensed.py 3 2 2>&1 | grep -v "^colours\|^run\|^   dof" | head -60
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python scripts/debug_condensed.py 3 2 2>&1 | grep -E "=====|Invalid|at |by " | head -30
