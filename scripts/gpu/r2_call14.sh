# Round 2, call 14 (1 GPU): the outer pieces on the device (tests/test_outer.py), the whole GPU suite after the graph /
# colour-check changes, continuation timing with the three outer modes, the 3-D continuation against the LU-solve fixture
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
timeout 500 python -m pytest tests/test_outer.py -q -m gpu -x > gpurun_out/r2_t_outer.log 2>&1; el outer-tests $?; tail -15 gpurun_out/r2_t_outer.log
timeout 700 python -m pytest tests -q -m gpu --deselect tests/test_outer.py > gpurun_out/r2_pytest_gpu_b.log 2>&1; el pytest-gpu $?; tail -6 gpurun_out/r2_pytest_gpu_b.log
for mode in host schur device; do
  timeout 300 python scripts/cont_bench.py 1 $mode > gpurun_out/r2_cont_bench_$mode.txt 2>&1; el cont-bench-$mode $?; tail -1 gpurun_out/r2_cont_bench_$mode.txt | cut -c1-700
done
timeout 1500 python scripts/cont3d.py device ldc3d-sv-k3-small tests/golden/_cont3d_snapshot.npz - host > gpurun_out/r2_cont3d_host.txt 2>&1; el cont3d-host $?; tail -3 gpurun_out/r2_cont3d_host.txt | cut -c1-900
timeout 1500 python scripts/cont3d.py device ldc3d-sv-k3-small tests/golden/_cont3d_snapshot.npz - device > gpurun_out/r2_cont3d_device.txt 2>&1; el cont3d-device $?; tail -3 gpurun_out/r2_cont3d_device.txt | cut -c1-900
el done 0
