# Round 2, call 10 (2 GPUs): bench.py --gpus 2 as the driver launches it (weak scaling, distributed vectors, LL exchanges);
# literal 3-D MacroStar GPU tests on GPU 0
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
CUDA_VISIBLE_DEVICES=0 timeout 400 python -m pytest tests/test_gpu_literal_macrostar.py -q -m gpu -x > gpurun_out/r2_t_literal.log 2>&1; el literal $?; tail -4 gpurun_out/r2_t_literal.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29950 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.log; el bench-n2 $?
grep -v "^\[synth\]\|^\[bricks\]" gpurun_out/r2_bench_n2.log | tail -8
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction", "scaling")}, d["e2e"]["ms_per_step"], d["config"]["workload"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()}, d["setup_s"])
except Exception as e:
    print("unreadable", e)
PY
el done 0
