# Round 2: bench.py --gpus N exactly as the driver launches it (N = $NGPU), without the CPU arm / continuation
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
N=${NGPU:-8}
nproc; free -g | head -2 | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29960 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.log; el bench-n$N $?
grep -v "^\[synth\]\|^\[bricks\]" gpurun_out/r2_bench_n$N.log | tail -12
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction", "scaling")}, d["e2e"]["ms_per_step"], d["config"]["workload"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()}, d["setup_s"])
except Exception as e:
    print("unreadable", e)
PY
el done 0
