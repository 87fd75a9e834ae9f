# Round 2, call 23 (2 GPUs): the distributed path on the final tree (batched loads in the dots' second pass + all-reduce
# kernel, fused single-rank dots untouched there): parity of rank-locally generated bricks against the serial oracle over
# NVLink peer memory, then bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ALFIB_PEER=1 timeout 300 $TR --master-port 29811 scripts/dist_check_bricks.py ldc3d-sv-k3-wtiny2 > gpurun_out/r2_bricks_n2_peer_c.log 2>&1; el bricks-peer $?; tail -3 gpurun_out/r2_bricks_n2_peer_c.log | cut -c1-300
timeout 600 $TR --master-port 29950 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.log; el bench-n2 $?
grep -v "^\[synth\]\|^\[bricks\]" gpurun_out/r2_bench_n2_c.log | tail -4 | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_c.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction", "scaling")}, d["e2e"]["ms_per_step"], d["config"]["workload"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()}, d["setup_s"])
except Exception as e:
    print("unreadable", e)
PY
el done 0
