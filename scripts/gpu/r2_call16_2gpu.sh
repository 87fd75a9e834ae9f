# Round 2, call 16 (2 GPUs): the distributed path after fusing the dots' second pass into the small all-reduce — parity of
# rank-locally generated bricks and of the sharded global problem against the serial oracle (NVLink peer-memory transport
# and NCCL), then bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ALFIB_PEER=1 timeout 400 $TR --master-port 29811 scripts/dist_check_bricks.py ldc3d-sv-k3-wtiny2 > gpurun_out/r2_bricks_n2_peer_b.log 2>&1; el bricks-peer $?; tail -3 gpurun_out/r2_bricks_n2_peer_b.log | cut -c1-300
ALFIB_PEER=0 timeout 400 $TR --master-port 29812 scripts/dist_check_bricks.py ldc3d-sv-k3-wtiny2 > gpurun_out/r2_bricks_n2_nccl_b.log 2>&1; el bricks-nccl $?; tail -3 gpurun_out/r2_bricks_n2_nccl_b.log | cut -c1-300
for cfgname in ldc3d-sv-k3-tiny ldc2d-pkp0-tiny; do
  ALFIB_PEER=1 timeout 300 $TR --master-port 29813 scripts/dist_check_halo.py $cfgname > gpurun_out/r2_halo_${cfgname}_n2_b.log 2>&1; el halo-$cfgname $?; tail -3 gpurun_out/r2_halo_${cfgname}_n2_b.log | cut -c1-300
done
timeout 900 $TR --master-port 29950 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.log; el bench-n2 $?
grep -v "^\[synth\]\|^\[bricks\]" gpurun_out/r2_bench_n2_b.log | tail -6 | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_b.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction", "scaling")}, d["e2e"]["ms_per_step"], d["config"]["workload"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()}, d["setup_s"])
except Exception as e:
    print("unreadable", e)
PY
el done 0
