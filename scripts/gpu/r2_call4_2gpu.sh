# Round 2, call 4 (2 GPUs): mailbox (push) transport of the distributed-vector exchanges
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
p=29700
for c in ldc3d-sv-k3-tiny ldc2d-pkp0-tiny ldc3d-pkp0-tiny; do
  p=$((p+1)); ALFIB_PEER=1 timeout 240 $TR --master-port $p scripts/dist_check_halo.py $c > gpurun_out/r2_mbox_${c}_n$N.log 2>&1; el mbox-$c $?
  grep "rel diff\|Error\|error" gpurun_out/r2_mbox_${c}_n$N.log | tail -9
done
p=$((p+1)); ALFIB_PEER=1 ALFIB_MBOX_OFF=1 timeout 240 $TR --master-port $p scripts/dist_check_halo.py ldc3d-sv-k3-tiny > gpurun_out/r2_pull_tiny_n$N.log 2>&1; el pull-tiny $?
grep "rel diff\|Error\|error" gpurun_out/r2_pull_tiny_n$N.log | tail -9
p=$((p+1)); ALFIB_PEER=1 timeout 300 $TR --master-port $p scripts/dist_check_bricks.py > gpurun_out/r2_mbox_bricks_n$N.log 2>&1; el mbox-bricks $?
grep "world" gpurun_out/r2_mbox_bricks_n$N.log | tail -3
p=$((p+1)); ALFIB_PEER=1 timeout 600 $TR --master-port $p scripts/dist_check_halo.py ldc3d-sv-k3 --time > gpurun_out/r2_mbox_cfg5_n$N.log 2>&1; el mbox-cfg5 $?
grep "world\|events\|Error" gpurun_out/r2_mbox_cfg5_n$N.log | tail -6
p=$((p+1)); timeout 900 $TR --master-port $p bench.py --gpus $N --steps 5 --warmup 3 --scaling weak --peer-memory 1 --no-cpu-baseline > gpurun_out/r2_bench_weak_mbox_n$N.json 2> gpurun_out/r2_bench_weak_mbox_n$N.log; el bench-weak-mbox $?
tail -3 gpurun_out/r2_bench_weak_mbox_n$N.log; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_weak_mbox_n$N.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "residual_reduction", "setup_s")}, d["e2e"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["breakdown_ms"].items()})
except Exception as e:
    print("unreadable", e)
PY
el done 0
