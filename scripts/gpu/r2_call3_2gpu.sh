# Round 2, call 3 (2 GPUs): distributed level vectors (alfib_level_set_halo) against the serial oracle, NCCL and
# NVLink peer-memory transports, rank-locally generated bricks, timings against the replicated design, weak bench.
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
p=29600
for c in ldc3d-sv-k3-tiny ldc2d-pkp0-tiny ldc3d-pkp0-tiny; do
  p=$((p+1)); timeout 240 $TR --master-port $p scripts/dist_check_halo.py $c > gpurun_out/r2_halo_${c}_n$N.log 2>&1; el halo-$c $?
  grep "rel diff\|Error\|error" gpurun_out/r2_halo_${c}_n$N.log | tail -12
done
p=$((p+1)); ALFIB_PEER=1 timeout 240 $TR --master-port $p scripts/dist_check_halo.py ldc3d-sv-k3-tiny > gpurun_out/r2_halo_peer_tiny_n$N.log 2>&1; el halo-peer-tiny $?
grep "rel diff\|Error\|error" gpurun_out/r2_halo_peer_tiny_n$N.log | tail -12
p=$((p+1)); timeout 300 $TR --master-port $p scripts/dist_check_bricks.py > gpurun_out/r2_bricks_n$N.log 2>&1; el bricks $?
grep -v "^\[synth\|Warning\|warn" gpurun_out/r2_bricks_n$N.log | tail -12
p=$((p+1)); ALFIB_PEER=1 timeout 300 $TR --master-port $p scripts/dist_check_bricks.py > gpurun_out/r2_bricks_peer_n$N.log 2>&1; el bricks-peer $?
grep -v "^\[synth\|Warning\|warn" gpurun_out/r2_bricks_peer_n$N.log | tail -12
p=$((p+1)); timeout 600 $TR --master-port $p scripts/dist_check_halo.py ldc3d-sv-k3 --time > gpurun_out/r2_halo_cfg5_n$N.log 2>&1; el halo-cfg5 $?
grep "world\|events\|Error" gpurun_out/r2_halo_cfg5_n$N.log | tail -6
p=$((p+1)); ALFIB_PEER=1 timeout 600 $TR --master-port $p scripts/dist_check_halo.py ldc3d-sv-k3 --time > gpurun_out/r2_halo_peer_cfg5_n$N.log 2>&1; el halo-peer-cfg5 $?
grep "world\|events\|Error" gpurun_out/r2_halo_peer_cfg5_n$N.log | tail -6
p=$((p+1)); timeout 900 $TR --master-port $p bench.py --gpus $N --steps 5 --warmup 3 --scaling weak --no-cpu-baseline > gpurun_out/r2_bench_weak_n$N.json 2> gpurun_out/r2_bench_weak_n$N.log; el bench-weak $?
tail -4 gpurun_out/r2_bench_weak_n$N.log; cut -c1-600 gpurun_out/r2_bench_weak_n$N.json
el done 0
