# Round 2, call 8 (2 GPUs): LL-protocol mailbox exchanges (parity + cost); TMA tile-op configurations with compile-time trip counts
mkdir -p gpurun_out
export ALFIB_PROBLEM_CACHE=/tmp/alfib_cache
t0=$(date +%s)
el() { echo "[$1] rc=$2 after $(( $(date +%s) - t0 ))s"; }
CUDA_VISIBLE_DEVICES=0 timeout 400 python scripts/apply_variants.py ldc3d-sv-k3 100 > gpurun_out/r2_apply_variants5.txt 2> gpurun_out/r2_apply_variants5.err; el variants $?; grep variant gpurun_out/r2_apply_variants5.txt | cut -c1-200; tail -2 gpurun_out/r2_apply_variants5.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
p=29900
for c in ldc3d-sv-k3-tiny ldc2d-pkp0-tiny ldc3d-pkp0-tiny; do
  p=$((p+1)); ALFIB_PEER=1 timeout 240 $TR --master-port $p scripts/dist_check_halo.py $c > gpurun_out/r2_ll_${c}_n2.log 2>&1; el ll-$c $?
  grep "rel diff\|Error\|error" gpurun_out/r2_ll_${c}_n2.log | tail -8
done
p=$((p+1)); ALFIB_PEER=1 timeout 300 $TR --master-port $p scripts/dist_check_bricks.py > gpurun_out/r2_ll_bricks_n2.log 2>&1; el ll-bricks $?
grep "world" gpurun_out/r2_ll_bricks_n2.log | tail -3
p=$((p+1)); ALFIB_PEER=1 timeout 400 $TR --master-port $p scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_ll.log 2>&1; el dkb-n2-ll $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_ll.log | cut -c1-300
p=$((p+1)); ALFIB_PEER=1 ALFIB_DEBUG_SKIP_SPMV=1 timeout 400 $TR --master-port $p scripts/dist_kernel_bench.py ldc3d-sv-k3 50 > gpurun_out/r2_dkb_n2_ll_nospmv.log 2>&1; el dkb-n2-ll-nospmv $?; grep "^level\|cycle_ms" gpurun_out/r2_dkb_n2_ll_nospmv.log | cut -c1-300
el done 0
