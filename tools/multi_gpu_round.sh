#!/bin/bash
# Usage (through gpurun --gpus N): tools/multi_gpu_round.sh N <out_dir under gpurun_out>
# One target spanning N GPUs checked against the CPU oracle (tile sharding, draw-order sharding) and the bench lines at N.
n=$1; out=gpurun_out/$2; mkdir -p "$out"
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
$T --master-port 29511 tests/multi_gpu/tile_sharding_check.py > "$out/tile$n.json" 2> "$out/tile$n.err"; echo "tile rc=$?"
$T --master-port 29512 tests/multi_gpu/order_sharding_check.py > "$out/order$n.json" 2> "$out/order$n.err"; echo "order rc=$?"
p=29520
for c in 3 4 5; do
  $T --master-port $p bench.py --gpus $n --config $c --no-cpu-baseline > "$out/bench_c${c}_$n.json" 2> "$out/bench_c${c}_$n.err"; echo "bench c$c rc=$?"; p=$((p+1))
done
timeout 600 python -m pytest tests -m gpu -x -q -k "sharded" > "$out/pytest_sharded_$n.log" 2>&1; tail -2 "$out/pytest_sharded_$n.log"
