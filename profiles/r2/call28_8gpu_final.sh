#!/bin/bash
# Round 2, GPU call 28 (8 GPUs): the 34 q strong-scaling points at 8 and 4 ranks with the final build; every line carries the
# 24 q sharded-vs-single-GPU parity check made before timing.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c28; mkdir -p $O
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572"
{
run "8gpu default" timeout 300 $T8 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu
run "4gpu default" timeout 300 $T4 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu
} > $O/bench.log 2>&1
tail -c 600 $O/bench.log
