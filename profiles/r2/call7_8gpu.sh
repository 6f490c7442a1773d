#!/bin/bash
# Round 2, GPU call 7 (8 GPUs): hardware parity (quick worker + bench.py's parity line) and the 34 q strong-scaling point at 8
# (direct fused exchange = default, vs the round-1 fused flavour), then the same at 4 ranks on the same box.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c7; mkdir -p $O
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572"
QSB_DIST_QUICK=1 $T8 tests/dist_gpu_worker.py > $O/dist_quick8.log 2>&1; echo "dist quick rc=$?" | tee -a $O/dist_quick8.log
{
run "8gpu default(fused direct)" $T8 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu
run "8gpu fused round1"          $T8 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e --fused 4
run "4gpu default(fused direct)" $T4 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu
run "4gpu fused round1"          $T4 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu --no-e2e --fused 4
} > $O/bench.log 2>&1
tail -2 $O/dist_quick8.log
