#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): sharded parity on hardware at this revision (all exchange flavours, incl. the direct
# fused exchange) and the 34 q strong-scaling point with each flavour; bench.py's own parity line.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c6; mkdir -p $O
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "2gpu default(pipelined)" $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu
run "2gpu fused direct"       $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --fused 1
run "2gpu fused round1"       $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --fused 4
} > $O/bench.log 2>&1
tail -3 $O/pytest_multi.log
