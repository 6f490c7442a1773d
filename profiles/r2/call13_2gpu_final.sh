#!/bin/bash
# Round 2, GPU call 13 (2 GPUs): hardware parity of the sharded path at the shipped revision (threshold search included) + bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c13; mkdir -p $O
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{ run "2gpu default" $T bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu; } > $O/bench.log 2>&1
tail -3 $O/pytest_multi.log
