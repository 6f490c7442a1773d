#!/bin/bash
# Round 2, GPU call 27 (2 GPUs): the final build: whole GPU suite (with 2 GPUs visible the multi-GPU test runs too: all exchange
# flavours, f32 and f64, against the oracle) + the 2-GPU bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c27; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{ run "2gpu default" $T bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu; } > $O/bench.log 2>&1
tail -3 $O/pytest_gpu.log
