#!/bin/bash
# Round 2, GPU call 2: the flat op-stream interpreter (first version): parity suite + 30 q timings.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c2; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "v2 f32" $B
run "v2 f64" $B --precision 64
run "v2 qft f32" $B --workload qft
run "v2 f32 cap12" $B --cost-cap 12
} > $O/bench.log 2>&1
tail -3 $O/pytest_gpu.log
