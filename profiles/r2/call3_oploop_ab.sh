#!/bin/bash
# Round 2, GPU call 3: op-loop flavours of the flat-stream interpreter, same box back to back.
#   default build: QSB_OPLOOP=0 (plain loop, per-thread constant load of the selected set)
#   _c: QSB_OPLOOP=2 (+ header of the next record prefetched)   _a: QSB_OPLOOP=1 (uniform prefetch + selects; call 2)
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c3; mkdir -p $O
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for sfx in "" _c _a; do
  run "oploop$sfx f32" env QSB_LIB_SUFFIX=$sfx $B
  run "oploop$sfx f64" env QSB_LIB_SUFFIX=$sfx $B --precision 64
done
run "oploop qft f32" $B --workload qft
} > $O/bench.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
