#!/bin/bash
# Round 2, GPU call 4: group interpreter (round 1) + candidates, same box back to back, 30 q depth 20.
#   default: + Hadamard-like slot form S_UNIT_H      --res4 3: without it
#   _pf: + L2 prefetch of the tile 592 CTAs ahead    _occ5: 5 CTAs/SM (96 registers, spills)   _tb6: 64-thread CTAs
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c4; mkdir -p $O
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "default f32" $B
run "default f64" $B --precision 64
run "no-hform f32" $B --res4 3
run "no-hform f64" $B --res4 3 --precision 64
for sfx in _pf _occ5 _tb6; do
  run "$sfx f32" env QSB_LIB_SUFFIX=$sfx $B
  run "$sfx f64" env QSB_LIB_SUFFIX=$sfx $B --precision 64
done
run "default qft f32" $B --workload qft
run "_pf qft f32" env QSB_LIB_SUFFIX=_pf $B --workload qft
run "_pf f32 cap12" env QSB_LIB_SUFFIX=_pf $B --cost-cap 12
} > $O/bench.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
