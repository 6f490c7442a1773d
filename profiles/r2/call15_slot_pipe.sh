#!/bin/bash
# Round 2, GPU call 15: slot select (predicate, coefficient select, deferred-X toggle) of slot k+1 issued before the packed
# arithmetic of slot k (-DQSB_SLOT_PIPE, build _sp) vs the default build.  Same box, back to back.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c15; mkdir -p $O
QSB_LIB_SUFFIX=_sp python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q > $O/pytest_sp.log 2>&1; echo "pytest sp rc=$?" | tee -a $O/pytest_sp.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "default f32 #$rep" $B
run "slot-pipe f32 #$rep" env QSB_LIB_SUFFIX=_sp $B
done
run "default f64" $B --precision 64
run "slot-pipe f64" env QSB_LIB_SUFFIX=_sp $B --precision 64
run "default qft" $B --workload qft
run "slot-pipe qft" env QSB_LIB_SUFFIX=_sp $B --workload qft
} > $O/bench.log 2>&1
tail -2 $O/pytest_sp.log
