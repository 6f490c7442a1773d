#!/bin/bash
# Round 2, GPU call 11: per-round overhead: exchange base offsets precomputed during the gather (default build) vs computed
# per round (_nosbt = previous kernel) vs + exchange addresses from 4 XOR basis words (_xb).  Same box, back to back.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c11; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/pytest_default.log
QSB_LIB_SUFFIX=_xb python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q > $O/pytest_xb.log 2>&1; echo "pytest xb rc=$?" | tee -a $O/pytest_xb.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "nosbt f32 #$rep" env QSB_LIB_SUFFIX=_nosbt $B
run "sbt f32 #$rep" $B
run "sbt+xb f32 #$rep" env QSB_LIB_SUFFIX=_xb $B
done
run "nosbt f64" env QSB_LIB_SUFFIX=_nosbt $B --precision 64
run "sbt f64" $B --precision 64
run "sbt+xb f64" env QSB_LIB_SUFFIX=_xb $B --precision 64
run "nosbt qft" env QSB_LIB_SUFFIX=_nosbt $B --workload qft
run "sbt qft" $B --workload qft
run "sbt+xb qft" env QSB_LIB_SUFFIX=_xb $B --workload qft
} > $O/bench.log 2>&1
tail -2 $O/pytest_default.log $O/pytest_xb.log
