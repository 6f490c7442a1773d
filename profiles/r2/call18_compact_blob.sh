#!/bin/bash
# Round 2, GPU call 18: constant-cache footprint: GRound with XOR-basis words only (96 instead of 192 bytes) + every round's
# tables contiguous in the pass descriptor (default build) vs the previous commit (_prev).  Same box, back to back.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c18; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_readout.py -m "gpu and not slow" -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "prev f32 #$rep" env QSB_LIB_SUFFIX=_prev $B
run "compact f32 #$rep" $B
done
run "prev f64" env QSB_LIB_SUFFIX=_prev $B --precision 64
run "compact f64" $B --precision 64
run "prev qft" env QSB_LIB_SUFFIX=_prev $B --workload qft
run "compact qft" $B --workload qft
run "prev f32 cap12" env QSB_LIB_SUFFIX=_prev $B --cost-cap 12
run "compact f32 cap12" $B --cost-cap 12
} > $O/bench.log 2>&1
tail -2 $O/pytest.log
