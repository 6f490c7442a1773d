#!/bin/bash
# Round 2, GPU call 21: merged phase runs on the pack qubit (G_DIAGA + QSB_NVB) vs the same kernel with the merge off (_nopack).
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c21; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_readout.py -m "gpu and not slow" -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "nopack qft #$rep" env QSB_LIB_SUFFIX=_nopack $B --workload qft
run "pack qft #$rep" $B --workload qft
done
run "nopack f32 layered" env QSB_LIB_SUFFIX=_nopack $B
run "pack f32 layered" $B
run "pack f64 layered" $B --precision 64
run "pack f64 qft" $B --precision 64 --workload qft
} > $O/bench.log 2>&1
tail -2 $O/pytest.log
