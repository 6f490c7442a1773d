#!/bin/bash
# Round 2, GPU call 25: thread-phase loop without nvcc's 4x unrolling (f64 layered lost 1.7 % in call 24: code size) vs the
# build of call 24 (_prev).  Same box, back to back.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c25; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "prev f64 layered #$rep" env QSB_LIB_SUFFIX=_prev $B --precision 64
run "new f64 layered #$rep" $B --precision 64
run "prev f32 layered #$rep" env QSB_LIB_SUFFIX=_prev $B
run "new f32 layered #$rep" $B
done
run "prev qft" env QSB_LIB_SUFFIX=_prev $B --workload qft
run "new qft" $B --workload qft
run "prev qft f64" env QSB_LIB_SUFFIX=_prev $B --workload qft --precision 64
run "new qft f64" $B --workload qft --precision 64
} > $O/bench.log 2>&1
tail -2 $O/pytest.log
