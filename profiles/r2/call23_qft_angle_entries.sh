#!/bin/bash
# Round 2, GPU call 23: QFT follow-up of call 22 (ncu: a QFT pass executes 8600 instructions per thread; 21 % in G_DIAG_GEN's
# predicated sweep, 26 % in angle entries at ~10.5 instructions each, generic LD in the merged-run loop):
# 8-byte angle entries tested with one 32-bit mask against the predicate word, specials addressed by index (constant-bank
# loads), one G_DIAG_GEN body per mask.  Default build vs the previous commit (_prev), same box.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c23; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_readout.py -m "gpu and not slow" -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "prev qft #$rep" env QSB_LIB_SUFFIX=_prev $B --workload qft
run "new qft #$rep" $B --workload qft
done
run "prev qft f64" env QSB_LIB_SUFFIX=_prev $B --workload qft --precision 64
run "new qft f64" $B --workload qft --precision 64
for rep in 1 2; do
run "prev f32 layered #$rep" env QSB_LIB_SUFFIX=_prev $B
run "new f32 layered #$rep" $B
done
run "prev f64 layered" env QSB_LIB_SUFFIX=_prev $B --precision 64
run "new f64 layered" $B --precision 64
} > $O/bench.log 2>&1
tail -2 $O/pytest.log
