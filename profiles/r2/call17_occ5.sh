#!/bin/bash
# Round 2, GPU call 17: 5 CTAs/SM (96 registers) WITH the shared-memory carve-out that lets 5 CTAs be resident
# (call 4 measured this build at 4 resident CTAs: the carve-out was left to the driver).  QSB_VERBOSE_OCC prints the residency.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c17; mkdir -p $O
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | grep -E "^\{|resident" | tail -3; }
{
run "default f32" env QSB_VERBOSE_OCC=1 $B
run "occ5 f32" env QSB_VERBOSE_OCC=1 QSB_LIB_SUFFIX=_occ5 $B
run "default f64" env QSB_VERBOSE_OCC=1 $B --precision 64
run "occ5 f64" env QSB_VERBOSE_OCC=1 QSB_LIB_SUFFIX=_occ5 $B --precision 64
run "occ5 f32 cap12" env QSB_LIB_SUFFIX=_occ5 $B --cost-cap 12
run "occ5 qft f32" env QSB_LIB_SUFFIX=_occ5 $B --workload qft
} > $O/bench.log 2>&1
QSB_LIB_SUFFIX=_occ5 python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q > $O/pytest_occ5.log 2>&1; echo "pytest occ5 rc=$?" | tee -a $O/pytest_occ5.log
