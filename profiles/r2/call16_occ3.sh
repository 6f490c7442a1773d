#!/bin/bash
# Round 2, GPU call 16: 3 CTAs/SM (ptxas takes 166 registers instead of the 128 it is capped at with 4 CTAs/SM) vs default.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c16; mkdir -p $O
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "default f32" $B
run "occ3 f32" env QSB_LIB_SUFFIX=_occ3 $B
run "default f64" $B --precision 64
run "occ3 f64" env QSB_LIB_SUFFIX=_occ3 $B --precision 64
run "occ3 f32 cap12" env QSB_LIB_SUFFIX=_occ3 $B --cost-cap 12
} > $O/bench.log 2>&1
