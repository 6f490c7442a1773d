#!/bin/bash
# Round 2, GPU call 26: thread-phase loop unrolled per precision (default build) vs the build of call 24 (_prev), then the
# whole validation of the default build (GPU suite, smoke(), default bench line, reference arm).
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c26; mkdir -p $O
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "prev f64 layered #$rep" env QSB_LIB_SUFFIX=_prev $B --precision 64
run "new f64 layered #$rep" $B --precision 64
run "prev f32 layered #$rep" env QSB_LIB_SUFFIX=_prev $B
run "new f32 layered #$rep" $B
done
} > $O/ab.log 2>&1
bash profiles/r2/call20_final_validation.sh r2c26
