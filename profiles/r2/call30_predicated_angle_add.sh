#!/bin/bash
# Round 2, GPU call 30: angle entries as LOP3-to-predicate + predicated add (2 instructions per entry instead of 4 + selects)
# vs the build of call 29 (_prev); then the whole validation of the default build.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c30; mkdir -p $O
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
for rep in 1 2; do
run "prev qft #$rep" env QSB_LIB_SUFFIX=_prev $B --workload qft
run "new qft #$rep" $B --workload qft
done
run "prev qft f64" env QSB_LIB_SUFFIX=_prev $B --workload qft --precision 64
run "new qft f64" $B --workload qft --precision 64
run "prev f32 layered" env QSB_LIB_SUFFIX=_prev $B
run "new f32 layered" $B
run "prev f64 layered" env QSB_LIB_SUFFIX=_prev $B --precision 64
run "new f64 layered" $B --precision 64
} > $O/ab.log 2>&1
bash profiles/r2/call20_final_validation.sh r2c30
