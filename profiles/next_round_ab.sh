#!/bin/bash
# First GPU call of the next round: the A/B runs that could not be measured in round 1 (GPU budget spent).
#   1. tile choice: hill-climbed (default) vs first come (--no-sink 2 -> qsb_options_t.reserved[6] = 2)
#   2. tile geometry: QSB_TB = 6 / 7 (default) / 8   (64 / 128 / 256 threads, 8 / 4 / 2 CTAs per SM)
#   3. contiguous low bits: 3 vs 4 (f32; plans 18 instead of 22 passes at 30 q), 2 vs 3 (f64)
# Build the variants HERE (no GPU needed), then run this script on the box:
#   make -C gpu_quantum_simulator_b200/csrc SUFFIX=_tb6 EXTRA=-DQSB_TB=6
#   make -C gpu_quantum_simulator_b200/csrc SUFFIX=_tb8 EXTRA=-DQSB_TB=8
#   gpurun --timeout 400 -- 'bash profiles/next_round_ab.sh > gpurun_out/next_round_ab.log 2>&1'
# One JSON bench line per configuration, tagged by the "cfg=" line in front of it.
cd "$(dirname "$0")/.."
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
run "tb7 climbed f32"            $B
run "tb7 first-come f32"         $B --no-sink 2
run "tb7 climbed f64"            $B --precision 64
run "tb7 climbed low_bits=3 f32" $B --low-bits 3
run "tb7 climbed qft f32"        $B --workload qft
run "tb7 climbed low_bits=2 f64" $B --precision 64 --low-bits 2
for sfx in _tb6 _tb8; do
  if [ -f gpu_quantum_simulator_b200/libqsim_b200$sfx.so ]; then
    run "$sfx climbed f32" env QSB_LIB_SUFFIX=$sfx $B
    run "$sfx climbed f64" env QSB_LIB_SUFFIX=$sfx $B --precision 64
  fi
done
run "tb7 climbed 34q f32" python bench.py --qubits 34 --steps 3 --warmup 3 --no-e2e --no-cpu
