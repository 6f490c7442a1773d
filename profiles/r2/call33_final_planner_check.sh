#!/bin/bash
# Round 2, GPU call 33 (the last 30 GPU-seconds of the round): the final planner (lane relocation + CX -> controlled phase next to
# an h + four hill-climbing orders) on hardware through the C ABI: 22 q parity against the CPU oracle (f32, f64; default, old
# planner, conflicts-only relocation) and 30 q layered f32 timed with the old planner / lane relocation only / the default.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c33; mkdir -p $O
C=tests/tools/circuits
timeout 24 tests/tools/_bin/gpu_lane_reloc_check $C/layered_22q_d5.qasm 3 $C/layered_30q_d20.qasm $C/layered_30q_d20.qasm:64 $C/qft_30q.qasm > $O/final_planner_check.jsonl 2>&1
echo "check rc=$?" | tee -a $O/final_planner_check.jsonl
cat $O/final_planner_check.jsonl
