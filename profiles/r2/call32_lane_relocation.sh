#!/bin/bash
# Round 2, GPU call 32 (the last 2.7 GPU-minutes of the round): hardware parity and A/B of the end-of-pass lane relocation.
#   1. tests/tools/gpu_lane_reloc_check (C, no Python start-up): 22 q layered circuit against the CPU oracle, f32 + f64, three lane
#      policies; then 30 q layered f32 / QFT f32 / layered f64 timed with reserved[6] = 3 (no relocation), 0 (default), 4 (conflicts only)
#   2. as much of `pytest -m gpu` as the remaining budget allows (the whole suite took 90 s in call 30)
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c32; mkdir -p $O
C=tests/tools/circuits
timeout 60 tests/tools/_bin/gpu_lane_reloc_check $C/layered_22q_d5.qasm 4 $C/layered_30q_d20.qasm $C/qft_30q.qasm $C/layered_30q_d20.qasm:64 > $O/lane_reloc_check.jsonl 2>&1
echo "check rc=$?" | tee -a $O/lane_reloc_check.jsonl
PYTHONUNBUFFERED=1 timeout 110 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -12 $O/lane_reloc_check.jsonl; tail -3 $O/pytest_gpu.log
