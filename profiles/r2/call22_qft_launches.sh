#!/bin/bash
# Round 2, GPU call 22: where the QFT time goes: per-pass launch list (time, DRAM bytes, issue, FMA pipe, instructions) and one
# --set full capture of the pass with the longest phase runs.  Plain run first; nothing printed under ncu is a bench value.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c22; mkdir -p $O
B="python bench.py --qubits 30 --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras --workload qft"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.per_cycle_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed.sum
$B > $O/plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:k_tile_pass -s 12 -c 4 --csv --log-file $O/launches_qft30.csv $B > $O/ncu_l.log 2>&1
$B > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 14 -c 1 -o $O/prof_qft30_pass3 $B > $O/ncu_full.log 2>&1
tail -1 $O/plain.log
