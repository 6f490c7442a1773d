#!/bin/bash
# Round 2, GPU call 1: evidence for the kernel as shipped at the end of round 1 (HEAD 744211b), before any change.
#   1. pytest -m gpu   2. the queued A/B runs (profiles/next_round_ab.sh)   3. ncu: --set full of k_tile_pass at 30 q
#   (default schedule and --cost-cap 12), per-launch DRAM bytes at 30 q and 34 q.
# gpurun --timeout 1500 -- 'bash profiles/r2/call1_evidence_at_head.sh'
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c1; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
bash profiles/next_round_ab.sh > $O/next_round_ab.log 2>&1
B30="python bench.py --qubits 30 --steps 1 --warmup 3 --no-e2e --no-cpu"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.per_cycle_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed.sum
$B30 > $O/plain30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 72 -c 3 -o $O/prof_head_30q $B30 > $O/ncu_full30.log 2>&1
$B30 > $O/plain30b.log 2>&1 && ncu --metrics $M --clock-control none -k regex:k_tile_pass -s 66 -c 22 --csv --log-file $O/launches_30q.csv $B30 > $O/ncu_l30.log 2>&1
$B30 --cost-cap 12 > $O/plain30c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 170 -c 2 -o $O/prof_head_30q_cap12 $B30 --cost-cap 12 > $O/ncu_full30c.log 2>&1
B34="python bench.py --qubits 34 --steps 1 --warmup 3 --no-e2e --no-cpu"
$B34 > $O/plain34.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_tile_pass -s 78 -c 26 --csv --log-file $O/launches_34q.csv $B34 > $O/ncu_l34.log 2>&1
ls -la $O
