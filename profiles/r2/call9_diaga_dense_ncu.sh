#!/bin/bash
# Round 2, GPU call 9: merged controlled phases (G_DIAGA) A/B on QFT, sanity of the restored one-CTA-per-tile kernel,
# ncu of the dense-k kernels (sm__throughput vs dram__throughput per k, BASELINE configuration 4) and of k_tile_pass.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c9; mkdir -p $O
python -m pytest tests -m "gpu and not slow" -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "layered f32" $B
run "layered f64" $B --precision 64
run "qft f32 diaga" $B --workload qft
run "qft f32 no-diaga" $B --workload qft --res4 4
run "qft f64 diaga" $B --workload qft --precision 64
run "qft f64 no-diaga" $B --workload qft --precision 64 --res4 4
} > $O/bench.log 2>&1
M=gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread
D="python profiles/dense_k_sweep.py 30 1"
$D > $O/dense_plain_30q.jsonl 2>&1 && ncu --metrics $M --clock-control none -k regex:k_dense --csv --log-file $O/dense_k_ncu_30q.csv $D > $O/ncu_dense.log 2>&1
B1="python bench.py --qubits 30 --steps 1 --warmup 3 --no-e2e --no-cpu"
$B1 > $O/plain30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 72 -c 2 -o $O/prof_r2_30q $B1 > $O/ncu_full30.log 2>&1
tail -3 $O/pytest_gpu.log
