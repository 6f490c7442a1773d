#!/bin/bash
# Round 2, GPU call 10: 32 vectors per thread (-DQSB_NVB=5 -DQSB_TB=6, 64-thread CTAs x 4 per SM, 254 registers) against the
# default geometry (16 vectors, 128-thread CTAs x 4 per SM), same box back to back; _nvb5t7 = 32 vectors x 128 threads x 2 CTAs/SM (2^13 tiles).
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c10; mkdir -p $O
QSB_LIB_SUFFIX=_nvb5 python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q > $O/pytest_nvb5.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_nvb5.log
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "default f32" $B
run "nvb5 f32" env QSB_LIB_SUFFIX=_nvb5 $B
run "default f64" $B --precision 64
run "nvb5 f64" env QSB_LIB_SUFFIX=_nvb5 $B --precision 64
run "default qft f32" $B --workload qft
run "nvb5 qft f32" env QSB_LIB_SUFFIX=_nvb5 $B --workload qft
run "nvb5 f32 cap12" env QSB_LIB_SUFFIX=_nvb5 $B --cost-cap 12
run "nvb5t7 f32" env QSB_LIB_SUFFIX=_nvb5t7 $B
run "nvb5t7 f64" env QSB_LIB_SUFFIX=_nvb5t7 $B --precision 64
run "nvb5t7 qft f32" env QSB_LIB_SUFFIX=_nvb5t7 $B --workload qft
} > $O/bench.log 2>&1
# dense k = 4, 5: sm__throughput vs dram__throughput (call 9 captured k = 2, 3 before its time limit; 12 launches per k suffice)
M=gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
for k in 4 5; do
  D="python profiles/dense_k_sweep.py 30 1 $k"
  $D > $O/dense_plain_k$k.jsonl 2>&1 && ncu --metrics $M --clock-control none -k regex:k_dense -s 20 -c 12 --csv --log-file $O/dense_k${k}_ncu_30q.csv $D > $O/ncu_dense_k$k.log 2>&1
done
B1="python bench.py --qubits 30 --steps 1 --warmup 3 --no-e2e --no-cpu"
$B1 > $O/plain30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 72 -c 2 -o $O/prof_r2_30q $B1 > $O/ncu_full30.log 2>&1
tail -3 $O/pytest_nvb5.log
