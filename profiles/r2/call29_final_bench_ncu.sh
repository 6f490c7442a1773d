#!/bin/bash
# Round 2, GPU call 29: the final build: default bench line (N = 1) and reference arm, then the ncu --set full capture of
# k_tile_pass at 30 q (plain run first; nothing measured under ncu is a bench value) and the QFT pass of call 22 again.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c29; mkdir -p $O
( time python bench.py --steps 5 --warmup 3 ) > $O/bench_n1.log 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/bench_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc=$?" | tee -a $O/bench_ref.err
B30="python bench.py --qubits 30 --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras"
$B30 > $O/plain30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 72 -c 2 -o $O/prof_final_30q $B30 > $O/ncu_full30.log 2>&1
$B30 --workload qft > $O/plainqft.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 14 -c 1 -o $O/prof_final_qft30_pass3 $B30 --workload qft > $O/ncu_fullqft.log 2>&1
tail -1 $O/plain30.log | cut -c1-200
