#!/bin/bash
# Round 2, GPU call 5: the new bench line (configs, e2e_full_30q, reference_gpu, full-depth cpu_baseline) and the large parity tests.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c5; mkdir -p $O
python -m pytest tests/test_gpu_large.py -m gpu -x -q --durations=5 > $O/pytest_large.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_large.log
( time python bench.py --steps 5 --warmup 3 ) > $O/bench_n1.log 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/bench_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.log 2> $O/bench_ref.err
tail -3 $O/pytest_large.log; tail -5 $O/bench_n1.err
