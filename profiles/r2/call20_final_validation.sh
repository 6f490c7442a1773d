#!/bin/bash
# Round 2, GPU call 20: the build shipped at the end of the round (compact descriptor): whole GPU suite, smoke(), default bench line
# (34 q + the per-configuration extras) and the reference arm.
cd "$(dirname "$0")/../.."
O=gpurun_out/${1:-r2c20}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
( time python bench.py --steps 5 --warmup 3 ) > $O/bench_n1.log 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/bench_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc=$?" | tee -a $O/bench_ref.err
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log
