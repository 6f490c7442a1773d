#!/bin/bash
# Round 2, GPU call 12: the shipped build: whole GPU suite (incl. the large-size tests), smoke(), XOR-basis A/B, default bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c12; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
B="python bench.py --qubits 30 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "noxb f32 #1" env QSB_LIB_SUFFIX=_noxb $B
run "default f32 #1" $B
run "noxb f32 #2" env QSB_LIB_SUFFIX=_noxb $B
run "default f32 #2" $B
run "noxb f64" env QSB_LIB_SUFFIX=_noxb $B --precision 64
run "default f64" $B --precision 64
} > $O/ab.log 2>&1
( time python bench.py --steps 5 --warmup 3 ) > $O/bench_n1.log 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/bench_n1.err
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log
