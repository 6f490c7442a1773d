#!/bin/bash
# Round 2, GPU call 8: persistent tile loop + cp.async prefetch: parity suite, racecheck of the shared-buffer reuse,
# A/B against one-CTA-per-tile launches of the same kernel (QSB_PERSIST=0), dense-k sweep at 32 q.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c8; mkdir -p $O
python -m pytest tests -m "gpu and not slow" -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
B="python bench.py --qubits 30 --steps 4 --warmup 3 --no-e2e --no-cpu"
run() { echo "cfg=$1"; shift; "$@" 2>&1 | tail -1; }
{
run "persist f32" $B
run "one-cta-per-tile f32" env QSB_PERSIST=0 $B
run "persist f64" $B --precision 64
run "one-cta-per-tile f64" env QSB_PERSIST=0 $B --precision 64
run "persist qft f32" $B --workload qft
run "one-cta-per-tile qft f32" env QSB_PERSIST=0 $B --workload qft
run "persist f32 cap12" $B --cost-cap 12
run "one-cta-per-tile f32 cap12" env QSB_PERSIST=0 $B --cost-cap 12
run "persist 28q f32" python bench.py --qubits 28 --steps 4 --warmup 3 --no-e2e --no-cpu
run "one-cta-per-tile 28q f32" env QSB_PERSIST=0 python bench.py --qubits 28 --steps 4 --warmup 3 --no-e2e --no-cpu
} > $O/bench.log 2>&1
python profiles/dense_k_sweep.py 32 3 > $O/dense_k_sweep_32q.jsonl 2>&1
cat > /tmp/race.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers, gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits
for prec, tol in ((q.F32, 1e-5), (q.F64, 1e-12)):
    circ = circuits.random_layered(22, depth=2, seed=5)
    want = helpers.oracle_run_circuit(circ, 22)
    with q.Simulator(22, precision=prec) as s:
        st = s.apply(q.gates_from_circuit(circ))
        err = float(np.max(np.abs(s.state() - want)))
        print("racecheck run", prec, "passes", st["passes"], "err", err)
        assert err <= tol
PY
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/race.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
tail -3 $O/pytest_gpu.log; tail -5 $O/racecheck.log
