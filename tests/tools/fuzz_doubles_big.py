"""Like fuzz_doubles.py on 15..19 qubits (several tiles: exercises the tile choice).
Usage: python tests/tools/fuzz_doubles_big.py <seconds> <first seed>"""
import os
import sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits
t_end = time.time() + float(sys.argv[1]); seed = int(sys.argv[2]); runs = bad = 0
while time.time() < t_end:
    rng = np.random.RandomState(seed)
    n = int(rng.randint(15, 20)); prec = 32 if rng.rand() < .5 else 64
    blob = bool(rng.rand() < .5); helpers.hostcheck_use_blob(blob)
    helpers.hostcheck_set_climb(int(seed) % 9 - 1)   # the product's search (-1) and every fixed order of the hill climbing's swaps
    mode = rng.randint(3)
    if mode == 0: circ = circuits.random_superset(n, int(rng.randint(50, 400)), seed)
    elif mode == 1: circ = circuits.random_layered(n, depth=int(rng.randint(1, 10)), seed=seed)
    else: circ = circuits.qft(n) + circuits.random_reference_gates(n, int(rng.randint(20, 200)), seed)
    try:
        got, rep = helpers.hostcheck_run(q.gates_from_circuit(circ), n, prec)
        want = helpers.oracle_run_circuit(circ, n)
        err = float(np.max(np.abs(got - want))); ok = err < (3e-6 if (blob and prec == 32) else 1e-11) and rep["bad_slots"] == 0 and rep["max_conflict"] == 1 and rep["noncontig"] == 0
    except Exception as e:
        ok, err, rep = False, repr(e), {}
    runs += 1
    if not ok:
        bad += 1; print("FAIL", dict(seed=seed, n=n, prec=prec, blob=blob, mode=int(mode), err=err, rep=rep), flush=True)
    seed += 1
print("done runs", runs, "bad", bad, flush=True)
