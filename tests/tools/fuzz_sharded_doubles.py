"""Offline fuzz of the sharded schedules (plain and fused exchange, 2/4/8 ranks) on the CPU doubles.
Usage: python tests/tools/fuzz_sharded_doubles.py <seconds> <first seed>"""
import os
import sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits
BLOB = os.environ.get("QSB_FUZZ_BLOB") == "1"   # interpret the device encoding (what the kernel reads) instead of the logical tables
helpers.hostcheck_use_blob(BLOB)
t_end = time.time() + float(sys.argv[1]); seed = int(sys.argv[2]); runs = bad = 0
while time.time() < t_end:
    rng = np.random.RandomState(seed)
    world = int(rng.choice([2, 4, 8])); g = {2: 1, 4: 2, 8: 3}[world]
    prec = 32 if rng.rand() < .5 else 64
    T = 13 if prec == 32 else 12
    n = int(rng.randint(T + 2 * g, T + 2 * g + 4))
    fused = bool(rng.rand() < .5)
    smo = int(rng.choice([0, 2, 4, 10, 14, 20]))   # 0: the product's search over thresholds and lane policies
    mode = rng.randint(3)
    if mode == 0: circ = circuits.random_superset(n, int(rng.randint(20, 200)), seed)
    elif mode == 1: circ = circuits.random_layered(n, depth=int(rng.randint(1, 6)), seed=seed)
    else: circ = circuits.qft(n) if rng.rand() < .5 else circuits.random_reference_gates(n, int(rng.randint(20, 200)), seed)
    try:
        got, rep = helpers.sharded_host_run(q.gates_from_circuit(circ), n, world, prec, swap_min_ops=smo, fused=fused)
        want = helpers.oracle_run_circuit(circ, n)
        err = float(np.max(np.abs(got - want))); ok = err < (3e-6 if BLOB and prec == 32 else 1e-11) and rep["bad_slots"] == 0 and rep["max_conflict"] == 1 and rep["noncontig"] == 0
    except Exception as e:
        ok, err, rep = False, repr(e), {}
    runs += 1
    if not ok:
        bad += 1; print("FAIL", dict(seed=seed, n=n, world=world, prec=prec, fused=fused, smo=smo, mode=int(mode), err=err, rep=rep), flush=True)
    seed += 1
print("done runs", runs, "bad", bad, flush=True)
