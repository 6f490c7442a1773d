"""Offline fuzz of the planner + device encoding on the CPU doubles (tests/hostcheck) against the oracle:
random superset / reference-gate / generic-matrix / phase-ladder / CX-chain circuits on 1..15 qubits, both
precisions, both doubles, low-bits variants.  Usage: python tests/tools/fuzz_doubles.py <seconds> <first seed>
Prints FAIL lines and a final "done runs N bad M".  Round 1: > 400k circuits without a failure."""
import os
import sys, os, math, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits, Gate

def rand_unitary(rng):
    a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    qq, r = np.linalg.qr(a)
    return qq * (np.diag(r) / np.abs(np.diag(r)))

def rand_generic(n, ng, rng):
    """gates given directly as matrices with random control masks (incl. multi-control, near-singular m00)"""
    gates, ops = [], []
    for _ in range(ng):
        t = int(rng.randint(n))
        U = rand_unitary(rng)
        kind = rng.randint(6)
        if kind == 0: U = np.array([[0, 1], [1, 0]], dtype=complex)
        if kind == 1: U = np.diag([1, np.exp(1j * rng.uniform(-3, 3))])
        if kind == 2:
            th = rng.uniform(-3.2, 3.2) if rng.rand() < .7 else math.pi - 1e-9 * rng.rand()
            U = np.array([[math.cos(th / 2), -1j * math.sin(th / 2)], [-1j * math.sin(th / 2), math.cos(th / 2)]])
        if kind == 3: U = np.array([[1, 1], [1, -1]]) / math.sqrt(2)
        ctrl = 0
        if n > 1 and rng.rand() < 0.4:
            others = [x for x in range(n) if x != t]
            k = 1 + (rng.rand() < 0.25 and len(others) > 1)
            for c in rng.choice(others, size=k, replace=False): ctrl |= 1 << int(c)
        g = Gate(); g.controls = ctrl; g.target = t
        m = [U[0, 0], U[0, 1], U[1, 0], U[1, 1]]
        for k in range(4): g.m[2 * k], g.m[2 * k + 1] = m[k].real, m[k].imag
        gates.append(g); ops.append((ctrl, t, U))
    return gates, ops

def oracle_generic(ops, n):
    v = np.zeros(1 << n, dtype=complex); v[0] = 1
    idx = np.arange(1 << n)
    for ctrl, t, U in ops:
        sel = ((idx & ctrl) == ctrl) & (((idx >> t) & 1) == 0)
        i0 = idx[sel]; i1 = i0 | (1 << t)
        a, b = v[i0].copy(), v[i1].copy()
        v[i0] = U[0, 0] * a + U[0, 1] * b; v[i1] = U[1, 0] * a + U[1, 1] * b
    return v

def main():
    t_end = time.time() + float(sys.argv[1])
    seed = int(sys.argv[2]); bad = 0; runs = 0
    while time.time() < t_end:
        rng = np.random.RandomState(seed)
        n = int(rng.randint(1, 16)); ng = int(rng.randint(1, 260))
        prec = 32 if rng.rand() < .5 else 64
        low = 0 if rng.rand() < .7 else int(rng.choice([3, 5, 6] if prec == 32 else [2, 4, 5]))
        blob = bool(rng.rand() < .5)
        helpers.hostcheck_use_blob(blob)
        helpers.hostcheck_set_climb(int(seed) % 9 - 1)   # the product's search (-1) and every fixed order of the hill climbing's swaps
        mode = rng.randint(5)
        try:
            if mode == 0:
                circ = circuits.random_superset(n, ng, seed); gates = q.gates_from_circuit(circ); want = helpers.oracle_run_circuit(circ, n)
            elif mode == 1:
                circ = circuits.random_reference_gates(n, ng, seed); gates = q.gates_from_circuit(circ); want = helpers.oracle_run_circuit(circ, n)
            elif mode == 4:
                circ = [("h", (k,), ()) for k in range(n)]
                for _ in range(int(rng.randint(2, 14))):
                    t = int(rng.randint(n))
                    for _ in range(int(rng.randint(2, 10))):
                        r = rng.rand()
                        if n > 1 and r < 0.6:
                            c = int(rng.choice([x for x in range(n) if x != t])); circ.append(("cx", (c, t), ()))
                        elif r < 0.75: circ.append(("rz", (t,), (float(rng.uniform(-3, 3)),)))
                        elif r < 0.85: circ.append(("x", (t,), ()))
                        elif r < 0.93 or n < 2: circ.append(("sx", (t,), ()))
                        else:
                            c = int(rng.choice([x for x in range(n) if x != t])); circ.append(("cx", (t, c), ()))
                gates = q.gates_from_circuit(circ); want = helpers.oracle_run_circuit(circ, n)
            elif mode == 3:
                circ = []
                for _ in range(ng):
                    r = rng.rand()
                    if r < 0.15: circ.append(("h", (int(rng.randint(n)),), ()))
                    elif r < 0.25: circ.append((["t", "s", "z", "sdg", "tdg"][rng.randint(5)], (int(rng.randint(n)),), ()))
                    elif r < 0.45 or n < 2: circ.append(("rz", (int(rng.randint(n)),), (float(rng.uniform(-7, 7)),)))
                    elif r < 0.9:
                        a, b = rng.choice(n, size=2, replace=False)
                        circ.append((["cp", "cz"][rng.rand() < .2], (int(a), int(b)), (float(math.pi / 2 ** rng.randint(0, 12)) if rng.rand() < .5 else float(rng.uniform(-7, 7)),)))
                    else:
                        a, b = rng.choice(n, size=2, replace=False); circ.append(("swap", (int(a), int(b)), ()))
                circ = [c if c[0] != "cz" else ("cz", c[1], ()) for c in circ]
                gates = q.gates_from_circuit(circ); want = helpers.oracle_run_circuit(circ, n)
            else:
                gl, ops = rand_generic(n, ng, rng); gates = (Gate * len(gl))(*gl); want = oracle_generic(ops, n)
            if len(gates) == 0: seed += 1; continue
            got, rep = helpers.hostcheck_run(gates, n, prec, low)
            err = float(np.max(np.abs(got - want)))
            ok = err < (3e-6 if (blob and prec == 32) else 1e-11) and rep["bad_slots"] == 0 and rep["max_conflict"] == 1 and rep["noncontig"] == 0
        except Exception as e:
            ok, err, rep = False, repr(e), {}
        runs += 1
        if not ok:
            bad += 1
            print("FAIL", dict(seed=seed, n=n, ng=ng, prec=prec, low=low, blob=blob, mode=int(mode), err=err, rep=rep), flush=True)
        seed += 1
    print("done runs", runs, "bad", bad, "last seed", seed, flush=True)
main()
