"""Synthetic circuits of BASELINE.json and their spellings.

A circuit is a list of (name, qubits, params) tuples, little-endian qubits, gate names as
in the QASM front end.  `rz` ALWAYS means the reference's phase gate diag(1, e^{i theta})
(/root/reference/quantum_simulator.c:205-208).

Two writers:
  to_qasm(circ, n)            native text for this framework's front end (superset gates)
  to_reference_qasm(circ, n)  the same unitary spelled in the reference's gate set
                              {cx x sx z s sdg t tdg rz h} (quantum_simulator.c:13-23), plus the
                              global phase the respelling drops -- so the reference program can
                              be the oracle for circuits it cannot parse directly.
"""
import math

import numpy as np


def random_layered(n, depth=20, seed=12345):
    """Per layer: one of {h, rx(t), rz(t)} on every qubit (t ~ U(-pi, pi)), then CX on a random
    perfect matching (random permutation, consecutive pairs, random direction).  SURVEY.md §8d."""
    rng = np.random.RandomState(seed)
    circ = []
    for _ in range(depth):
        kinds = rng.randint(0, 3, size=n)
        thetas = rng.uniform(-math.pi, math.pi, size=n)
        for q in range(n):
            if kinds[q] == 0:
                circ.append(("h", (q,), ()))
            elif kinds[q] == 1:
                circ.append(("rx", (q,), (float(thetas[q]),)))
            else:
                circ.append(("rz", (q,), (float(thetas[q]),)))
        perm = rng.permutation(n)
        flips = rng.randint(0, 2, size=n // 2)
        for k in range(n // 2):
            a, b = int(perm[2 * k]), int(perm[2 * k + 1])
            if flips[k]:
                a, b = b, a
            circ.append(("cx", (a, b), ()))
    return circ


def qft(n, with_h_layer=True, swaps=True):
    """H layer (non-trivial input), then the textbook QFT: H + controlled-phase ladder + bit reversal."""
    circ = []
    if with_h_layer:
        for q in range(n):
            circ.append(("h", (q,), ()))
    for j in reversed(range(n)):
        circ.append(("h", (j,), ()))
        for k in reversed(range(j)):
            circ.append(("cp", (k, j), (math.pi / (1 << (j - k)),)))
    if swaps:
        for q in range(n // 2):
            circ.append(("swap", (q, n - 1 - q), ()))
    return circ


def random_reference_gates(n, n_gates, seed=0):
    """Random circuit over exactly the reference's gate set."""
    rng = np.random.RandomState(seed)
    names = ["cx", "x", "sx", "z", "s", "sdg", "t", "tdg", "rz", "h"]
    circ = []
    for _ in range(n_gates):
        g = names[rng.randint(len(names))]
        if g == "cx" and n >= 2:
            a, b = rng.choice(n, size=2, replace=False)
            circ.append(("cx", (int(a), int(b)), ()))
        elif g == "cx":
            continue
        elif g == "rz":
            circ.append(("rz", (int(rng.randint(n)),), (float(rng.uniform(-math.pi, math.pi)),)))
        else:
            circ.append((g, (int(rng.randint(n)),), ()))
    return circ


def random_superset(n, n_gates, seed=0):
    """Random circuit over the superset the front end accepts (for fused-vs-oracle tests)."""
    rng = np.random.RandomState(seed)
    one = ["x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "rz", "rx", "ry"]
    two = ["cx", "cz", "cp", "swap"]
    circ = []
    for _ in range(n_gates):
        r = rng.rand()
        if r < 0.6 or n < 2:
            g = one[rng.randint(len(one))]
            p = (float(rng.uniform(-math.pi, math.pi)),) if g in ("rz", "rx", "ry") else ()
            circ.append((g, (int(rng.randint(n)),), p))
        elif r < 0.95 or n < 3:
            g = two[rng.randint(len(two))]
            a, b = rng.choice(n, size=2, replace=False)
            p = (float(rng.uniform(-math.pi, math.pi)),) if g == "cp" else ()
            circ.append((g, (int(a), int(b)), p))
        else:
            a, b, c = rng.choice(n, size=3, replace=False)
            circ.append(("ccx", (int(a), int(b), int(c)), ()))
    return circ


def _fmt(x):
    return repr(float(x))


def to_qasm(circ, n, decl="qubit[{n}] q;"):
    lines = ["OPENQASM 3.0;", 'include "stdgates.inc";', decl.format(n=n)]
    for name, qubits, params in circ:
        head = name + ("(" + ",".join(_fmt(p) for p in params) + ")" if params else "")
        lines.append(head + " " + ", ".join(f"q[{q}]" for q in qubits) + ";")
    return "\n".join(lines) + "\n"


def to_cuda_variant_text(circ, n):
    """The bare '<num_q> <num_g>' format the reference's .cu programs read (naive.cu:239-240)."""
    text, _ = to_reference_qasm(circ, n)
    body = text.split("\n")[3:]
    body = [b for b in body if b]
    return f"{n} {len(body)}\n" + "\n".join(body) + "\n"


def to_reference_gates(circ):
    """-> (list over the reference gate set, global phase angle phi with U = e^{i phi} * U_ref)."""
    out, phi = [], 0.0
    for name, q, p in circ:
        if name in ("cx", "x", "sx", "z", "s", "sdg", "t", "tdg", "rz", "h"):
            out.append((name, q, p))
        elif name == "p":
            out.append(("rz", q, p))
        elif name == "y":            # Y = i X Z
            out += [("z", q, ()), ("x", q, ())]
            phi += math.pi / 2
        elif name == "rx":           # RX(t) = e^{-it/2} H P(t) H
            out += [("h", q, ()), ("rz", q, p), ("h", q, ())]
            phi -= p[0] / 2
        elif name == "ry":           # RY(t) = S RX(t) Sdg
            out += [("sdg", q, ()), ("h", q, ()), ("rz", q, p), ("h", q, ()), ("s", q, ())]
            phi -= p[0] / 2
        elif name == "cz":
            a, b = q
            out += [("h", (b,), ()), ("cx", (a, b), ()), ("h", (b,), ())]
        elif name == "cp":           # CP(t) = P(t/2)_a CX P(-t/2)_b CX P(t/2)_b
            a, b = q
            t = p[0]
            out += [("rz", (a,), (t / 2,)), ("cx", (a, b), ()), ("rz", (b,), (-t / 2,)), ("cx", (a, b), ()),
                    ("rz", (b,), (t / 2,))]
        elif name == "swap":
            a, b = q
            out += [("cx", (a, b), ()), ("cx", (b, a), ()), ("cx", (a, b), ())]
        elif name == "ccx":          # textbook 6-CX Toffoli, exact with T = diag(1, e^{i pi/4})
            a, b, c = q
            out += [("h", (c,), ()), ("cx", (b, c), ()), ("tdg", (c,), ()), ("cx", (a, c), ()), ("t", (c,), ()),
                    ("cx", (b, c), ()), ("tdg", (c,), ()), ("cx", (a, c), ()), ("t", (b,), ()), ("t", (c,), ()),
                    ("h", (c,), ()), ("cx", (a, b), ()), ("t", (a,), ()), ("tdg", (b,), ()), ("cx", (a, b), ())]
        else:
            raise ValueError(f"no reference spelling for gate {name}")
    return out, phi


def to_reference_qasm(circ, n):
    ref, phi = to_reference_gates(circ)
    return to_qasm(ref, n), phi
