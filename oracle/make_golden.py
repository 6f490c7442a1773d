"""Generate tests/golden/*.npz from the UNMODIFIED reference build (oracle/_ref).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Each fixture stores the circuit (native spelling), n and the fp64 amplitudes the reference's
compute_state_vector() produced -- for superset circuits, the reference ran the respelling in
its own gate set and the dropped global phase was re-applied (circuits.to_reference_qasm).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from gpu_quantum_simulator_b200 import circuits  # noqa: E402


def main():
    assert helpers.have_ref(), "build oracle/_ref first (make -C oracle)"
    os.makedirs(helpers.GOLDEN, exist_ok=True)
    cases = []
    for name in ("entanglement", "grover_3_18"):
        path = os.path.join(helpers.REFERENCE_DIR, name + ".qasm")
        circ, n = helpers.parse_reference_style_file(path)
        _, amps = helpers.ref_run_file(path)          # the shipped bytes, CRLF and all
        cases.append((name, circ, n, amps, "reference file, run unmodified"))
    for n, ng, seed in ((1, 12, 1), (3, 60, 2), (5, 150, 3), (8, 300, 4), (10, 400, 5), (13, 300, 6), (14, 200, 7)):
        circ = circuits.random_reference_gates(n, ng, seed)
        cases.append((f"refgates_n{n}_s{seed}", circ, n, helpers.ref_run_circuit(circ, n), "reference gate set"))
    for n, ng, seed in ((4, 80, 11), (7, 200, 12), (11, 300, 13), (14, 250, 14)):
        circ = circuits.random_superset(n, ng, seed)
        cases.append((f"superset_n{n}_s{seed}", circ, n, helpers.ref_run_circuit(circ, n), "superset, respelled for the reference"))
    circ = circuits.qft(9)
    cases.append(("qft_n9", circ, 9, helpers.ref_run_circuit(circ, 9), "QFT after H layer"))
    circ = circuits.random_layered(14, depth=6, seed=12345)
    cases.append(("layered_n14_d6", circ, 14, helpers.ref_run_circuit(circ, 14), "BASELINE random-layered family"))
    for name, circ, n, amps, note in cases:
        helpers.save_case(os.path.join(helpers.GOLDEN, name + ".npz"), circ, n, amps, note)
        print(f"{name}: n={n} gates={len(circ)} norm={abs((amps.conj() * amps).sum()):.15f}")


if __name__ == "__main__":
    main()
