"""Child of test_planner.py::test_full_outer_condition_table_falls_back: runs with QSB_HOSTCHECK_SUFFIX naming host doubles
built with a tiny outer-condition table (-DQSB_MAX_COND=k), so that slots, merged phase runs and angle entries all find
the table full and take their fallbacks (generic specials, plain phase ops, 64-bit-mask thread phases).  The device
encoding, read byte for byte by the blob double, must still reproduce the oracle.  Prints one line per case."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits


def main():
    n = 18
    rng = np.random.RandomState(11)
    mixed = [("h", (k,), ()) for k in range(n)]
    for _ in range(2):
        for a in range(n - 1, n - 7, -1):
            for b in range(a - 1, n - 7, -1):
                mixed.append(("ccx", (a, b, int(rng.randint(0, 6))), ()))
        for t in range(0, n, 3):
            mixed += [("cp", (c, t), (0.1 + 0.01 * c + 0.3 * t,)) for c in range(n) if c != t]
        mixed += [("h", (k,), ()) for k in range(0, n, 2)]
    # QFT of a product state (an H layer in front would make the exact answer |0..0>: every error would have to survive a cancellation)
    qft_in = [("rx", (k,), (0.3 + 0.17 * k,)) for k in range(n)] + circuits.qft(n, with_h_layer=False)
    cases = (("qft", qft_in), ("layered", circuits.random_layered(n, depth=6, seed=4)), ("mixed", mixed))
    G_DIAGA, G_DIAG_V = 32, 24
    ok = True
    for name, circ in cases:
        want = helpers.oracle_run_circuit(circ, n)
        for prec in (32, 64):
            helpers.hostcheck_use_blob(True)
            helpers.hostcheck_blob_code_count(0, reset=True); helpers.hostcheck_blob_max_cond(reset=True)
            got, rep = helpers.hostcheck_run(q.gates_from_circuit(circ), n, prec)
            mc = helpers.hostcheck_blob_max_cond()
            runs = sum(helpers.hostcheck_blob_code_count(G_DIAGA + k) for k in range(5))
            plain = sum(helpers.hostcheck_blob_code_count(G_DIAG_V + k) for k in range(4))
            helpers.hostcheck_use_blob(False)
            err = float(np.max(np.abs(got - want)))
            good = rep["bad_slots"] == 0 and err < (3e-6 if prec == 32 else 1e-12)
            ok = ok and good
            print(f"case={name} prec={prec} max_cond={mc} merged_runs={runs} plain_phase_ops={plain} err={err:.3e} bad={rep['bad_slots']} {'ok' if good else 'FAIL'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
