"""Fusion / scheduling pass on the CPU: the planner's tables, interpreted by the host test double
(tests/hostcheck), must reproduce the oracle; slot maps must be bijective and bank-conflict free."""
import os

import numpy as np
import pytest

import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits


def run_both(circ, n, precision, low_bits=0):
    gates = q.gates_from_circuit(circ)
    got, rep = helpers.hostcheck_run(gates, n, precision, low_bits)
    want = helpers.oracle_run_circuit(circ, n)
    return got, want, rep


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("n,ng,seed", [(1, 10, 0), (2, 30, 1), (5, 120, 2), (9, 200, 3), (13, 250, 4), (14, 200, 5), (15, 160, 6)])
def test_tables_reproduce_oracle_on_superset_circuits(n, ng, seed, precision):
    circ = circuits.random_superset(n, ng, seed)
    got, want, rep = run_both(circ, n, precision)
    assert rep["bad_slots"] == 0 and rep["noncontig"] == 0
    assert rep["max_conflict"] == 1, "shared-memory exchange must be bank-conflict free"
    assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("precision,low_bits", [(32, 3), (32, 5), (32, 6), (64, 2), (64, 4), (64, 5)])
def test_low_bits_option(precision, low_bits):
    circ = circuits.random_layered(15, depth=4, seed=7)
    got, want, rep = run_both(circ, 15, precision, low_bits)
    assert rep["bad_slots"] == 0 and rep["max_conflict"] == 1
    assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("path", helpers.golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_tables_reproduce_golden_vectors(path):
    circ, n, amps, _ = helpers.load_case(path)
    got, rep = helpers.hostcheck_run(q.gates_from_circuit(circ), n, 32)
    assert rep["bad_slots"] == 0 and rep["max_conflict"] == 1
    assert np.max(np.abs(got - amps)) < 1e-12


def test_qft_and_layered_family():
    for circ, n in ((circuits.qft(14), 14), (circuits.random_layered(16, depth=5, seed=1), 16)):
        got, want, rep = run_both(circ, n, 32)
        assert rep["bad_slots"] == 0 and rep["max_conflict"] == 1
        assert np.max(np.abs(got - want)) < 1e-12


def test_fusion_factor_on_headline_workload():
    """30 q, depth 20: the pass count is what HBM sees; keep it from regressing."""
    circ = circuits.random_layered(30, 20, 12345)
    st = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32)
    assert st["source_gates"] == 900
    # 2^12-amplitude tiles, hill-climbed tiles, lane relocation, h cx h -> cz, rx cx rx -> rx cx
    assert st["passes"] <= 18 and st["rounds"] <= 106 and st["device_ops"] <= 820
    assert st["bytes_moved"] == st["passes"] * 2 * (1 << 30) * 8
    st64 = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=64)
    assert st64["passes"] <= 19 and st64["rounds"] <= 110
    first_come = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32, reserved=[0, 0, 0, 0, 0, 0, 2])
    assert first_come["passes"] >= st["passes"] + 4
    # without the end-of-pass lane relocation: more passes, and empty rounds that only turn the registers
    no_reloc = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32, reserved=[0, 0, 0, 0, 0, 0, 3])
    assert no_reloc["passes"] >= st["passes"] + 1 and no_reloc["rounds"] >= st["rounds"] + 8
    conflict_only = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32, reserved=[0, 0, 0, 0, 0, 0, 4])
    assert st["passes"] <= conflict_only["passes"] <= no_reloc["passes"]
    # CX kept as CX next to an h on its target (reserved[4] = 5): the schedule of call 32
    cx_kept = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32, reserved=[0, 0, 0, 0, 5])
    assert cx_kept["passes"] >= st["passes"] and cx_kept["rounds"] >= st["rounds"] + 10 and cx_kept["device_ops"] >= st["device_ops"] + 50
    # rewritten with an h on ONE side of the CX too (reserved[4] = 6, the schedule of call 33): fewer passes and rounds, but every
    # such CX is now a gate plus a phase instead of one multiplexed gate
    one_sided = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32, reserved=[0, 0, 0, 0, 6])
    assert one_sided["rounds"] < st["rounds"] and one_sided["device_ops"] >= st["device_ops"] + 50
    # both off: the planner of round 1 / the start of round 2 (22 passes, 130 rounds, 11 of them empty)
    old = q.plan_dry_run(30, q.gates_from_circuit(circ), precision=32, reserved=[0, 0, 0, 0, 5, 0, 3])
    assert old["passes"] == 22 and old["rounds"] == 130


def test_cx_next_to_a_hadamard_becomes_a_controlled_phase():
    """h cx h on the target is the reference's spelling of CZ: no matrix op may be left of it, and every
    variant (h before, h after, both, s.h, controls on several qubits) must reproduce the oracle."""
    n = 6
    base = [("h", (k,), ()) for k in range(n)] + [("rx", (k,), (0.3 + k,)) for k in range(n)]
    cases = [
        [("h", (1,), ()), ("cx", (0, 1), ()), ("h", (1,), ())],
        [("cx", (0, 1), ()), ("h", (1,), ())],
        [("h", (1,), ()), ("cx", (0, 1), ())],
        [("rx", (1,), (0.7,)), ("cx", (0, 1), ()), ("h", (1,), ()), ("cx", (2, 1), ()), ("h", (1,), ()), ("cx", (1, 3), ())],
        [("h", (2,), ()), ("ccx", (0, 1, 2), ()), ("h", (2,), ())],
        [("h", (1,), ()), ("s", (1,), ()), ("cx", (0, 1), ()), ("sdg", (1,), ()), ("h", (1,), ())],
        [("cx", (0, 1), ()), ("rz", (0,), (0.4,)), ("x", (0,), ()), ("h", (1,), ()), ("cx", (1, 0), ()), ("h", (0,), ())],
        # gates that commute with X hop over the CX and multiply (rx cx rx -> one multiplexed gate)
        [("rx", (1,), (0.7,)), ("cx", (0, 1), ()), ("rx", (1,), (-1.9,)), ("cx", (2, 1), ()), ("x", (1,), ()), ("ccx", (0, 3, 1), ()), ("rx", (1,), (0.2,))],
    ]
    for prec in (32, 64):
        for tail in cases:
            got, want, rep = run_both(base + tail + base, n, prec)
            assert rep["bad_slots"] == 0 and rep["noncontig"] == 0
            assert np.max(np.abs(got - want)) < 1e-12
            os.environ["QSB_HC_RES4"] = "6"           # the one-sided rewrite (A/B knob) through the same double
            try:
                got, want, rep = run_both(base + tail + base, n, prec)
            finally:
                del os.environ["QSB_HC_RES4"]
            assert rep["bad_slots"] == 0 and np.max(np.abs(got - want)) < 1e-12
    cz = [("h", (1,), ()), ("cx", (0, 1), ()), ("h", (1,), ())]
    st = q.plan_dry_run(4, q.gates_from_circuit(cz))
    assert st["device_ops"] <= 2 and st["rounds"] == 1   # one controlled phase (+ the rounding of h.h as a global scalar)
    kept = q.plan_dry_run(4, q.gates_from_circuit(cz), reserved=[0, 0, 0, 0, 5])
    assert kept["device_ops"] == 3 and kept["rounds"] == 2
    hop = [("rx", (1,), (0.7,)), ("cx", (0, 1), ()), ("rx", (1,), (0.4,))]
    assert q.plan_dry_run(4, q.gates_from_circuit(hop))["device_ops"] < q.plan_dry_run(4, q.gates_from_circuit(hop), reserved=[0, 0, 0, 0, 5])["device_ops"]


def test_doubles_run_the_schedule_the_product_picks():
    """The host doubles call the same plan search as the library (tiled_plan_search: four hill-climbing orders, cheapest kept),
    so what they emulate is what a GPU run executes: same pass and round counts as the library's dry run, and the fixed
    orders the search chooses from are really different schedules."""
    n = 17
    circ = circuits.random_layered(n, depth=8, seed=3)
    gates = q.gates_from_circuit(circ)
    want = helpers.oracle_run_circuit(circ, n)
    for prec in (32, 64):
        st = q.plan_dry_run(n, gates, precision=prec)
        got, rep = helpers.hostcheck_run(gates, n, prec)
        assert (rep["passes"], rep["rounds"]) == (st["passes"], st["rounds"])
        assert np.max(np.abs(got - want)) < 1e-12
    seen = set()
    try:
        for v in range(8):
            helpers.hostcheck_set_climb(v)
            got, rep = helpers.hostcheck_run(gates, n, 32)
            assert np.max(np.abs(got - want)) < 1e-12 and rep["bad_slots"] == 0 and rep["noncontig"] == 0
            seen.add((rep["passes"], rep["rounds"]))
    finally:
        helpers.hostcheck_set_climb(-1)
    assert len(seen) >= 2
    best = min(1.5 * p + 0.25 * r for p, r in seen)
    st = q.plan_dry_run(n, gates, precision=32)
    assert 1.5 * st["passes"] + 0.25 * st["rounds"] <= best + 1e-9 or st["passes"] <= min(p for p, _ in seen) + 1


def test_multi_control_and_global_phase_gates():
    circ = [("h", (0,), ()), ("h", (1,), ()), ("h", (2,), ()), ("ccx", (0, 1, 2), ()), ("y", (1,), ()),
            ("z", (0,), ()), ("sx", (2,), ()), ("cp", (2, 0), (0.3,)), ("cz", (1, 2), ())]
    for prec in (32, 64):
        got, want, rep = run_both(circ, 3, prec)
        assert np.max(np.abs(got - want)) < 1e-13


@pytest.fixture
def device_encoding():
    """Interpret the kernel-parameter blob (what k_tile_pass reads) instead of the planner's logical tables."""
    helpers.hostcheck_use_blob(True)
    yield
    helpers.hostcheck_use_blob(False)


@pytest.mark.parametrize("precision,tol", [(32, 2e-5), (64, 1e-12)])
@pytest.mark.parametrize("n,ng,seed", [(2, 30, 1), (7, 150, 2), (13, 250, 4), (14, 200, 5), (16, 120, 6)])
def test_device_encoding_reproduces_oracle(device_encoding, n, ng, seed, precision, tol):
    """The lowering (groups of slots, specials, thread-phase lists, gather / scatter / smem tables, deferred X)
    interpreted byte for byte as the kernel does; f32 blobs carry float coefficients, hence 2e-5."""
    circ = circuits.random_superset(n, ng, seed)
    got, want, rep = run_both(circ, n, precision)
    assert rep["bad_slots"] == 0
    assert np.max(np.abs(got - want)) < tol


def test_device_encoding_on_the_circuit_families(device_encoding):
    for circ, n in ((circuits.qft(15), 15), (circuits.random_layered(17, depth=5, seed=3), 17),
                    (helpers.load_case(os.path.join(helpers.GOLDEN, "grover_3_18.npz"))[0], 6)):
        for precision, tol in ((32, 2e-5), (64, 1e-12)):
            got, want, rep = run_both(circ, n, precision)
            assert rep["bad_slots"] == 0
            assert np.max(np.abs(got - want)) < tol


def test_device_encoding_of_fused_exchange_passes(device_encoding):
    """Peer-scatter offsets (rank field << 48 | local byte offset) of the fused exchange, all ranks in one process."""
    for n, world, precision, tol in ((20, 4, 32, 2e-5), (19, 2, 64, 1e-12)):
        circ = circuits.random_layered(n, depth=5, seed=2)
        want = helpers.oracle_run_circuit(circ, n)
        for fused in (True, False):
            got, rep = helpers.sharded_host_run(q.gates_from_circuit(circ), n, world, precision, fused=fused)
            assert rep["swaps"] >= 1 and rep["bad_slots"] == 0
            assert np.max(np.abs(got - want)) < tol


def test_fusion_depth_cap_reschedules_without_changing_the_plan_totals():
    circ = circuits.random_layered(30, 20, 12345)
    g = q.gates_from_circuit(circ)
    full = q.plan_dry_run(30, g, precision=32)
    capped = q.plan_dry_run(30, g, precision=32, reserved=[0, 0, 0, 12])
    assert capped["passes"] > 1.5 * full["passes"]
    assert abs(capped["rounds"] - full["rounds"]) <= 0.1 * full["rounds"]      # same SM work, more HBM sweeps


@pytest.mark.parametrize("precision", [32, 64])
def test_edge_case_circuits_on_the_host_double(precision):
    """Empty gate list, diagonal-only, CX-only and global-phase-only circuits; 1-qubit register (padded to a tile)."""
    cases = [
        ([], 3),
        ([("rz", (k,), (0.1 * (k + 1),)) for k in range(5)] + [("cz", (0, 4), ()), ("cp", (1, 3), (0.7,))], 5),
        ([("x", (0,), ())] + [("cx", (k, k + 1), ()) for k in range(13)], 14),
        ([("h", (0,), ()), ("z", (0,), ()), ("s", (0,), ()), ("h", (0,), ())], 1),
    ]
    for circ, n in cases:
        gates = q.gates_from_circuit(circ)
        got, rep = helpers.hostcheck_run(gates, n, precision)
        want = helpers.oracle_run_circuit(circ, n) if circ else np.eye(1, 1 << n, 0, dtype=complex)[0]
        assert rep["bad_slots"] == 0
        assert np.max(np.abs(got - want)) < 1e-12


@pytest.mark.parametrize("precision", [32, 64])
def test_swaps_become_a_relabelling(precision):
    """Three alternating CX on one pair (= `swap`) cost nothing: the two qubits trade wires and the final qubit map
    absorbs the permutation.  Near misses (an op on one of the qubits in between, a repeated direction, a second
    control) must still run as ordinary CX."""
    n = 14
    rng = np.random.RandomState(3)
    circ = circuits.random_layered(n, depth=2, seed=5)
    for _ in range(12):
        a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
        circ.append(("swap", (a, b), ()))
        c = int(rng.randint(n))
        circ.append(("h", (c,), ()))
        circ.append(("rz", (int(rng.randint(n)),), (0.3,)))
    # interleaved with ops on other qubits: still a swap
    circ += [("cx", (0, 1), ()), ("h", (5,), ()), ("cx", (1, 0), ()), ("cp", (6, 7), (0.4,)), ("cx", (0, 1), ())]
    # near misses
    circ += [("cx", (2, 3), ()), ("cx", (3, 2), ()), ("t", (2,), ()), ("cx", (2, 3), ())]
    circ += [("cx", (4, 5), ()), ("cx", (4, 5), ()), ("cx", (5, 4), ())]
    circ += [("cx", (6, 7), ()), ("ccx", (7, 8, 6), ()), ("cx", (6, 7), ())]
    circ += [("cx", (9, 10), ()), ("cx", (10, 9), ()), ("cz", (9, 3), ()), ("cx", (9, 10), ())]
    circ += circuits.random_layered(n, depth=1, seed=6)
    gates = q.gates_from_circuit(circ)
    for blob in (False, True):
        helpers.hostcheck_use_blob(blob)
        got, rep = helpers.hostcheck_run(gates, n, precision)
        want = helpers.oracle_run_circuit(circ, n)
        assert rep["bad_slots"] == 0
        assert np.max(np.abs(got - want)) < (2e-6 if blob and precision == 32 else 1e-12)
    helpers.hostcheck_use_blob(False)
    # a circuit of swaps only needs no pass at all, and the QFT's final bit reversal is free
    only = [("swap", (k, n - 1 - k), ()) for k in range(n // 2)]
    assert q.plan_dry_run(n, q.gates_from_circuit(only), precision=precision)["passes"] == 0
    full = q.plan_dry_run(30, q.gates_from_circuit(circuits.qft(30)), precision=precision)
    bare = q.plan_dry_run(30, q.gates_from_circuit(circuits.qft(30, swaps=False)), precision=precision)
    assert full["passes"] == bare["passes"] and full["rounds"] == bare["rounds"]


@pytest.mark.parametrize("precision", [32, 64])
def test_same_qubit_products(precision):
    """The reference's 2x2 preprocessing (preproces.cu:215-269) in fp64: consecutive one-qubit gates on a qubit
    multiply when the product keeps a cheap form; phases with equal masks multiply; anything touching the qubit
    in between (a control, a phase) stops the product."""
    n = 13
    hh = [("h", (k,), ()) for k in range(n)] * 2
    assert q.plan_dry_run(n, q.gates_from_circuit(hh), precision=precision)["device_ops"] <= 1       # identity (+ global scalar)
    rx = [("rx", (k,), (0.1 * (k + 1) * (d + 1),)) for d in range(5) for k in range(n)]
    one = [("rx", (k,), (0.1 * (k + 1),)) for k in range(n)]
    a, b = (q.plan_dry_run(n, q.gates_from_circuit(c), precision=precision) for c in (rx, one))
    assert a["rounds"] == b["rounds"] and a["passes"] == b["passes"]
    zz = [("rz", (3,), (0.2,)), ("cp", (3, 5), (0.3,)), ("rz", (3,), (0.4,)), ("cp", (5, 3), (0.1,)), ("t", (3,), ())]
    assert q.plan_dry_run(n, q.gates_from_circuit(zz), precision=precision)["device_ops"] == 2         # one phase on q3, one on {q3, q5}
    off = q.plan_dry_run(n, q.gates_from_circuit(rx), precision=precision, reserved=[0, 0, 0, 0, 2])
    assert off["device_ops"] > 3 * a["device_ops"]
    rng = np.random.RandomState(11)
    circ = []
    for _ in range(400):
        t = int(rng.randint(n)); r = rng.rand()
        if r < 0.5:
            circ.append((["h", "x", "y", "sx", "rx", "ry"][int(rng.randint(6))], (t,), ()))
            if circ[-1][0] in ("rx", "ry"):
                circ[-1] = (circ[-1][0], (t,), (float(rng.uniform(-3, 3)),))
        elif r < 0.75:
            circ.append((["z", "s", "t", "tdg", "rz"][int(rng.randint(5))], (t,), ()))
            if circ[-1][0] == "rz":
                circ[-1] = ("rz", (t,), (float(rng.uniform(-3, 3)),))
        else:
            c = int(rng.choice([x for x in range(n) if x != t]))
            circ.append((["cx", "cz", "cp"][int(rng.randint(3))], (c, t), ()))
            if circ[-1][0] == "cp":
                circ[-1] = ("cp", (c, t), (float(rng.uniform(-3, 3)),))
    gates = q.gates_from_circuit(circ)
    want = helpers.oracle_run_circuit(circ, n)
    for blob in (False, True):
        helpers.hostcheck_use_blob(blob)
        got, rep = helpers.hostcheck_run(gates, n, precision)
        assert rep["bad_slots"] == 0
        assert np.max(np.abs(got - want)) < (2e-6 if blob and precision == 32 else 1e-12)
    helpers.hostcheck_use_blob(False)
    fused = q.plan_dry_run(n, gates, precision=precision)
    plain = q.plan_dry_run(n, gates, precision=precision, reserved=[0, 0, 0, 0, 2])
    assert fused["device_ops"] < 0.9 * plain["device_ops"]


def test_sx_runs_in_a_cheap_form():
    """SX = e^{i pi/4} RX(pi/2) (quantum_simulator.c:187): the scalar is pulled out (global factor, or a phase gate on
    the controls), the rest is an rx-form matrix -- no general complex op is left in the plan."""
    n = 13
    circ = [("sx", (k,), ()) for k in range(n)] + [("cx", (0, 5), ()), ("sx", (5,), ()), ("rz", (5,), (0.3,)), ("sx", (5,), ())]
    gates = q.gates_from_circuit(circ)
    # a controlled SX: phase on the control + controlled rx-form
    g = q.Gate()
    g.controls, g.target = 1 << 3, 7
    for k, v in enumerate([0.5, 0.5, 0.5, -0.5, 0.5, -0.5, 0.5, 0.5]):
        g.m[k] = v
    arr = (q.Gate * (len(gates) + 1))(*list(gates), g)
    want = helpers.oracle_run_circuit(circ, n)
    idx = np.arange(1 << n)
    sel = ((idx >> 3) & 1 == 1) & ((idx >> 7) & 1 == 0)
    a, b = want[idx[sel]].copy(), want[idx[sel] | (1 << 7)].copy()
    want[idx[sel]] = (0.5 + 0.5j) * a + (0.5 - 0.5j) * b
    want[idx[sel] | (1 << 7)] = (0.5 - 0.5j) * a + (0.5 + 0.5j) * b
    for precision in (32, 64):
        for blob in (False, True):
            helpers.hostcheck_use_blob(blob)
            got, rep = helpers.hostcheck_run(arr, n, precision)
            assert rep["bad_slots"] == 0
            assert np.max(np.abs(got - want)) < (2e-6 if blob and precision == 32 else 1e-12)
    helpers.hostcheck_use_blob(False)


def test_cx_chains_on_one_target_share_a_round():
    """A deferred X is free but closes its qubit for the round; a chain of CX on one target (the Toffoli
    decompositions of grover_3_18.qasm) would then cost a round per CX.  Chains run their X as matrices instead:
    the shipped Grover circuit needs a third of the rounds, the layered workloads keep their plans."""
    circ, n, amps, _ = helpers.load_case(os.path.join(helpers.GOLDEN, "grover_3_18.npz"))
    st = q.plan_dry_run(n, q.gates_from_circuit(circ), precision=32)
    assert st["source_gates"] == 2445 and st["rounds"] < 450 and st["passes"] < 60        # was 1011 rounds in 140 passes
    got, rep = helpers.hostcheck_run(q.gates_from_circuit(circ), n, 64)
    assert rep["bad_slots"] == 0 and np.max(np.abs(got - amps)) < 1e-12


def test_phase_ladders_onto_the_pack_qubit_are_merged():
    """G_DIAGA + QSB_NVB (f32): controlled phases whose target sits on the PACK qubit (the two lanes of a packed
    register) and whose controls are thread / CTA bits used to cost one multiply of all 16 vectors each (G_DIAG_ALL);
    a run of them is one angle sum + one multiply.  The byte-exact double must see the merged code and still
    reproduce the oracle; f64 tiles have no pack qubit and must never carry it."""
    G_DIAGA, G_DIAG_ALL, NVB = 32, 40, 4
    n = 18
    for circ in (circuits.qft(n),
                 [("h", (q0,), ()) for q0 in range(n)] + [("cp", (c, 0), (0.21 * c + 0.1,)) for c in range(1, n)] + [("h", (0,), ())]):
        want = helpers.oracle_run_circuit(circ, n)
        seen = {}
        for precision in (32, 64):
            helpers.hostcheck_use_blob(True)
            helpers.hostcheck_blob_code_count(0, reset=True)
            try:
                got, rep = helpers.hostcheck_run(q.gates_from_circuit(circ), n, precision)
                seen[precision] = (helpers.hostcheck_blob_code_count(G_DIAGA + NVB), helpers.hostcheck_blob_code_count(G_DIAG_ALL))
            finally:
                helpers.hostcheck_use_blob(False)
            assert rep["bad_slots"] == 0
            assert np.max(np.abs(got - want)) < (3e-6 if precision == 32 else 1e-12)
        assert seen[64] == (0, 0), seen
        assert seen[32][0] > 0, seen          # the ladder onto qubit 0 (the pack qubit of the first pass) is merged


def test_full_outer_condition_table_falls_back():
    """Slots, merged phase runs and angle entries all name outer controls through the pass's 24-entry outer-condition
    table; a gate that finds it full must take its fallback (generic special, plain phase op, thread phase with its
    own 64-bit mask).  Real circuits rarely fill 24 entries, so the host doubles are rebuilt with a 3-entry table and
    run in a child process: QFT, a layered circuit and a ccx / cp mix must still reproduce the oracle byte-exactly
    interpreted, and the table must actually have been full."""
    import subprocess, sys, os
    hc = os.path.join(helpers.ROOT, "tests", "hostcheck")
    subprocess.run(["make", "-C", hc, "SUFFIX=_smallcond", "EXTRA=-DQSB_MAX_COND=3"], check=True, capture_output=True)
    env = dict(os.environ, QSB_HOSTCHECK_SUFFIX="_smallcond")
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "hostcheck_small_table_worker.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("case=")]
    assert len(lines) == 6 and all(l.endswith(" ok") for l in lines), r.stdout
    assert all("max_cond=3 " in l for l in lines), r.stdout                   # the table was full in every case
    qft32 = [l for l in lines if l.startswith("case=qft prec=32")][0]
    assert "merged_runs=0 " not in qft32 and "plain_phase_ops=0 " not in qft32, qft32     # both the merged form and its fallback ran


@pytest.mark.parametrize("precision", [32, 64])
def test_controlled_phase_ladders_are_merged(precision):
    """Round 2 (G_DIAGA): a run of controlled phases on one vector bit (a QFT ladder) is lowered to ONE op -- per-thread
    fixed-point angle sum, one sincospi, one multiply -- instead of one multiply per gate.  The device encoding, read
    byte for byte by the blob double, must still reproduce the oracle, with and without the merge, and the merge must
    remove device ops' worth of specials (fewer bytes of pass descriptor is the visible trace on the host)."""
    n = 18
    circ = circuits.qft(n)
    gates = q.gates_from_circuit(circ)
    want = helpers.oracle_run_circuit(circ, n)
    helpers.hostcheck_use_blob(True)
    try:
        got, rep = helpers.hostcheck_run(gates, n, precision)
        assert rep["bad_slots"] == 0
        assert np.max(np.abs(got - want)) < (3e-6 if precision == 32 else 1e-12)
    finally:
        helpers.hostcheck_use_blob(False)
    # a ladder of cp gates onto one target, controls everywhere: exactly the pattern the merge is for
    ladder = [("h", (0,), ())] + [("cp", (c, 0), (0.37 * c,)) for c in range(1, n)] + [("h", (0,), ())]
    want = helpers.oracle_run_circuit(ladder, n)
    for blob in (False, True):
        helpers.hostcheck_use_blob(blob)
        got, rep = helpers.hostcheck_run(q.gates_from_circuit(ladder), n, precision)
        helpers.hostcheck_use_blob(False)
        assert rep["bad_slots"] == 0
        assert np.max(np.abs(got - want)) < (3e-6 if blob and precision == 32 else 1e-12)


def test_dense_fusion_plans_fewer_sweeps_with_wider_blocks():
    """QSB_MODE_DENSE (the k = 2..5 experiment of BASELINE.json configuration 4): the host fusion needs no GPU.  Wider
    blocks -> fewer sweeps; every sweep is one read + one write of the state; a gate wider than k is refused."""
    circ = circuits.random_layered(24, 20, 12345)
    gates = q.gates_from_circuit(circ)
    sweeps = []
    for k in (1, 2, 3, 4, 5):
        if k == 1:
            with pytest.raises(q.QsbError):
                q.plan_dry_run(24, gates, mode=q.MODE_DENSE, dense_k=1)     # a CX does not fit one qubit
            continue
        st = q.plan_dry_run(24, gates, mode=q.MODE_DENSE, dense_k=k)
        assert st["bytes_moved"] == st["passes"] * 2 * (1 << 24) * 8
        sweeps.append(st["passes"])
    assert sweeps == sorted(sweeps, reverse=True) and sweeps[-1] < 0.5 * sweeps[0]
    tiled = q.plan_dry_run(24, gates)
    assert tiled["passes"] * 4 < sweeps[-1]                                  # the sparse register-tile schedule: far fewer sweeps
