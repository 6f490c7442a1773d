"""Parity tests proper: the CUDA path (through the C ABI) against the oracle, the golden vectors
and size-independent properties.  Tolerances are BASELINE.json's: max|d amp| <= 1e-5 (f32),
<= 1e-12 (f64), identical measurement argmax (as a tie-tolerant set)."""
import math
import os
import subprocess

import numpy as np
import pytest

import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits, F32, F64, MODE_TILED, MODE_SWEEP, MODE_DENSE

pytestmark = pytest.mark.gpu

TOL = {F32: 1e-5, F64: 1e-12}


def run_gpu(circ, n, precision, mode=MODE_TILED, low_bits=0):
    with q.Simulator(n, precision=precision, mode=mode, low_bits=low_bits) as s:
        st = s.apply(q.gates_from_circuit(circ))
        return s.state(), st


@pytest.mark.parametrize("mode", [MODE_TILED, MODE_SWEEP], ids=["tiled", "sweep"])
@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
@pytest.mark.parametrize("path", helpers.golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_golden_vectors(path, precision, mode):
    circ, n, amps, _ = helpers.load_case(path)
    got, _ = run_gpu(circ, n, precision, mode)
    assert np.max(np.abs(got - amps)) <= TOL[precision]


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
def test_grover_argmax_set_and_probabilities(precision):
    circ, n, amps, _ = helpers.load_case(os.path.join(helpers.GOLDEN, "grover_3_18.npz"))
    with q.Simulator(n, precision=precision) as s:
        s.apply(q.gates_from_circuit(circ))
        norm, idx, p = s.norm_argmax()
        probs = s.probabilities()
    assert idx in (3, 18)                      # p[3] - p[18] = 1.4e-15 in the reference
    assert abs(p - 0.49959115777166263) < 1e-5
    assert abs(norm - 1.0) < (1e-4 if precision == F32 else 1e-12)   # 2445 sequential f32 gates
    assert set(np.argsort(probs)[-2:]) == {3, 18}


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
@pytest.mark.parametrize("n,ng,seed", [(16, 300, 21), (18, 400, 22), (20, 300, 23), (22, 150, 24)])
def test_random_superset_vs_oracle(n, ng, seed, precision):
    circ = circuits.random_superset(n, ng, seed)
    want = helpers.oracle_run_circuit(circ, n)
    got, st = run_gpu(circ, n, precision)
    assert np.max(np.abs(got - want)) <= TOL[precision]
    assert st["passes"] < ng


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
def test_layered_family_and_qft_vs_oracle(precision):
    for circ, n in ((circuits.random_layered(22, depth=8, seed=12345), 22), (circuits.qft(20), 20)):
        want = helpers.oracle_run_circuit(circ, n)
        got, _ = run_gpu(circ, n, precision)
        assert np.max(np.abs(got - want)) <= TOL[precision]


@pytest.mark.parametrize("low_bits", [3, 5, 6])
def test_low_bits_variants(low_bits):
    circ = circuits.random_layered(18, depth=6, seed=3)
    want = helpers.oracle_run_circuit(circ, 18)
    got, _ = run_gpu(circ, 18, F32, low_bits=low_bits)
    assert np.max(np.abs(got - want)) <= TOL[F32]


def test_tiled_equals_sweep_and_f32_close_to_f64_at_26_qubits():
    """Sizes the oracle cannot reach in seconds: fused vs unfused schedule, f32 vs f64, norm."""
    n = 26
    circ = circuits.random_layered(n, depth=6, seed=99)
    gates = q.gates_from_circuit(circ)
    with q.Simulator(n, precision=F64) as s:
        s.apply(gates)
        ref = s.state_native().copy()
        norm64 = s.norm_argmax()[0]
    with q.Simulator(n, precision=F64, mode=MODE_SWEEP) as s:
        s.apply(gates)
        sw = s.state_native()
        assert np.max(np.abs(sw - ref)) <= 1e-12
    with q.Simulator(n, precision=F32) as s:
        s.apply(gates)
        f32 = s.state_native()
        norm32 = s.norm_argmax()[0]
        assert np.max(np.abs(f32.astype(np.float64) - ref)) <= 1e-5
    assert abs(norm64 - 1.0) < 1e-12 and abs(norm32 - 1.0) < 1e-4


def test_unitarity_roundtrip_30_qubits():
    """BASELINE size: circuit followed by its inverse returns |0...0> (f32)."""
    n = 30
    circ = circuits.random_layered(n, depth=4, seed=5)
    inv = []
    for name, qs, p in reversed(circ):
        inv.append((name, qs, tuple(-x for x in p)) if name in ("rx", "rz") else (name, qs, p))
    with q.Simulator(n, precision=F32) as s:
        s.apply(q.gates_from_circuit(circ))
        norm, _, _ = s.norm_argmax()
        assert abs(norm - 1.0) < 1e-4
        s.apply(q.gates_from_circuit(inv))
        norm, idx, p = s.norm_argmax()
        assert idx == 0 and abs(p - 1.0) < 1e-4
        head = s.state(0, 8)
        assert abs(abs(head[0]) - 1.0) < 1e-4 and np.max(np.abs(head[1:])) < 1e-5


def test_reference_named_operations_and_cdf():
    n = 10
    circ = circuits.random_reference_gates(n, 200, seed=8)
    want = helpers.oracle_run_circuit(circ, n)
    L = helpers.oracle_lib()
    cdf_want = np.zeros(1 << n)
    L.oc_cdf(want.ctypes.data, n, cdf_want.ctypes.data)
    with q.Simulator(n, precision=F64) as s:
        s.set_state(want)
        cdf = s.compute_state_cumulative_distribution()
        assert np.max(np.abs(cdf - cdf_want)) < 1e-12
        shots = s.measurement(64, seed=7)
        assert all(cdf[k] > 0 for k in shots)
        # one reference-style call at a time
        s.reset()
        h = np.array([1, 1, 1, -1]) / math.sqrt(2)
        s.execute_single_qubit_gate(h, 0)
        s.execute_cnot(0, 1)
        v = s.state()
        assert abs(v[0] - 1 / math.sqrt(2)) < 1e-15 and abs(v[3] - 1 / math.sqrt(2)) < 1e-15


def test_ref_shim_and_cli(tmp_path):
    import ctypes as C
    circ, n, amps, _ = helpers.load_case(os.path.join(helpers.GOLDEN, "grover_3_18.npz"))
    path = tmp_path / "grover.qasm"
    path.write_bytes(circuits.to_qasm(circ, n).replace("\n", "\r\n").encode())   # CRLF like the shipped file
    nq = C.c_int()
    p = q.lib.qsb_ref_compute_state_vector(str(path).encode(), C.byref(nq))
    assert p and nq.value == n
    got = np.ctypeslib.as_array(p, shape=(2 << n,)).copy().view(np.complex128)
    q.lib.qsb_free(p)
    assert np.max(np.abs(got - amps)) <= 1e-12
    exe = os.path.join(helpers.ROOT, "gpu_quantum_simulator_b200", "bin", "qsim")
    out = subprocess.run([exe, str(path), "0", "--precision", "64", "--dump-amplitudes", "--precision-out", "17"],
                         capture_output=True, text=True, check=True).stdout.split("\n")
    float(out[0])                                             # line 1: "%lf" seconds, like the reference
    vals = {}
    for line in out[1:]:
        if " : " in line:
            k, rest = line.split(" : ")
            re_, im_ = rest.replace(" i", "").split(" + ")
            vals[int(k)] = complex(float(re_), float(im_))
    assert max(abs(vals.get(k, 0) - amps[k]) for k in range(1 << n)) <= 1e-12
    assert any(l.startswith("MOST LIKELY MEASUREMENT: ") for l in out)
    bad = subprocess.run([exe], capture_output=True, text=True)
    assert bad.returncode == 1 and "Usage:" in bad.stdout


def test_reference_main_links_against_gpu_library(tmp_path):
    """the UNMODIFIED reference program (oracle/_ref/libqsim_ref.so = quantum_simulator.c
    built -fPIC -Dmain=ref_main) runs its own main() while compute_state_vector / the CDF resolve to
    libqsim_b200_refcompat.so, i.e. run on the GPU."""
    ref_so = os.path.join(helpers.ROOT, "oracle", "_ref", "libqsim_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    lib_dir = os.path.join(helpers.ROOT, "gpu_quantum_simulator_b200")
    launcher = tmp_path / "launcher.c"
    launcher.write_text("int ref_main(int, char **); int main(int c, char **v) { return ref_main(c, v); }\n")
    exe = tmp_path / "CExe_gpu"
    subprocess.run(["gcc", str(launcher), "-o", str(exe), "-Wl,--no-as-needed", "-L" + lib_dir, "-lqsim_b200_refcompat",
                    "-L" + os.path.dirname(ref_so), "-lqsim_ref", "-lqsim_b200", "-lm",
                    "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.dirname(ref_so)], check=True)
    circ, n, amps, _ = helpers.load_case(os.path.join(helpers.GOLDEN, "grover_3_18.npz"))
    path = tmp_path / "grover.qasm"
    path.write_text(circuits.to_qasm(circ, n))
    env = dict(os.environ, LD_DEBUG="bindings")
    r = subprocess.run([str(exe), str(path), "1"], capture_output=True, text=True, env=env)
    assert r.returncode == 0
    float(r.stdout.split()[0])                                  # the reference's only output: "%lf" seconds
    bound = [l for l in r.stderr.splitlines() if "`compute_state_vector'" in l and "normal symbol" in l]
    assert bound and all("libqsim_b200_refcompat" in l.split(" to ")[1] for l in bound), bound[:3]
    bad = subprocess.run([str(exe)], capture_output=True, text=True)
    assert bad.returncode == 1 and "Usage:" in bad.stdout          # quantum_simulator.c:39-43


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
def test_cx_heavy_circuits_vs_oracle(precision):
    """CX / X chains exercise the deferred swap (S_XDEF), its merge into the preceding slot and the
    closed-qubit rule of the round builder."""
    rng = np.random.RandomState(5)
    for n in (17, 21):
        circ = []
        for _ in range(6 * n):
            k = rng.randint(0, 5)
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            if k == 0:
                circ.append(("h", (a,), ()))
            elif k == 1:
                circ.append(("x", (a,), ()))
            elif k == 2:
                circ.append(("rx", (a,), (float(rng.uniform(-3, 3)),)))
            else:
                circ.append(("cx", (a, b), ()))
        want = helpers.oracle_run_circuit(circ, n)
        got, _ = run_gpu(circ, n, precision)
        assert np.max(np.abs(got - want)) <= TOL[precision]


def test_planner_knobs_do_not_change_the_result():
    """Fusion depth cap, lazy diagonals, phase sinking, tail trimming, 2x2 products and the tile choice only
    reschedule: same state (f64, 1e-12)."""
    n = 24
    circ = circuits.random_layered(n, depth=8, seed=11) + circuits.qft(n, with_h_layer=False, swaps=False)[:200]
    gates = q.gates_from_circuit(circ)
    ref = None
    for reserved in (None, [0, 0, 0, 12], [0, 2, 1, 0], [0, 0, 0, 0, 1, 0, 1], [0, 0, 3, 20, 0, 0, 0], [0, 0, 0, 0, 2, 0, 2]):
        with q.Simulator(n, precision=F64, reserved=reserved) as s:
            st = s.apply(gates)
            v = s.state_native().copy()
        if ref is None:
            ref, ref_passes = v, st["passes"]
        else:
            assert np.max(np.abs(v - ref)) <= 1e-12, reserved
    with q.Simulator(n, precision=F64, reserved=[0, 0, 0, 12]) as s:
        assert s.apply(gates)["passes"] > ref_passes          # the cap really produced a different schedule


def test_qft_of_zero_state_is_uniform_at_32_qubits():
    """BASELINE size property: QFT|0...0> = uniform superposition, every amplitude 2^-16 (f32, 32 GiB state)."""
    n = 32
    circ = circuits.qft(n, with_h_layer=False, swaps=True)
    with q.Simulator(n, precision=F32) as s:
        st = s.apply(q.gates_from_circuit(circ))
        norm, idx, p = s.norm_argmax()
        head = s.state(0, 1 << 12)
        tail = s.state((1 << n) - (1 << 12), 1 << 12)
    amp = 2.0 ** (-n / 2)
    assert abs(norm - 1.0) < 1e-4
    assert abs(p - amp * amp) < 1e-4 * amp * amp * 100
    for part in (head, tail):
        assert np.max(np.abs(part - amp)) < 1e-5 * amp * 10 + 1e-9
    assert st["passes"] < len(circ) // 20


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
def test_edge_case_circuits(precision):
    """Empty gate list, diagonal-only, CX-only chains, a 1-qubit register and a global phase (reference has no such tests;
    these are the degenerate inputs of its grammar)."""
    with q.Simulator(3, precision=precision) as s:
        st = s.apply(q.gates_from_circuit([]))
        v = s.state()
        assert st["passes"] == 0 and v[0] == 1 and np.count_nonzero(v) == 1
    cases = [
        ([("h", (k,), ()) for k in range(5)] + [("rz", (k,), (0.1 * (k + 1),)) for k in range(5)] + [("cz", (0, 4), ()), ("cp", (1, 3), (0.7,))], 5),
        ([("x", (0,), ())] + [("cx", (k, k + 1), ()) for k in range(19)], 20),
        ([("h", (0,), ()), ("z", (0,), ()), ("s", (0,), ()), ("h", (0,), ())], 1),
    ]
    for circ, n in cases:
        want = helpers.oracle_run_circuit(circ, n)
        got, _ = run_gpu(circ, n, precision)
        assert np.max(np.abs(got - want)) <= TOL[precision]
    nq, g = q.parse_qasm_string('OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[2] q;\nh q[0];\ngphase(0.5);\nctrl @ gphase(0.25) q[1];\n')
    with q.Simulator(nq, precision=precision) as s:
        s.apply(g)
        want = np.exp(0.5j) * np.array([1, 1, 0, 0]) / math.sqrt(2)
        assert np.max(np.abs(s.state() - want)) <= TOL[precision]


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
def test_graph_replay_of_a_plan_equals_plain_launches(precision):
    """qsb_options_t.use_graph: the first execute captures the pass launches, later ones replay the graph;
    results must be bit-identical to plain stream launches and correct against the oracle."""
    n = 15
    circ = circuits.random_layered(n, depth=12, seed=21)
    gates = q.gates_from_circuit(circ)
    want = helpers.oracle_run_circuit(circ, n)
    with q.Simulator(n, precision=precision) as s:
        plan = s.plan(gates)
        s.execute(plan)
        plain = s.state()
        plan.close()
    with q.Simulator(n, precision=precision, use_graph=True) as s:
        plan = s.plan(gates)
        for _ in range(3):                      # capture + launch, then two replays
            s.reset()
            st = s.execute(plan)
            assert np.array_equal(s.state(), plain)
            assert st["passes"] >= 2 and st["device_ms"] > 0
        plan.close()
        s.reset()
        s.apply(gates)                          # plan + capture + execute + free in one call
        assert np.array_equal(s.state(), plain)
    assert np.max(np.abs(plain - want)) <= TOL[precision]


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
@pytest.mark.parametrize("n", [9, 18, 21])
def test_phase_ladders_and_swaps(n, precision):
    """Rounds with many thread-level phases run them as fixed-point angles (one sincospi per round); `swap`
    (three alternating CX) is a relabelling of the final qubit map.  Random and dyadic angles, f32 and f64."""
    rng = np.random.RandomState(40 + n)
    circ = [("h", (k,), ()) for k in range(n)]
    for _ in range(6 * n):
        a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
        r = rng.rand()
        if r < 0.45:
            circ.append(("cp", (a, b), (float(rng.uniform(-7, 7)),)))
        elif r < 0.7:
            circ.append(("cp", (a, b), (math.pi / 2 ** int(rng.randint(0, 14)),)))
        elif r < 0.8:
            circ.append(("rz", (a,), (float(rng.uniform(-7, 7)),)))
        elif r < 0.87:
            circ.append((["t", "sdg", "z"][int(rng.randint(3))], (a,), ()))
        elif r < 0.9:
            circ.append(("cz", (a, b), ()))
        else:
            circ.append(("swap", (a, b), ()))
    circ += [("h", (k,), ()) for k in range(n)] + circuits.qft(n)
    want = helpers.oracle_run_circuit(circ, n)
    got, st = run_gpu(circ, n, precision)
    assert np.max(np.abs(got - want)) <= TOL[precision]
    sweep, _ = run_gpu(circ, n, precision, mode=q.MODE_SWEEP)
    assert np.max(np.abs(got - sweep)) <= 2 * TOL[precision]


def test_a_plan_is_refused_from_another_layout():
    """Fused passes leave the qubits permuted; executing a plan from a layout it was not made for would compute on
    the wrong bits, so it is an error (qsim_b200.h).  After a reset the same plan runs again, bit-identically."""
    n = 16
    gates = q.gates_from_circuit(circuits.random_layered(n, depth=6, seed=9))
    with q.Simulator(n) as s:
        plan = s.plan(gates)
        s.execute(plan)
        first = s.state()
        perm, _ = s.layout()
        if not np.array_equal(perm[:n], np.arange(n)):
            with pytest.raises(q.QsbError) as e:
                s.execute(plan)
            assert "another qubit layout" in str(e.value)
            assert np.array_equal(s.state(), first)          # nothing ran
        s.reset()
        s.execute(plan)
        assert np.array_equal(s.state(), first)
        s.apply(gates)                                       # plans from the current layout: always fine
        plan.close()


@pytest.mark.parametrize("precision", [F32, F64], ids=["f32", "f64"])
@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_dense_block_mode_vs_oracle(k, precision):
    """QSB_MODE_DENSE (the k = 2..5 fusion experiment of BASELINE.json configuration 4): greedy dense k-qubit blocks,
    one sweep per block, must give the oracle's state -- otherwise the k sweep in profiles/ compares nothing."""
    for circ, n in ((circuits.random_layered(18, depth=6, seed=40 + k), 18), (circuits.random_superset(16, 200, 50 + k), 16),
                    (circuits.qft(15), 15)):
        if k < 3 and any(name == "ccx" for name, _, _ in circ):
            circ = [g for g in circ if g[0] != "ccx"]          # a Toffoli is wider than k = 2
        want = helpers.oracle_run_circuit(circ, n)
        with q.Simulator(n, precision=precision, mode=MODE_DENSE, dense_k=k) as s:
            st = s.apply(q.gates_from_circuit(circ))
            assert np.max(np.abs(s.state() - want)) <= TOL[precision]
            assert st["passes"] < len(circ)
