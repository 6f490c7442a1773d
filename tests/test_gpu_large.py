"""Parity at BASELINE sizes (task 1b/1c): full-state oracle diffs at 26 q (f32 and f64) and 28 q (f32),
and a 30 q full-state comparison of the f32 run against the f64 run of the depth-20 headline circuit.

The oracle costs ~5 ns per amplitude and gate on one core, so the 26 / 28 q circuits are a few layers deep (every
qubit still takes a non-diagonal gate and a CX, the tile choice, round schedule and permuted readout are all
exercised); the 30 q test runs the full depth-20 circuit and compares the two precisions chunk by chunk.
Tolerances are BASELINE.json's: max |d amp| <= 1e-5 (f32), <= 1e-12 (f64)."""
import numpy as np
import pytest

import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits, F32, F64

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

_oracle_cache = {}


def oracle_state(n, depth, seed):
    key = (n, depth, seed)
    if key not in _oracle_cache:
        _oracle_cache.clear()                      # one multi-GiB state at a time
        circ = circuits.random_layered(n, depth=depth, seed=seed)
        _oracle_cache[key] = (circ, helpers.oracle_run_circuit(circ, n))
    return _oracle_cache[key]


def max_abs_diff_chunked(sim, want, chunk_bits=24):
    """max |state - want| without holding a second full copy: the state comes down in 2^chunk_bits pieces."""
    n = sim.num_qubits
    worst = 0.0
    step = 1 << min(chunk_bits, n)
    for first in range(0, 1 << n, step):
        got = sim.state(first, step)
        worst = max(worst, float(np.max(np.abs(got - want[first:first + step]))))
    return worst


@pytest.mark.parametrize("precision,tol", [(F32, 1e-5), (F64, 1e-12)], ids=["f32", "f64"])
def test_26q_layered_vs_oracle(precision, tol):
    circ, want = oracle_state(26, 3, 2601)
    with q.Simulator(26, precision=precision) as s:
        st = s.apply(q.gates_from_circuit(circ))
        assert st["passes"] >= 2                   # the circuit does not fit one tile: gathers of high qubits are exercised
        assert max_abs_diff_chunked(s, want) <= tol
        norm, _, _ = s.norm_argmax()
        assert abs(norm - 1.0) < (1e-5 if precision == F32 else 1e-12)


def test_28q_layered_vs_oracle_f32():
    circ, want = oracle_state(28, 1, 2801)
    with q.Simulator(28, precision=F32) as s:
        s.apply(q.gates_from_circuit(circ))
        assert max_abs_diff_chunked(s, want) <= 1e-5


def test_30q_depth20_f32_matches_f64_full_state():
    """The headline single-GPU circuit, every amplitude: the f32 run against the f64 run (whose own parity with the
    oracle is pinned at 26 q above and at <= 22 q in test_gpu_parity.py)."""
    n = 30
    circ = circuits.random_layered(n, 20, 12345)
    g = q.gates_from_circuit(circ)
    step = 1 << 25
    worst, norm32 = 0.0, 0.0
    with q.Simulator(n, precision=F64) as a, q.Simulator(n, precision=F32) as b:
        a.apply(g)
        b.apply(g)
        for first in range(0, 1 << n, step):
            x, y = a.state(first, step), b.state(first, step)
            worst = max(worst, float(np.max(np.abs(x - y))))
            norm32 += float(np.vdot(y, y).real)
        na, _, _ = a.norm_argmax()
    assert worst <= 1e-5, worst
    assert abs(na - 1.0) < 1e-12 and abs(norm32 - 1.0) < 1e-4
