"""QASM front end (csrc/qasm.c) against the reference grammar (quantum_simulator.c:133-242). CPU only."""
import math
import os

import numpy as np
import pytest

import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits


def mat(g):
    m = np.array(list(g.m)).reshape(4, 2)
    return (m[:, 0] + 1j * m[:, 1]).reshape(2, 2)


def test_bell_both_declaration_styles_and_crlf():
    for text in ('OPENQASM 3.0;\r\ninclude "stdgates.inc";\r\nqubit q[2];\r\nh q[0];\r\ncx q[0], q[1];\r\n',
                 'OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[2] q;\nh $0;\ncx $0, $1;\n'):
        n, g = q.parse_qasm_string(text)
        assert n == 2 and len(g) == 2
        assert g[0].target == 0 and g[0].controls == 0
        assert np.allclose(mat(g[0]), np.array([[1, 1], [1, -1]]) / math.sqrt(2))
        assert g[1].controls == 1 and g[1].target == 1          # first operand is the control (:229-235)
        assert np.array_equal(mat(g[1]), np.array([[0, 1], [1, 0]]))


def test_rz_is_the_phase_gate():
    n, g = q.parse_qasm_string('OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[1] q;\nrz(1.5707963267948966) q[0];\n')
    m = mat(g[0])
    assert m[0, 0] == 1 and m[0, 1] == 0 and m[1, 0] == 0
    assert m[1, 1] == complex(math.cos(1.5707963267948966), math.sin(1.5707963267948966))


def test_reference_gate_table_matches_oracle_table():
    L = helpers.oracle_lib()
    import ctypes as C
    for name in ["x", "sx", "z", "s", "sdg", "t", "tdg", "h", "y"]:
        n, g = q.parse_qasm_string(f'OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[1] q;\n{name} q[0];\n')
        m = (C.c_double * 8)()
        assert L.oc_gate_matrix(name.encode(), 0.0, m) == 0
        assert list(g[0].m) == list(m), name
    for name in ["rz", "rx", "ry", "p"]:
        n, g = q.parse_qasm_string(f'OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[1] q;\n{name}(-0.7321) q[0];\n')
        m = (C.c_double * 8)()
        assert L.oc_gate_matrix(name.encode(), -0.7321, m) == 0
        assert list(g[0].m) == list(m), name


def test_unknown_token_is_an_error_with_reference_wording():
    with pytest.raises(q.QsbError) as e:
        q.parse_qasm_string('OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[2] q;\nfoo q[0];\n')
    assert "Unknown token: foo" in str(e.value)


def test_missing_file():
    with pytest.raises(q.QsbError) as e:
        q.parse_qasm_file("/nonexistent/file.qasm")
    assert "cannot open circuit file" in str(e.value)


def test_operand_out_of_range_and_repeated_operand():
    with pytest.raises(q.QsbError):
        q.parse_qasm_string('OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[2] q;\nh q[2];\n')
    with pytest.raises(q.QsbError):
        q.parse_qasm_string('OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[2] q;\ncx q[1], q[1];\n')


def test_pi_expressions_comments_and_ignored_statements():
    text = ('OPENQASM 3.0;\ninclude "stdgates.inc";\n// comment\nqubit[3] q;\nbit[3] c;\n'
            'rx(pi/2) q[0];\nu(pi/2, 0, -pi / 4) q[1];\nbarrier q[0], q[1];\ncp(2*pi/8) q[0], q[2];\nc = measure q;\n')
    n, g = q.parse_qasm_string(text)
    assert n == 3 and len(g) == 3
    assert np.allclose(mat(g[0]), np.array([[1, -1j], [-1j, 1]]) / math.sqrt(2))
    assert g[2].controls == 1 and g[2].target == 2
    assert np.allclose(mat(g[2])[1, 1], np.exp(1j * math.pi / 4))


def test_cuda_variant_header():
    circ = [("h", (0,), ()), ("cx", (0, 1), ()), ("rz", (1,), (0.25,))]
    n, g = q.parse_qasm_string(circuits.to_cuda_variant_text(circ, 2))
    assert n == 2 and len(g) == 3


def test_swap_and_ccx():
    n, g = q.parse_qasm_string('OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[3] q;\nswap q[0], q[2];\nccx q[0], q[1], q[2];\n')
    assert len(g) == 4
    assert [(x.controls, x.target) for x in g] == [(1, 2), (4, 0), (1, 2), (3, 2)]


@pytest.mark.skipif(not os.path.isdir(helpers.REFERENCE_DIR), reason="/root/reference absent")
def test_shipped_files_parse():
    n, g = q.parse_qasm_file(os.path.join(helpers.REFERENCE_DIR, "grover_3_18.qasm"))
    assert n == 6 and len(g) == 2445
    n, g = q.parse_qasm_file(os.path.join(helpers.REFERENCE_DIR, "entanglement.qasm"))
    assert n == 2 and len(g) == 2


def test_writer_roundtrip_matches_generator():
    circ = circuits.random_superset(7, 120, seed=3)
    n, g = q.parse_qasm_string(circuits.to_qasm(circ, 7))
    g2 = q.gates_from_circuit(circ)
    assert n == 7 and len(g) == len(g2)
    for a, b in zip(g, g2):
        assert a.controls == b.controls and a.target == b.target and list(a.m) == list(b.m)
