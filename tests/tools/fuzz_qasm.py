"""Mutation fuzz of the QASM front end: random edits of valid programs must fail with a message or parse into
gates inside the declared register.  Usage: python tests/tools/fuzz_qasm.py <seed> <iterations>"""
import os
import sys, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits, Gate
lib = q.lib
seed = int(sys.argv[1]); iters = int(sys.argv[2])
rnd = random.Random(seed)
base = [
 'OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[4] q;\nh q[0];\ncx q[0], q[1];\nrz(0.5) q[2];\nsx q[3];\ntdg q[1];\n',
 'OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit q[3];\nh $0;\ncx $0, $1;\nrz(1.25) $2;\n',
 'OPENQASM 3.0;\ninclude "stdgates.inc";\nqubit[3] a;\nqubit[2] b;\ngate foo(t) x, y { rx(t/2) x; cx x, y; p(pi*t) y; }\nfoo(0.3) a[0], b[1];\nctrl @ inv @ foo(1) a[1], a[2], b[0];\npow(3) @ s a;\ngphase(0.1);\nnegctrl(2) @ x a[0], a[1], b[0];\nbarrier a;\nmeasure a;\n',
 '3 4\nh q[0];\ncx q[0], q[1];\nrz(0.5) q[2];\nx q[1];\n',
 circuits.to_qasm(circuits.random_superset(5, 40, 3), 5),
]
tokens = ["q[", "]", "(", ")", ",", ";", "{", "}", "@", "pi", "ctrl", "gate", "qubit", "$", "9999999999", "-", "1e309", "/0", "inv", "pow(", "rz(", "\n", "\r\n", "//", "/*", "\"", "\x00x", "[" * 50, "(" * 50, "a" * 300, "q[-1]", "q[4]", "q[99999999999]"]
crashes = 0
for it in range(iters):
    s = rnd.choice(base)
    for _ in range(rnd.randint(1, 6)):
        k = rnd.random()
        pos = rnd.randrange(len(s) + 1)
        if k < 0.3 and len(s) > 2:
            e = min(len(s), pos + rnd.randint(1, 8)); s = s[:pos] + s[e:]
        elif k < 0.7:
            s = s[:pos] + rnd.choice(tokens) + s[pos:]
        elif k < 0.85 and len(s) > 2:
            s = s[:pos] + chr(rnd.randint(1, 255)) + s[pos + 1:]
        else:
            a = rnd.randrange(len(s) + 1); s = s[:pos] + s[min(a, pos):max(a, pos)] + s[pos:]
    nq = C.c_int(); gp = C.POINTER(Gate)(); n = C.c_size_t()
    data = s.encode("latin-1", "replace").replace(b"\x00", b" ")
    rc = lib.qsb_parse_qasm_string(data, C.byref(nq), C.byref(gp), C.byref(n))
    if rc == 0:
        assert 0 < nq.value <= 64, (nq.value, s)
        for i in range(n.value):
            g = gp[i]
            assert 0 <= g.target < nq.value and not (g.controls >> nq.value) and not (g.controls >> g.target) & 1, (g.target, g.controls, nq.value, s)
        lib.qsb_free(gp)
    else:
        assert lib.qsb_last_error()
print("ok", seed, iters)
