"""Full-depth runs of the UNMODIFIED reference CPU program (oracle/_ref/ref_cexe = /root/reference/quantum_simulator.c,
gcc -O2, single-threaded by construction) on the reference-gate spelling of random_layered(n, depth 20, seed 12345):
BASELINE.md section 3 / SURVEY section 8(d).  One JSON line per size; its own stdout line (gate-loop seconds) is the figure.
usage: python profiles/cpu_reference_full_depth.py 22 24 26"""
import importlib.util, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("qsb_circuits", os.path.join(ROOT, "gpu_quantum_simulator_b200", "circuits.py"))
circuits = importlib.util.module_from_spec(spec); spec.loader.exec_module(circuits)   # no GPU library is loaded
exe = os.path.join(ROOT, "oracle", "_ref", "ref_cexe")
for n in [int(a) for a in sys.argv[1:]] or [22]:
    circ = circuits.random_layered(n, 20, 12345)
    ref, _ = circuits.to_reference_gates(circ)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c.qasm"); open(p, "w").write(circuits.to_qasm(ref, n))
        out = subprocess.run([exe, p, "0"], capture_output=True, text=True, check=True).stdout.split()
    t = float(out[0])
    print(json.dumps({"qubits": n, "source_gates": len(circ), "reference_set_gates": len(ref), "seconds": t,
                      "gates_per_sec": len(circ) / t, "ns_per_amplitude_per_reference_gate": t / len(ref) / (1 << n) * 1e9,
                      "cores_used": 1, "host_cores": os.cpu_count(), "program": "oracle/_ref/ref_cexe (unmodified quantum_simulator.c, gcc -O2)"}), flush=True)
