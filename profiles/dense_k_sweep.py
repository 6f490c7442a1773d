"""BASELINE.json configuration 4: 32 q random circuit, depth 20, fusion k = 2..5 sweep (QSB_MODE_DENSE) next to the
sparse register-tile schedule (QSB_MODE_TILED).  One JSON line per configuration: ms per circuit (CUDA events, 3 warm-up
+ 3 timed executions), sweeps, achieved GB/s.  Run under ncu with -k regex:k_dense for sm__throughput vs dram__throughput.
usage: python profiles/dense_k_sweep.py [qubits=32] [steps=3] [ks, e.g. 4,5 -- dense only, for a short ncu run]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
circ = circuits.random_layered(n, 20, 12345)
gates = q.gates_from_circuit(circ)
configs = ((q.MODE_TILED, 0), (q.MODE_DENSE, 2), (q.MODE_DENSE, 3), (q.MODE_DENSE, 4), (q.MODE_DENSE, 5))
if len(sys.argv) > 3:
    configs = tuple((q.MODE_DENSE, int(k)) for k in sys.argv[3].split(","))
warm = 3 if len(sys.argv) <= 3 else 0
for mode, k in configs:
    with q.Simulator(n, precision=q.F32, mode=mode, dense_k=k) as s:
        plan = s.plan(gates)
        ms = []
        for i in range(warm + steps):
            s.reset()
            st = s.execute(plan)
            if i >= warm:
                ms.append(st["device_ms"])
        norm, _, _ = s.norm_argmax()
        t = sum(ms) / len(ms)
        print(json.dumps({"qubits": n, "mode": "tiled" if mode == q.MODE_TILED else f"dense k={k}", "sweeps": st["passes"],
                          "ms_per_circuit": t, "gates_per_sec": len(circ) / (t * 1e-3), "ms_per_sweep": t / st["passes"],
                          "GBps": st["passes"] * 2 * (1 << n) * 8 / (t * 1e-3) / 1e9, "norm": norm}), flush=True)
        plan.close()
