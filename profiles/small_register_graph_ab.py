"""A/B on small registers: one plan executed many times with plain stream launches vs a captured CUDA graph
(qsb_options_t.use_graph).  Prints one JSON line per (qubits, mode).  Run on the GPU box:
    python profiles/small_register_graph_ab.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_quantum_simulator_b200 as q  # noqa: E402
from gpu_quantum_simulator_b200 import circuits  # noqa: E402


def main():
    for n in (6, 12, 16, 20, 22):
        circ = circuits.random_layered(n, depth=40, seed=3)
        gates = q.gates_from_circuit(circ)
        for use_graph in (False, True):
            with q.Simulator(n, use_graph=use_graph) as s:
                plan = s.plan(gates)
                for _ in range(5):
                    s.reset()
                    s.execute(plan)
                reps = 200
                dev = 0.0
                t0 = time.perf_counter()
                for _ in range(reps):
                    s.reset()
                    dev += s.execute(plan)["device_ms"]
                wall = (time.perf_counter() - t0) * 1e3
                st = plan.stats()
                print(json.dumps({"qubits": n, "gates": len(circ), "passes": st["passes"], "use_graph": use_graph,
                                  "device_us_per_circuit": round(dev / reps * 1e3, 2),
                                  "host_us_per_execute": round(wall / reps * 1e3, 2),
                                  "gates_per_s_device": round(len(circ) / (dev / reps) * 1e3)}))
                plan.close()


if __name__ == "__main__":
    main()
