"""Cost probe of k_tile_pass: homogeneous circuits at 28 qubits (f32 and f64), device time per circuit with
passes / rounds / device ops, so that time ~ a * passes + b * rounds + c_kind * gates can be read off.
One JSON line per circuit.  Run on the GPU box:  python profiles/gate_cost_probe.py"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import gpu_quantum_simulator_b200 as q  # noqa: E402
from gpu_quantum_simulator_b200 import circuits  # noqa: E402

N = 28


def layers(kind, depth, rng):
    c = []
    for _ in range(depth):
        if kind == "h":
            c += [("h", (k,), ()) for k in range(N)]
        elif kind == "rx":
            c += [("rx", (k,), (float(rng.uniform(-3, 3)),)) for k in range(N)]
        elif kind == "ry":
            c += [("ry", (k,), (float(rng.uniform(-3, 3)),)) for k in range(N)]
        elif kind == "u3":   # general complex 2x2: h . rz . rx fused by nobody -> three ops; use y-rotation + phase mix
            c += [("rz", (k,), (float(rng.uniform(-3, 3)),)) for k in range(N)]
            c += [("ry", (k,), (float(rng.uniform(-3, 3)),)) for k in range(N)]
        elif kind == "rz":
            c += [("rz", (k,), (float(rng.uniform(-3, 3)),)) for k in range(N)]
        elif kind == "cx":
            p = rng.permutation(N)
            c += [("cx", (int(p[2 * i]), int(p[2 * i + 1])), ()) for i in range(N // 2)]
        elif kind == "cp":
            p = rng.permutation(N)
            c += [("cp", (int(p[2 * i]), int(p[2 * i + 1])), (float(rng.uniform(-3, 3)),)) for i in range(N // 2)]
        elif kind == "h+cx":
            c += [("h", (k,), ()) for k in range(N)]
            p = rng.permutation(N)
            c += [("cx", (int(p[2 * i]), int(p[2 * i + 1])), ()) for i in range(N // 2)]
    return c


def main():
    rng = np.random.RandomState(1)
    cases = [(k, d) for k in ("h", "rx", "ry", "rz", "cx", "cp", "h+cx", "u3") for d in (1, 4)]
    cases.append(("layered", 20))
    for prec in (q.F32, q.F64):
        for kind, depth in cases:
            circ = circuits.random_layered(N, depth, 12345) if kind == "layered" else layers(kind, depth, rng)
            with q.Simulator(N, precision=prec) as s:
                plan = s.plan(q.gates_from_circuit(circ))
                st = plan.stats()
                for _ in range(3):
                    s.reset()
                    s.execute(plan)
                reps, ms = 10, 0.0
                for _ in range(reps):
                    s.reset()
                    ms += s.execute(plan)["device_ms"]
                ms /= reps
                floor = st["passes"] * 2 * (1 << N) * (8 if prec == q.F32 else 16) / 6546.6e9 * 1e3
                print(json.dumps({"qubits": N, "dtype": "f32" if prec == q.F32 else "f64", "kind": kind, "depth": depth,
                                  "gates": len(circ), "passes": st["passes"], "rounds": st["rounds"], "device_ops": st["device_ops"],
                                  "ms": round(ms, 4), "hbm_floor_ms": round(floor, 4), "frac_of_hbm_peak": round(floor / ms, 3) if ms else None}))
                plan.close()


if __name__ == "__main__":
    main()
