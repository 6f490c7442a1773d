"""torchrun worker: sharded GPU run (one rank per GPU, NCCL exchanges) checked against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
import gpu_quantum_simulator_b200 as q  # noqa: E402
from gpu_quantum_simulator_b200 import circuits, dist as qdist  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    # exchange flavours (qsb_options_t.reserved[5]): 0 = default by world size, 1 = fused peer scatter, 2 = NCCL all-to-all,
    # 3 = pipelined copy-engine exchange
    for prec, tol, mode in ((q.F32, 1e-5, 0), (q.F64, 1e-12, 0), (q.F32, 1e-5, 1), (q.F64, 1e-12, 1), (q.F32, 1e-5, 2), (q.F32, 1e-5, 3), (q.F64, 1e-12, 3)):
        for n, depth, seed in ((22, 8, 7), (23, 5, 8)):
            circ = circuits.random_layered(n, depth=depth, seed=seed)
            sim = q.Simulator(n, precision=prec, rank=rank, world_size=world, device=local, reserved=[0, 0, 0, 0, 0, mode])
            qdist.init_comm(sim, dist)
            st = sim.apply(q.gates_from_circuit(circ))
            got = qdist.gather_state(sim, dist)
            sim.close()
            if rank == 0:
                want = helpers.oracle_run_circuit(circ, n)
                err = float(np.max(np.abs(got - want)))
                worst = max(worst, err / tol)
                print(f"n={n} prec={prec} world={world} exchange_mode={mode} swaps={st['swaps']} passes={st['passes']} err={err:.3e}")
                assert st["swaps"] >= 1
                assert err <= tol, (n, prec, err)
    if rank == 0:
        print(f"sharded_gpu_ok worst_err_over_tol={worst:.3f}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
