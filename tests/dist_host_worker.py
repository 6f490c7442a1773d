"""world_size-2 CPU run of the sharded schedule over torch.distributed (gloo): each process plans with
the product planner for its own rank, interprets its passes with the host test double and performs the
qubit exchange with isend/irecv -- the same chunk pattern the NCCL path uses on GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
import gpu_quantum_simulator_b200 as q  # noqa: E402
from gpu_quantum_simulator_b200 import circuits  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 19
    circ = circuits.random_layered(n, depth=6, seed=4)
    run = helpers.ShardedHostRun(q.gates_from_circuit(circ), n, world, rank, 32)
    for i in range(run.steps):
        if run.is_swap(i):
            ch = run.chunks()
            send = [torch.from_numpy(ch[j].copy().view(np.float64)) for j in range(world)]
            recv = [torch.empty_like(send[j]) for j in range(world)]
            reqs = []
            for j in range(world):
                if j == rank:
                    recv[j].copy_(send[j])
                    continue
                reqs.append(dist.isend(send[j], j))
                reqs.append(dist.irecv(recv[j], j))
            for r in reqs:
                r.wait()
            for j in range(world):
                ch[j] = recv[j].numpy().view(np.complex128)
        else:
            run.run(i)
    nloc = run.nloc
    shard = torch.from_numpy(run.shard.view(np.float64).copy())
    perm, rep = run.finish()
    gathered = [torch.empty_like(shard) for _ in range(world)]
    dist.all_gather(gathered, shard)
    if rank == 0:
        shards = [g.numpy().view(np.complex128) for g in gathered]
        got = helpers.gather_logical(shards, perm, n, nloc)
        want = helpers.oracle_run_circuit(circ, n)
        print(f"swaps={rep['swaps']} passes={rep['passes']} max_abs_err={np.max(np.abs(got - want)):.3e} ok")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
