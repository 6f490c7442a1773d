"""One process per GPU: torch.distributed is the plumbing (rendezvous, id broadcast, barriers); the
qubit exchanges themselves run inside libqsim_b200 on NCCL send/recv over NVLink."""
import ctypes as C

import numpy as np

from ._lib import lib, check


def init_comm(sim, dist):
    """Create the NCCL communicator of `sim` (a Simulator built with rank / world_size).  Rank 0 makes
    the ncclUniqueId, torch.distributed broadcasts its 128 bytes."""
    import torch
    buf = (C.c_ubyte * 128)()
    if dist.get_rank() == 0:
        check(lib.qsb_comm_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().numpy().tolist())
    check(lib.qsb_comm_init(sim._h, raw))


def gather_state(sim, dist):
    """Full state in logical order on every rank (test / small-n helper)."""
    import torch
    shard = torch.from_numpy(sim.shard_physical().view(np.float64).copy())
    perm, nloc = sim.layout()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    shard = shard.to(dev)
    parts = [torch.empty_like(shard) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, shard)
    full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
    n = sim.num_qubits
    idx = np.arange(1 << n, dtype=np.uint64)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(int(perm[q]))
    return full[phys]
