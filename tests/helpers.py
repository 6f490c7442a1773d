"""Test-side access to the CPU checker (oracle/) and the host emulator (tests/hostcheck/).

Only tests, smoke() and bench.py's CPU-baseline legs import this; the product package never does.
"""
import ctypes as C
import math
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libqsim_ref.so")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "ref_cexe")
# QSB_HOSTCHECK_SUFFIX: a variant of the host doubles built by the caller (make -C tests/hostcheck SUFFIX=... EXTRA=...)
_HC_SUFFIX = os.environ.get("QSB_HOSTCHECK_SUFFIX", os.environ.get("QSB_LIB_SUFFIX", ""))
HOSTCHECK_SO = os.path.join(ROOT, "tests", "hostcheck", "libqsb_hostcheck%s.so" % _HC_SUFFIX)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_DIR = "/root/reference"


_made = set()


def _build(path, directory, always=False):
    """always: let make decide (once per process) -- the host test double links the product planner, so a stale build
    would silently test yesterday's planner."""
    if not os.path.exists(path) or (always and directory not in _made and not (_HC_SUFFIX and path == HOSTCHECK_SO)):
        subprocess.run(["make", "-C", directory], check=True, capture_output=True)
    _made.add(directory)
    return path


def oracle_lib():
    L = C.CDLL(_build(ORACLE_SO, os.path.join(ROOT, "oracle")))
    L.oc_apply_1q.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_uint64]
    L.oc_apply_cx.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int]
    L.oc_gate_matrix.argtypes = [C.c_char_p, C.c_double, C.POINTER(C.c_double)]
    L.oc_gate_matrix.restype = C.c_int
    L.oc_run_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.oc_run_file.restype = C.c_int
    L.oc_cdf.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.oc_measure.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.oc_measure.restype = C.c_uint64
    L.oc_free.argtypes = [C.c_void_p]
    return L


def oracle_run_file(path):
    """Oracle restatement on a QASM file -> (n, complex128 state)."""
    L = oracle_lib()
    p, n = C.c_void_p(), C.c_int()
    rc = L.oc_run_file(str(path).encode(), C.byref(p), C.byref(n))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    N = 1 << n.value
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * N,)).copy()
    L.oc_free(p)
    return n.value, out.view(np.complex128)


def oracle_run_circuit(circ, n, state=None):
    """Oracle restatement on an in-memory circuit (superset gates) -> complex128 state."""
    L = oracle_lib()
    N = 1 << n
    if state is None:
        v = np.zeros(N, dtype=np.complex128)
        v[0] = 1.0
    else:
        v = np.array(state, dtype=np.complex128)
    m = (C.c_double * 8)()
    ptr = v.ctypes.data
    for name, q, p in circ:
        arg = p[0] if p else 0.0
        if name == "cx":
            L.oc_apply_cx(ptr, n, 1 << q[0], q[1])
        elif name == "ccx":
            L.oc_apply_cx(ptr, n, (1 << q[0]) | (1 << q[1]), q[2])
        elif name == "x":
            L.oc_apply_cx(ptr, n, 0, q[0])
        elif name == "swap":
            L.oc_apply_cx(ptr, n, 1 << q[0], q[1]); L.oc_apply_cx(ptr, n, 1 << q[1], q[0]); L.oc_apply_cx(ptr, n, 1 << q[0], q[1])
        elif name in ("cz", "cp"):
            assert L.oc_gate_matrix(b"z" if name == "cz" else b"p", arg, m) == 0
            L.oc_apply_1q(ptr, n, m, q[1], 1 << q[0])
        else:
            assert L.oc_gate_matrix(name.encode(), arg, m) == 0, name
            L.oc_apply_1q(ptr, n, m, q[0], 0)
    return v


def have_ref():
    return os.path.exists(REF_SO)


def ref_run_file(path):
    """The UNMODIFIED reference compute_state_vector (oracle/_ref build) -> (n, complex128 state)."""
    L = C.CDLL(REF_SO)
    L.compute_state_vector.restype = C.c_void_p
    L.compute_state_vector.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    n = C.c_int()
    p = L.compute_state_vector(str(path).encode(), C.byref(n))
    if not p:
        raise RuntimeError("reference returned NULL")
    N = 1 << n.value
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * N,)).copy()
    C.CDLL(None).free(C.c_void_p(p))
    return n.value, out.view(np.complex128)


def ref_run_circuit(circ, n):
    """Reference program on a circuit spelled in its own gate set; returns the state INCLUDING the
    global phase the respelling dropped, i.e. directly comparable with the native circuit."""
    from gpu_quantum_simulator_b200 import circuits
    text, phi = circuits.to_reference_qasm(circ, n)
    with tempfile.NamedTemporaryFile("w", suffix=".qasm", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        _, v = ref_run_file(path)
    finally:
        os.unlink(path)
    return v * complex(math.cos(phi), math.sin(phi))


def hostcheck_use_blob(on):
    """Switch the host test double between the planner's logical tables (False) and the DEVICE ENCODING, i.e. the
    kernel-parameter blob exactly as k_tile_pass reads it (True)."""
    L = C.CDLL(_build(HOSTCHECK_SO, os.path.join(ROOT, "tests", "hostcheck"), always=True))
    L.qsb_hostcheck_use_blob(1 if on else 0)


def hostcheck_set_climb(variant):
    """-1 (the default): the double runs the schedule tiled_plan_search picks -- what a GPU run executes (four orders of the
    tile hill climbing on one GPU, exchange threshold x lane policy when sharded, cheapest kept).  0..7: one fixed order of
    the hill climbing's swaps (tiled_schedule's climb_variant) without the search."""
    L = C.CDLL(_build(HOSTCHECK_SO, os.path.join(ROOT, "tests", "hostcheck"), always=True))
    L.qsb_hostcheck_set_climb(int(variant))


def hostcheck_blob_code_count(code, reset=False):
    """How many special ops with this code (tiled.h G_*) the blob double has interpreted since the last reset."""
    L = C.CDLL(_build(HOSTCHECK_SO, os.path.join(ROOT, "tests", "hostcheck"), always=True))
    L.qsb_hostcheck_blob_code_count.restype = C.c_ulong
    return int(L.qsb_hostcheck_blob_code_count(int(code), 1 if reset else 0))


def hostcheck_blob_max_cond(reset=False):
    """Largest outer-condition table among the passes the blob double has interpreted since the last reset."""
    L = C.CDLL(_build(HOSTCHECK_SO, os.path.join(ROOT, "tests", "hostcheck"), always=True))
    L.qsb_hostcheck_blob_max_cond.restype = C.c_uint
    return int(L.qsb_hostcheck_blob_max_cond(1 if reset else 0))


def hostcheck_run(circ_gates, n, precision=32, low_bits=0, state=None):
    """Schedule with the product planner, interpret the tables on the host (tests/hostcheck).
    -> (complex128 state in LOGICAL order, report dict)"""
    from gpu_quantum_simulator_b200 import Gate
    L = C.CDLL(_build(HOSTCHECK_SO, os.path.join(ROOT, "tests", "hostcheck"), always=True))
    L.qsb_hostcheck_run.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(Gate), C.c_size_t, C.c_void_p,
                                    C.POINTER(C.c_int), C.c_void_p]
    L.qsb_hostcheck_run.restype = C.c_int
    T = L.qsb_hostcheck_tile_bits(precision)        # the tile geometry is a build-time switch (tiled.h QSB_TB)
    nloc = max(n, T)
    v = np.zeros(1 << nloc, dtype=np.complex128)
    if state is None:
        v[0] = 1.0
    else:
        v[: 1 << n] = state
    rep = (C.c_int * 5)()
    perm = np.zeros(64, dtype=np.int8)
    rc = L.qsb_hostcheck_run(n, precision, low_bits, circ_gates, len(circ_gates), v.ctypes.data, rep, perm.ctypes.data)
    if rc:
        raise RuntimeError(f"hostcheck rc={rc}")
    # physical -> logical
    idx = np.arange(1 << n, dtype=np.uint64)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(int(perm[q]))
    out = v[phys]
    report = dict(max_conflict=rep[0], bad_slots=rep[1], noncontig=rep[2], passes=rep[3], rounds=rep[4])
    return out, report


# ---- golden fixtures -------------------------------------------------------------------------
def save_case(path, circ, n, amps, note=""):
    names = np.array([c[0] for c in circ], dtype="U8")
    qs = -np.ones((len(circ), 3), dtype=np.int16)
    ps = np.full(len(circ), np.nan)
    for i, (_, q, p) in enumerate(circ):
        qs[i, : len(q)] = q
        if p:
            ps[i] = p[0]
    np.savez_compressed(path, names=names, qubits=qs, params=ps, n=n, amps=amps, note=note)


def load_case(path):
    z = np.load(path, allow_pickle=False)
    circ = []
    for name, q, p in zip(z["names"], z["qubits"], z["params"]):
        circ.append((str(name), tuple(int(x) for x in q if x >= 0), () if np.isnan(p) else (float(p),)))
    return circ, int(z["n"]), z["amps"], str(z["note"])


def golden_cases():
    if not os.path.isdir(GOLDEN):
        return []
    return sorted(os.path.join(GOLDEN, f) for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def parse_reference_style_file(path):
    """Tiny independent reader for the two shipped circuits (used only to build fixtures)."""
    circ, n = [], None
    import re
    for line in open(path, "rb").read().decode().replace("\r", "").split("\n"):
        line = line.strip()
        if not line or line.startswith("OPENQASM") or line.startswith("include"):
            continue
        if line.startswith("qubit"):
            n = int(re.search(r"\[(\d+)\]", line).group(1))
            continue
        m = re.match(r"([a-z]+)(?:\(([^)]*)\))?\s+(.*);", line)
        name, arg, ops = m.group(1), m.group(2), m.group(3)
        q = tuple(int(x) for x in re.findall(r"\[(\d+)\]", ops))
        circ.append((name, q, (float(arg),) if arg else ()))
    return circ, n


# ---- sharded schedules on the host -------------------------------------------------------------
def _hc():
    from gpu_quantum_simulator_b200 import Gate
    L = C.CDLL(_build(HOSTCHECK_SO, os.path.join(ROOT, "tests", "hostcheck"), always=True))
    L.qsb_hostcheck_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Gate), C.c_size_t]
    L.qsb_hostcheck_plan.restype = C.c_void_p
    L.qsb_hostcheck_plan_fused.argtypes = L.qsb_hostcheck_plan.argtypes
    L.qsb_hostcheck_plan_fused.restype = C.c_void_p
    L.qsb_hostcheck_step_kind.argtypes = [C.c_void_p, C.c_int]
    L.qsb_hostcheck_run_step_fused.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.qsb_hostcheck_num_steps.argtypes = [C.c_void_p]
    L.qsb_hostcheck_step_is_swap.argtypes = [C.c_void_p, C.c_int]
    L.qsb_hostcheck_nloc.argtypes = [C.c_void_p]
    L.qsb_hostcheck_run_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.qsb_hostcheck_finish.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
    L.qsb_hostcheck_step_props.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    return L


class ShardedHostRun:
    """One rank of a sharded schedule, interpreted on the host.  Exchanges are the caller's job."""

    def __init__(self, gates, n, world, rank, precision=32, low_bits=0, swap_min_ops=0, fused=False):
        self.L = _hc()
        plan = self.L.qsb_hostcheck_plan_fused if fused else self.L.qsb_hostcheck_plan
        self.h = plan(n, precision, low_bits, world, rank, swap_min_ops, gates, len(gates))
        if not self.h:
            raise RuntimeError("hostcheck planning failed")
        self.n, self.world, self.rank = n, world, rank
        self.nloc = self.L.qsb_hostcheck_nloc(self.h)
        self.shard = np.zeros(1 << self.nloc, dtype=np.complex128)
        if rank == 0:
            self.shard[0] = 1.0
        self.steps = self.L.qsb_hostcheck_num_steps(self.h)

    def is_swap(self, i):
        return bool(self.L.qsb_hostcheck_step_is_swap(self.h, i))

    def kind(self, i):
        """0 ordinary pass, 1 exchange marker, 2 fused-exchange pass"""
        return self.L.qsb_hostcheck_step_kind(self.h, i)

    def run(self, i):
        assert self.L.qsb_hostcheck_run_step(self.h, i, self.shard.ctypes.data) == 0

    def run_fused(self, i, new_shards):
        """new_shards: list of the P destination shards (complex128 arrays), written in place"""
        ptrs = (C.c_void_p * len(new_shards))(*[a.ctypes.data for a in new_shards])
        assert self.L.qsb_hostcheck_run_step_fused(self.h, i, self.shard.ctypes.data, ptrs) == 0

    def props(self, i):
        """pass i -> dict(rounds, out_of_place, moved, sync_scatter) or None for an exchange marker"""
        o = (C.c_int * 4)()
        if self.L.qsb_hostcheck_step_props(self.h, i, o):
            return None
        return dict(rounds=o[0], out_of_place=o[1], moved=o[2], sync_scatter=o[3])

    def chunks(self):
        return self.shard.reshape(self.world, -1)

    def finish(self):
        rep = (C.c_int * 5)()
        perm = np.zeros(64, dtype=np.int8)
        self.L.qsb_hostcheck_finish(self.h, rep, perm.ctypes.data)
        self.h = None
        return perm, dict(max_conflict=rep[0], bad_slots=rep[1], noncontig=rep[2], passes=rep[3], swaps=rep[4])


def gather_logical(shards, perm, n, nloc):
    """shards[rank][local physical index] -> complex128 state in logical order."""
    full = np.concatenate(shards)
    idx = np.arange(1 << n, dtype=np.uint64)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(int(perm[q]))
    return full[phys]


def sharded_host_run(circ_gates, n, world, precision=32, low_bits=0, swap_min_ops=0, fused=False):
    """All ranks in one process; the exchange is the chunk transpose the NCCL path performs, or (fused) the
    peer scatter of the pass itself."""
    ranks = [ShardedHostRun(circ_gates, n, world, r, precision, low_bits, swap_min_ops, fused) for r in range(world)]
    for i in range(ranks[0].steps):
        if ranks[0].kind(i) == 2:
            new = [np.zeros_like(r.shard) for r in ranks]
            for r in ranks:
                r.run_fused(i, new)
            for r in ranks:
                r.shard = new[r.rank]
        elif ranks[0].is_swap(i):
            ch = [r.chunks().copy() for r in ranks]
            for r in ranks:
                for j in range(world):
                    r.chunks()[j] = ch[j][r.rank]      # chunk r.rank of rank j lands at position j
        else:
            for r in ranks:
                r.run(i)
    nloc = ranks[0].nloc
    shards = [r.shard for r in ranks]
    out = [r.finish() for r in ranks]
    perm, rep = out[0]
    for p2, _ in out[1:]:
        assert np.array_equal(perm, p2)
    return gather_logical(shards, perm, n, nloc), rep
