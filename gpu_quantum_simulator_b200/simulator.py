"""Host-side mirror of the reference interface for the state-vector path.

`Simulator` wraps one qsb_t handle.  The method names follow the functions of
/root/reference/quantum_simulator.c they stand in for:
    compute_state_vector(file)                 :115-254
    execute_single_qubit_gate(U, target)       :81-92   (same U[4] indexing as the reference)
    execute_cnot(control, target)              :94-106
    compute_state_cumulative_distribution()    :256-268
    measurement(shots, seed)                   :270-283
"""
import ctypes as C
import os

import numpy as np

from ._lib import lib, check, Gate, Options, RunStats, F32, F64, MODE_TILED, MODE_SWEEP


def _gate_array(gates):
    if isinstance(gates, C.Array):
        return gates, len(gates)
    arr = (Gate * max(len(gates), 1))()
    for i, g in enumerate(gates):
        arr[i] = g
    return arr, len(gates)


def parse_qasm_file(path):
    """-> (num_qubits, ctypes array of Gate)"""
    nq = C.c_int()
    gp = C.POINTER(Gate)()
    n = C.c_size_t()
    check(lib.qsb_parse_qasm_file(str(path).encode(), C.byref(nq), C.byref(gp), C.byref(n)))
    arr = (Gate * max(n.value, 1))()
    C.memmove(arr, gp, C.sizeof(Gate) * n.value)
    lib.qsb_free(gp)
    return nq.value, (Gate * n.value).from_buffer(arr) if n.value else (Gate * 0)()


def parse_qasm_string(text):
    nq = C.c_int()
    gp = C.POINTER(Gate)()
    n = C.c_size_t()
    check(lib.qsb_parse_qasm_string(text.encode(), C.byref(nq), C.byref(gp), C.byref(n)))
    arr = (Gate * max(n.value, 1))()
    C.memmove(arr, gp, C.sizeof(Gate) * n.value)
    lib.qsb_free(gp)
    return nq.value, (Gate * n.value).from_buffer(arr) if n.value else (Gate * 0)()


def gates_from_circuit(circ):
    """circ: iterable of (name, qubits, params) -> ctypes array of Gate (via qsb_gate_from_name)."""
    out = []
    tmp = (Gate * 3)()
    k = C.c_int()
    for name, qubits, params in circ:
        p = (C.c_double * max(len(params), 1))(*params)
        q = (C.c_int * max(len(qubits), 1))(*qubits)
        check(lib.qsb_gate_from_name(name.encode(), p, len(params), q, len(qubits), tmp, C.byref(k)))
        for i in range(k.value):
            g = Gate()
            C.memmove(C.byref(g), C.byref(tmp[i]), C.sizeof(Gate))
            out.append(g)
    arr = (Gate * max(len(out), 1))()
    for i, g in enumerate(out):
        arr[i] = g
    return (Gate * len(out)).from_buffer(arr) if out else (Gate * 0)()


def _options(precision, mode, low_bits, rank, world_size, device, reserved=None, use_graph=False, dense_k=0):
    """reserved: planner tuning knobs (qsb_options_t.reserved): [0] min gates before a qubit exchange,
    [1] 2 = lazy diagonals on, [2] k+1 = trim tail rounds with < k gates (1 = off), [3] fusion-depth cost cap,
    [4] gate rewrites (5 = CX between two h / two rx left alone, 6 = CX -> controlled phase with an h on one side too),
    [5] exchange flavour, [6] 2 = first-come tiles, 3 / 4 = no / conflicts-only lane relocation; the full list is in
    include/qsim_b200.h.  Knobs left at 0 are searched by the library where it plans several candidates."""
    o = Options()
    lib.qsb_options_default(C.byref(o))
    for k, v in enumerate(reserved or ()):
        o.reserved[k] = int(v)
    o.precision = precision
    o.mode = mode
    o.low_bits = low_bits
    o.rank = rank
    o.world_size = world_size
    o.device = device
    o.use_graph = 1 if use_graph else 0
    o.tile_bits = dense_k          # MODE_DENSE: fusion width k (2..5)
    return o


def plan_dry_run(num_qubits, gates, precision=F32, low_bits=0, world_size=1, rank=0, reserved=None, mode=MODE_TILED, dense_k=0):
    """Host-only scheduling statistics (no GPU needed)."""
    arr, n = _gate_array(gates)
    o = _options(precision, mode, low_bits, rank, world_size, -1, reserved, False, dense_k)
    st = RunStats()
    check(lib.qsb_plan_dry_run(num_qubits, C.byref(o), arr, n, C.byref(st)))
    return st.as_dict()


class Plan:
    def __init__(self, sim, handle):
        self.sim, self.handle = sim, handle

    def stats(self):
        st = RunStats()
        check(lib.qsb_plan_stats(self.handle, C.byref(st)))
        return st.as_dict()

    def close(self):
        if self.handle:
            lib.qsb_plan_destroy(self.handle)
            self.handle = None

    __del__ = close


class Simulator:
    def __init__(self, num_qubits, precision=F32, mode=MODE_TILED, low_bits=0, rank=0, world_size=1, device=-1, reserved=None,
                 use_graph=False, dense_k=0):
        self._h = C.c_void_p()
        o = _options(precision, mode, low_bits, rank, world_size, device, reserved, use_graph, dense_k)
        check(lib.qsb_create(C.byref(self._h), num_qubits, C.byref(o)))
        self.num_qubits, self.precision = num_qubits, precision
        self.rank, self.world_size = rank, world_size

    # ---- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            lib.qsb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        check(lib.qsb_reset(self._h))

    # ---- hot path
    def apply(self, gates):
        arr, n = _gate_array(gates)
        check(lib.qsb_apply_gates(self._h, arr, n))
        return self.last_stats()

    def plan(self, gates):
        arr, n = _gate_array(gates)
        p = C.c_void_p()
        check(lib.qsb_plan_create(self._h, arr, n, C.byref(p)))
        return Plan(self, p)

    def execute(self, plan):
        check(lib.qsb_execute(self._h, plan.handle))
        return self.last_stats()

    def last_stats(self):
        st = RunStats()
        check(lib.qsb_last_run_stats(self._h, C.byref(st)))
        return st.as_dict()

    # ---- reference-named operations
    def execute_single_qubit_gate(self, U, target):
        """U: 4 complex numbers indexed as the reference does: v0' = v0*U[0] + v1*U[2] (quantum_simulator.c:88)."""
        U = np.asarray(U, dtype=np.complex128).reshape(4)
        g = Gate()
        g.target = target
        m = [U[0], U[2], U[1], U[3]]
        for k in range(4):
            g.m[2 * k], g.m[2 * k + 1] = m[k].real, m[k].imag
        return self.apply([g])

    def execute_cnot(self, control, target):
        g = Gate()
        g.controls, g.target = 1 << control, target
        g.m[2] = 1.0
        g.m[4] = 1.0
        return self.apply([g])

    def compute_state_vector(self, path):
        nq, gates = parse_qasm_file(path)
        if nq != self.num_qubits:
            raise ValueError(f"circuit declares {nq} qubits, simulator has {self.num_qubits}")
        self.reset()
        self.apply(gates)
        return self.state()

    # ---- readout
    def _own_range(self):
        n = 1 << self.num_qubits
        return 0, n

    def state(self, first=0, count=None):
        """complex128 amplitudes in logical order."""
        if count is None:
            count = (1 << self.num_qubits) - first
        out = np.empty(2 * count, dtype=np.float64)
        check(lib.qsb_download(self._h, out.ctypes.data, first, count))
        return out.view(np.complex128)

    def state_native(self, first=0, count=None, out=None):
        if count is None:
            count = (1 << self.num_qubits) - first
        dt = np.float32 if self.precision == F32 else np.float64
        if out is None:
            out = np.empty(2 * count, dtype=dt)
        check(lib.qsb_download_native(self._h, out.ctypes.data, first, count))
        return out

    def layout(self):
        """-> (perm, nloc): perm[q] = physical index bit of logical qubit q; bits >= nloc are rank bits."""
        perm = np.zeros(64, dtype=np.int8)
        nloc = C.c_int()
        check(lib.qsb_get_layout(self._h, perm.ctypes.data, C.byref(nloc)))
        return perm, nloc.value

    def shard_physical(self):
        """The local shard exactly as stored (physical index order), complex128."""
        _, nloc = self.layout()
        out = np.empty(2 << nloc, dtype=np.float64)
        check(lib.qsb_download_physical(self._h, out.ctypes.data, 0, 1 << nloc))
        return out.view(np.complex128)

    def shard_head(self, count, out=None):
        """The first `count` amplitudes of the local shard in physical order, as float64 (re, im) pairs."""
        if out is None:
            out = np.empty(2 * count, dtype=np.float64)
        check(lib.qsb_download_physical(self._h, out.ctypes.data, 0, count))
        return out

    def set_state(self, amps, first=0):
        a = np.ascontiguousarray(np.asarray(amps, dtype=np.complex128))
        check(lib.qsb_upload(self._h, a.view(np.float64).ctypes.data, first, a.size))

    def norm_argmax(self):
        norm, idx, p = C.c_double(), C.c_uint64(), C.c_double()
        check(lib.qsb_norm_argmax(self._h, C.byref(norm), C.byref(idx), C.byref(p)))
        return norm.value, idx.value, p.value

    def probabilities(self, first=0, count=None):
        if count is None:
            count = (1 << self.num_qubits) - first
        out = np.empty(count, dtype=np.float64)
        check(lib.qsb_probabilities(self._h, out.ctypes.data, first, count))
        return out

    def compute_state_cumulative_distribution(self, first=0, count=None):
        if count is None:
            count = (1 << self.num_qubits) - first
        out = np.empty(count, dtype=np.float64)
        check(lib.qsb_cdf(self._h, out.ctypes.data, first, count))
        return out

    def measurement(self, shots, seed=0):
        out = np.empty(shots, dtype=np.uint64)
        check(lib.qsb_sample(self._h, seed, shots, out.ctypes.data))
        return out

    def save_state(self, path):
        """Dump this rank's shard (device dtype, physical order) with its qubit map; one file per rank."""
        check(lib.qsb_save_state(self._h, os.fsencode(path)))

    def load_state(self, path):
        check(lib.qsb_load_state(self._h, os.fsencode(path)))


def sample_uniform(seed, k):
    """The r in [0, 1) that shot k of Simulator.measurement(seed=seed) searches for."""
    return lib.qsb_sample_uniform(seed, k)
