"""The C-ABI library loads on a CPU-only box and exports every symbol include/qsim_b200.h declares."""
import ctypes
import os
import re

import pytest

import helpers
import gpu_quantum_simulator_b200 as q
from gpu_quantum_simulator_b200 import _lib

HEADER = os.path.join(helpers.ROOT, "include", "qsim_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes binding"


def test_struct_sizes_match_header_layout():
    assert ctypes.sizeof(q.Gate) == 8 + 4 + 4 + 64
    assert ctypes.sizeof(q.Options) == 16 * 4
    assert ctypes.sizeof(q.RunStats) == 8 * 4 + 4 * 4 + 8 * 3


def test_refcompat_library_exports_reference_names():
    path = os.path.join(os.path.dirname(_lib.LIB_PATH), "libqsim_b200_refcompat.so")
    L = ctypes.CDLL(path)
    for n in ["compute_state_vector", "execute_single_qubit_gate", "execute_cnot",
              "compute_state_cumulative_distribution", "measurement"]:
        assert hasattr(L, n)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(q.QsbError) as e:
        q.Simulator(4)
    assert e.value.code == -5 and "no CPU path" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(helpers.ROOT, "gpu_quantum_simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "qsim_oracle" not in text and "hostcheck" not in text, f
