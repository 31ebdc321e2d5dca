"""The C-ABI boundary: every function declared in include/qsim_b200.h is exported
by libqsim_b200.so (and by the host emulator), with the ctypes signature table of
quantum_computations_b200/_capi.py covering exactly that set.  No compute calls
are made against the CUDA library here (there is no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

from quantum_computations_b200 import _capi, build_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "qsim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(qsim_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_and_ctypes_table_agree():
    assert declared_functions() == sorted(_capi.SIGNATURES)


@pytest.mark.parametrize("which", ["cuda", "emu"])
def test_library_exports_every_declared_symbol(which):
    path = build_native.build_cuda() if which == "cuda" else build_native.build_emu()
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{os.path.basename(path)} does not export {name}"
    _capi.declare(lib)
    assert lib.qsim_version() >= 100
    assert lib.qsim_has_cuda() == (1 if which == "cuda" else 0)


def test_host_side_entry_points_of_the_cuda_library():
    """Circuit container, planner and error reporting are host code and can be
    exercised without a GPU."""
    import numpy as np
    lib = ctypes.CDLL(build_native.build_cuda())
    _capi.declare(lib)
    circ = ctypes.c_void_p()
    assert lib.qsim_circuit_create(5, ctypes.byref(circ)) == 0
    h = (np.array([[1, 1], [1, -1]]) / np.sqrt(2)).astype(np.complex128)
    t = (ctypes.c_int * 1)(2)
    assert lib.qsim_circuit_add_matrix(circ, 1, t, h.view(np.float64).ctypes.data_as(_capi.c_double_p)) == 0
    bad = (ctypes.c_int * 1)(7)
    assert lib.qsim_circuit_add_matrix(circ, 1, bad, h.view(np.float64).ctypes.data_as(_capi.c_double_p)) == -1
    assert b"out of range" in lib.qsim_last_error()
    plan = ctypes.c_void_p()
    assert lib.qsim_plan_compile(circ, None, ctypes.byref(plan)) == 0
    stats = _capi.PlanStats()
    assert lib.qsim_plan_stats(plan, ctypes.byref(stats)) == 0
    assert stats.n_passes == 1 and stats.n_dense == 1
    # executing on a host pointer must be refused, not silently computed on the CPU
    buf = np.zeros(32, dtype=np.complex128)
    rc = lib.qsim_plan_execute(plan, buf.ctypes.data, 5, None, None)
    assert rc != 0
    lib.qsim_plan_destroy(plan)
    lib.qsim_circuit_destroy(circ)
