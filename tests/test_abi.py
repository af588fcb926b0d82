"""CPU-side checks of the drop-in boundary: libdilqr.so loads and exports every
symbol include/dilqr.h declares; argument validation returns error codes without
touching the GPU."""
import ctypes as C
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "dilqr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dilqr_\w+)\s*\(", src)))


def test_header_symbols_exported():
    lib = importlib.import_module("differentiable-ilqr_b200._lib")
    L = lib.lib()
    names = _header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), "libdilqr.so does not export %s" % n
    # and the Python binding table covers exactly the header
    assert sorted(lib.SYMBOLS) == names


def test_version_and_supported():
    lib = importlib.import_module("differentiable-ilqr_b200._lib")
    L = lib.lib()
    assert b"sm_100a" in L.dilqr_version()
    for dt in (lib.F32, lib.F64):
        assert L.dilqr_supported(dt, 5, 1, lib.DYN_CARTPOLE) == 1
        assert L.dilqr_supported(dt, 3, 1, lib.DYN_PENDULUM) == 1
        assert L.dilqr_supported(dt, 4, 2, lib.DYN_LINDX) == 1
        assert L.dilqr_supported(dt, 16, 4, lib.DYN_LINDX) == 1
        assert L.dilqr_supported(dt, 7, 7, lib.DYN_LINDX) == 0
    assert L.dilqr_supported(5, 5, 1, lib.DYN_CARTPOLE) == 0


def test_argument_validation_returns_codes():
    lib = importlib.import_module("differentiable-ilqr_b200._lib")
    L = lib.lib()
    s = lib.DilqrSolve()
    s.n_state, s.n_ctrl, s.T, s.n_batch, s.dtype = 5, 1, 10, 4, lib.F64
    s.dynamics = lib.DYN_CARTPOLE
    assert L.dilqr_mpc_begin(C.byref(s), None) == -1          # null x_init/C/c
    buf = (C.c_double * 4096)()
    base = C.addressof(buf)
    base += (-base) % 16
    s.x_init = s.C = s.c = base
    assert L.dilqr_mpc_begin(C.byref(s), None) == -1          # no workspace / status
    s.status = base
    s.workspace = base
    s.workspace_bytes = 16
    assert L.dilqr_mpc_iterate(C.byref(s), None) == -4        # workspace too small
    s.C = base + 8
    assert L.dilqr_mpc_begin(C.byref(s), None) == -3          # misaligned
    s.C = base
    s.dtype = 7
    assert L.dilqr_mpc_begin(C.byref(s), None) == -1
    s.dtype = lib.F32
    s.n_state = 7
    s.workspace_bytes = 1 << 40
    assert L.dilqr_mpc_begin(C.byref(s), None) == -2          # shape not compiled in
    need = L.dilqr_workspace_bytes(C.byref(s))
    assert need > 0
    k = lib.DilqrKkt()
    assert L.dilqr_kkt_grads(C.byref(k), None) == -1


def test_no_cpu_path():
    """The product refuses CPU tensors instead of silently computing on the host."""
    import pytest
    import torch
    d = importlib.import_module("differentiable-ilqr_b200")
    lib = importlib.import_module("differentiable-ilqr_b200._lib")
    m = d.MPC(4, 2, 5, lqr_iter=1, verbose=-1)
    C_, c_ = torch.eye(6).repeat(5, 3, 1, 1), torch.zeros(5, 3, 6)
    F = torch.zeros(4, 3, 4, 6)
    with pytest.raises(lib.DilqrLibraryError):
        m(torch.zeros(3, 4), d.QuadCost(C_, c_), d.LinDx(F, None))
