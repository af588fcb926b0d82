"""``pnqp`` -- drop-in for the reference's projected-Newton box QP (pnqp.py:5-82):
``pnqp(H, q, lower, upper, x_init=None, n_iter=20) -> (x, H_or_LU, If, i)`` with the
reference's batch-global termination / Armijo semantics."""
import ctypes as C

import torch

from . import _lib
from ._solver import _DT, _ptr, _stream


def pnqp(H, q, lower, upper, x_init=None, n_iter=20, solo=False):
    if n_iter != _lib.PNQP_MAX_ITER:
        raise NotImplementedError("n_iter is fixed at 20 (lqr_step.py:137)")
    if not H.is_cuda:
        raise _lib.DilqrLibraryError("pnqp runs on CUDA tensors only")
    B, n, _ = H.shape
    dtype, dev = H.dtype, H.device
    Hc = H.detach().contiguous()
    qc = q.detach().contiguous()

    def bound(v):
        if isinstance(v, float):
            return torch.full((B, n), v, dtype=dtype, device=dev)
        return v.detach().to(dtype).expand(B, n).contiguous()

    lo, hi = bound(lower), bound(upper)
    x0 = None if x_init is None else x_init.detach().contiguous()
    x = torch.empty(B, n, dtype=dtype, device=dev)
    lu = torch.empty(B, n, n, dtype=dtype, device=dev)
    piv = torch.empty(B, n, dtype=torch.int32, device=dev)
    If = torch.empty(B, n, dtype=dtype, device=dev)
    trace = torch.zeros(2 * _lib.PNQP_MAX_ITER, dtype=torch.int32, device=dev)
    status_dev = torch.zeros(64, dtype=torch.uint8, device=dev)
    for _ in range(64):
        _lib.call("dilqr_pnqp", _DT[dtype], n, B, _ptr(Hc), _ptr(qc), _ptr(lo), _ptr(hi), _ptr(x0),
                  _ptr(x), _ptr(lu), _ptr(piv), _ptr(If), _ptr(trace), 1 if solo else 0,
                  _ptr(status_dev), _stream())
        st = _lib.DilqrStatus.from_buffer_copy(status_dev.cpu().numpy().tobytes())
        if st.trace_match or solo:
            break
    else:
        raise _lib.DilqrLibraryError("pnqp control-flow trace did not stabilise")
    i = int(st.n_total_qp_iter) - 1
    if st.pnqp_unconverged:
        print("[WARNING] pnqp warning: Did not converge")          # pnqp.py:81
    fac = lu if n == 1 else (lu, piv)
    return x, fac, If, i
