"""Caller-facing batched helpers with the reference's names (util.py:32-153).
None of them is on the hot path any more (the fused kernels do this arithmetic in
registers); they are thin device-tensor utilities kept for drop-in compatibility."""
import torch

from . import _lib
from .definitions import QuadCost, LinDx


def bdiag(d):                                   # util.py:32-39
    return torch.diag_embed(d)


def bger(x, y):                                 # util.py:42
    return x.unsqueeze(2) * y.unsqueeze(1)


def bmv(X, y):                                  # util.py:46
    return torch.einsum("bij,bj->bi", X, y)


def bquad(x, Q):                                # util.py:50
    return torch.einsum("bi,bij,bj->b", x, Q, x)


def bdot(x, y):                                 # util.py:54
    return (x * y).sum(1)


def eclamp(x, lower, upper):                    # util.py:58-72  (in place, like the reference)
    lo = lower if not isinstance(lower, float) else torch.full_like(x, lower)
    hi = upper if not isinstance(upper, float) else torch.full_like(x, upper)
    I = x < lo
    x[I] = lo[I]
    I = x > hi
    x[I] = hi[I]
    return x


def get_data_maybe(x):                          # util.py:75
    return x.detach() if torch.is_tensor(x) else x


def detach_maybe(x):                            # util.py:156
    if x is None:
        return None
    return x.detach() if x.requires_grad else x


def get_traj(T, u, x_init, dynamics):           # util.py:104-127
    """Nominal rollout x_{t+1} = f(x_t, u_t): a kernel for the env_dx models, a short
    batched loop for LinDx."""
    if isinstance(dynamics, LinDx):
        x = [x_init.detach()]
        for t in range(T - 1):
            nx = bmv(dynamics.F[t].detach(), torch.cat((x[t], u[t].detach()), 1))
            if dynamics.f is not None and dynamics.f.nelement() > 0:
                nx = nx + dynamics.f[t].detach()
            x.append(nx)
        return torch.stack(x, 0)
    from ._solver import _DT, _ptr, _stream
    B = x_init.shape[0]
    out = torch.empty(T, B, dynamics.n_state, dtype=x_init.dtype, device=x_init.device)
    _lib.call("dilqr_rollout", _DT[x_init.dtype], dynamics._dilqr_kind, dynamics._theta(), T, B,
              _ptr(x_init.detach().contiguous()), _ptr(u.detach().contiguous()), _ptr(out),
              _stream())
    return out


def get_cost(T, u, cost, dynamics=None, x_init=None, x=None):   # util.py:130-153
    assert x_init is not None or x is not None
    assert isinstance(cost, QuadCost)
    if x is None:
        x = get_traj(T, u, x_init, dynamics)
    tau = torch.cat((x, u), 2)
    C, c = cost.C.detach(), cost.c.detach()
    return (0.5 * torch.einsum("tbi,tbij,tbj->tb", tau, C, tau)
            + (tau * c).sum(2)).sum(0)
