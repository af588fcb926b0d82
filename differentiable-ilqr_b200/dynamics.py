"""Dynamics containers of the upstream mpc.pytorch API kept by the reference
(dynamics.py:133-202): ``AffineDynamics`` and ``CtrlPassthroughDynamics``.

``AffineDynamics`` is a time-invariant LinDx: ``MPC.forward`` hands the solver
``F = [A B]`` and ``f = c`` broadcast over (T-1, B) and autograd sums the KKT
gradients ``dF, df`` back to ``A, B, c`` -- which is what the reference computes
through ``linearize_dynamics(diff=True)`` (mpc.py:504-523: ``f = x' - R x - S u``
is identically ``c``).  ``forward`` / ``grad_input`` are plain tensor expressions
for callers that simulate the plant; the solve itself never calls them.
"""
import torch
from torch import nn


class AffineDynamics(nn.Module):
    def __init__(self, A, B, c=None):
        super().__init__()
        assert A.ndimension() == 2
        assert B.ndimension() == 2
        if c is not None:
            assert c.ndimension() == 1
        self.A, self.B, self.c = A, B, c

    def forward(self, x, u):                               # dynamics.py:173-195
        x_dim = x.ndimension()
        if x_dim == 1:
            x = x.unsqueeze(0)
        if u.ndimension() == 1:
            u = u.unsqueeze(0)
        z = x.mm(self.A.t()) + u.mm(self.B.t())
        if self.c is not None:
            z = z + self.c
        return z.squeeze(0) if x_dim == 1 else z

    def grad_input(self, x, u):                            # dynamics.py:197-202
        n_batch = x.size(0)
        return (self.A.unsqueeze(0).repeat(n_batch, 1, 1),
                self.B.unsqueeze(0).repeat(n_batch, 1, 1))

    def as_lindx(self, T, n_batch, dtype, device):
        """F[T-1,B,ns,ns+nc], f[T-1,B,ns] (or None) as differentiable views."""
        ns = self.A.shape[0]
        Fm = torch.cat((self.A, self.B), 1).to(device=device, dtype=dtype)
        F = Fm.unsqueeze(0).unsqueeze(0).expand(T - 1, n_batch, ns, Fm.shape[1])
        f = None
        if self.c is not None:
            f = self.c.to(device=device, dtype=dtype).unsqueeze(0).unsqueeze(0).expand(
                T - 1, n_batch, ns)
        return F, f


class CtrlPassthroughDynamics(nn.Module):
    """x~ = (u_prev, x): the augmented plant of the slew-rate formulation
    (dynamics.py:133-156)."""

    def __init__(self, dynamics):
        super().__init__()
        self.dynamics = dynamics

    def forward(self, tilde_x, u):
        dim = tilde_x.ndimension()
        if dim == 1:
            tilde_x = tilde_x.unsqueeze(0)
        if u.ndimension() == 1:
            u = u.unsqueeze(0)
        n_ctrl = u.size(1)
        xtp1 = self.dynamics(tilde_x[:, n_ctrl:], u)
        out = torch.cat((u, xtp1), dim=1)
        return out.squeeze() if dim == 1 else out

    def grad_input(self, x, u):
        assert False, "Unimplemented"                      # dynamics.py:155-156
