"""Shared host-side plumbing of the env_dx dynamics models: every numeric method
is a kernel call through the C ABI (no torch arithmetic on the hot path)."""
import ctypes as C

import torch
from torch import nn

from .. import _lib
from .._solver import _DT, _ptr, _stream


class EnvDx(nn.Module):
    _dilqr_kind = None      # DILQR_DYN_*
    n_state = 0
    n_ctrl = 0

    def _theta(self):
        th = (C.c_double * 8)()
        for i, v in enumerate(self.params.detach().double().cpu().tolist()):
            th[i] = v
        return th

    def forward(self, x, u):
        """One Euler step x' = f(x, u); accepts [N,ns],[N,nc] or 1-D inputs."""
        squeeze = x.ndimension() == 1
        if squeeze:
            x, u = x.unsqueeze(0), u.unsqueeze(0)
        if not x.is_cuda:
            raise _lib.DilqrLibraryError("env_dx models run on CUDA tensors only")
        N = x.shape[0]
        x0 = x.detach().contiguous()
        # T=2 rollout: x[1] = f(x0, u[0])
        uu = torch.cat((u.detach().to(x.dtype).reshape(1, N, self.n_ctrl),
                        torch.zeros(1, N, self.n_ctrl, dtype=x.dtype, device=x.device)), 0)
        out = torch.empty(2, N, self.n_state, dtype=x.dtype, device=x.device)
        _lib.call("dilqr_rollout", _DT[x.dtype], self._dilqr_kind, self._theta(), 2, N,
                                            _ptr(x0), _ptr(uu), _ptr(out), _stream())
        y = out[1]
        return y.squeeze(0) if squeeze else y

    def get_linear_dyn(self, x, u):
        """Analytic Jacobian D[N, ns, ns+nc] of the step wrt (x, u)."""
        N = x.shape[0]
        n = self.n_state + self.n_ctrl
        xx = torch.stack((x.detach(), x.detach()), 0).contiguous()
        uu = torch.stack((u.detach(), u.detach()), 0).to(x.dtype).contiguous()
        D = torch.empty(1, N, self.n_state, n, dtype=x.dtype, device=x.device)
        _lib.call("dilqr_linearize", _DT[x.dtype], self._dilqr_kind, self._theta(), 2, N,
                                              _ptr(xx), _ptr(uu), _ptr(D), None, _stream())
        return D[0]

    def linearize_traj(self, x, u):
        """F[T-1,B,ns,n], f[T-1,B,ns] along a trajectory (mpc_explicit.py:516-546)."""
        T, B = x.shape[0], x.shape[1]
        n = self.n_state + self.n_ctrl
        F = torch.empty(T - 1, B, self.n_state, n, dtype=x.dtype, device=x.device)
        f = torch.empty(T - 1, B, self.n_state, dtype=x.dtype, device=x.device)
        _lib.call("dilqr_linearize", _DT[x.dtype], self._dilqr_kind, self._theta(), T, B,
                                              _ptr(x.detach().contiguous()),
                                              _ptr(u.detach().contiguous()), _ptr(F), _ptr(f),
                                              _stream())
        return F, f

    def get_true_obj(self):
        q = torch.cat((self.goal_weights, self.ctrl_penalty * torch.ones(self.n_ctrl)))
        px = -torch.sqrt(self.goal_weights) * self.goal_state
        p = torch.cat((px, torch.zeros(self.n_ctrl)))
        return q, p
