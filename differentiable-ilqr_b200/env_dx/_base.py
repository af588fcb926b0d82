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

    def _theta_list(self):
        """Host copy of the parameters (the kernels take them by value).  Cached per
        tensor version so one device->host sync serves every call of a step."""
        p = self.params
        key = (p.data_ptr(), p._version)
        if getattr(self, "_theta_key", None) != key:
            self._theta_cache = p.detach().double().cpu().tolist()
            self._theta_key = key
        return self._theta_cache

    def set_host_params(self, values):
        """Tell the model the current parameter values from a host copy the caller keeps
        (e.g. the training loop's master copy): no device->host read is ever needed."""
        p = self.params
        self._theta_cache = [float(v) for v in values]
        self._theta_key = (p.data_ptr(), p._version)

    def invalidate_params(self):
        """Forget the cached host copy (call after updating ``params`` through ``.data``,
        which does not bump the tensor version the cache is keyed on)."""
        self._theta_key = None

    def _theta(self):
        th = (C.c_double * 8)()
        for i, v in enumerate(self._theta_list()):
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

    _n_theta = None

    def get_matrices(self, x, u):
        """The reference's get_matrices (cartpole.py:105-716 / pendulum.py:152-382 /
        rocket.py:258-261): D, D_grad_params, D_grad_x, D_grad_u, x_grad_theta,
        x_grad_xtm1, x_grad_utm1 for rows x[N,ns], u[N,nc]."""
        N = x.shape[0]
        ns, nc = self.n_state, self.n_ctrl
        n, nth = ns + nc, len(self.params)
        dt, dev = x.dtype, x.device
        shapes = [(N, ns, n), (N, ns, n, nth), (N, ns, n, ns), (N, ns, n, nc), (N, ns, nth),
                  (N, ns, ns), (N, ns, nc)]
        outs = [torch.empty(s, dtype=dt, device=dev) for s in shapes]
        arr = (C.c_void_p * 7)(*[o.data_ptr() for o in outs])
        _lib.call("dilqr_env_tables", _DT[dt], self._dilqr_kind, self._theta(), N,
                  _ptr(x.detach().contiguous()), _ptr(u.detach().to(dt).contiguous()), arr,
                  _stream())
        return tuple(outs)

    def grad_input(self, X, U, K=None):
        """Closed-loop parameter-sensitivity rollout (cartpole.py:717-788, pendulum.py:
        383-443, rocket.py:263-323) returning the reference's seven tensors.  The tables
        come from the device kernel; the short recursion over T is batched device glue
        (the backward pass proper never materialises these tensors, see
        csrc/dilqr_backward.cuh)."""
        T, B, ns = X.shape
        nc = U.shape[2]
        n = ns + nc
        D, Dth, Dx, Du, xth, xx, xu = self.get_matrices(X.reshape(T * B, ns), U.reshape(T * B, nc))
        nth = Dth.shape[-1]
        D, Dth = D.reshape(T, B, ns, n), Dth.reshape(T, B, ns, n, nth)
        Dx, Du = Dx.reshape(T, B, ns, n, ns), Du.reshape(T, B, ns, n, nc)
        xth, xx, xu = xth.reshape(T, B, ns, nth), xx.reshape(T, B, ns, ns), xu.reshape(T, B, ns, nc)
        XU = torch.cat((X, U), -1)
        d_x = torch.einsum("tbnmk,tbm->tbnk", -Dx, XU)
        d_u = torch.einsum("tbnmk,tbm->tbnk", -Du, XU)
        G = torch.zeros(B, ns, nth, dtype=X.dtype, device=X.device)
        zK = torch.zeros(B, nc, ns, dtype=X.dtype, device=X.device)
        grad_D, grad_d = [], []
        Gm1 = None
        Ktm1 = zK
        for t in range(T):
            Kt = zK if K is None else K[t]
            if t > 0:
                Ktm1 = zK if K is None else K[t - 1]
                Gm1 = G
                G = xth[t] + torch.matmul(xx[t] + torch.matmul(xu[t], Ktm1), G)
            if t < T - 1:
                grad_D.append(Dth[t] + torch.matmul(Dx[t] + torch.matmul(Du[t], Kt.unsqueeze(1)),
                                                    G.unsqueeze(1)))
            if t > 0:
                Z = torch.cat((Gm1, torch.matmul(Ktm1, Gm1)), 1)
                grad_d.append(G - torch.einsum("bnmk,bm->bnk", grad_D[t - 1], XU[t - 1])
                              - torch.matmul(D[t - 1], Z))
        return (torch.stack(grad_D), torch.stack(grad_d), Dx[:T - 1], Du[:T - 1], D[:T - 1],
                d_x[:T - 1], d_u[:T - 1])

    def get_true_obj(self):
        q = torch.cat((self.goal_weights, self.ctrl_penalty * torch.ones(self.n_ctrl)))
        px = -torch.sqrt(self.goal_weights) * self.goal_state
        p = torch.cat((px, torch.zeros(self.n_ctrl)))
        return q, p
