"""IL_Env -- the caller either side of the hot path (il_env.py:32-188 of the reference):
expert-data generation by open-loop MPC (`populate_data`), by closed-loop
receding-horizon MPC (`populate_data2`) and the `mpc` helper the training loop calls.

Same names and arguments as the reference.  What is B200-first here:

* the diagonal cost is handed to the solver as the `[n,n]` / `[n]` tensors it is
  (broadcast-aware kernels) instead of being `.repeat`-ed to `[T,B,n,n]`
  (il_env.py:159-162); `tile=True` restores the dense layout, results are identical;
* `populate_data2` runs its `n_total` closed loops as ONE batch: the reference calls
  MPC once per (sample, time step) with n_batch=1 (il_env.py:112-143), i.e. every problem
  has its own pnqp / line-search / stop-rule flags.  `solo=2` gives exactly those
  per-problem semantics inside a batched launch, so T solves of B=n_total replace
  n_total*T solves of B=1; the plant step and the warm-start shift stay on the device.
"""
import numpy as np
import torch

from . import mpc_explicit
from .definitions import QuadCost
from .env_dx import cartpole, pendulum
from .il import TileCost
from .mpc_explicit import GradMethods


class IL_Env:
    def __init__(self, env, lqr_iter=100, mpc_T=35, slew_rate_penalty=None, dtype=None,
                 device=None, tile=False):
        self.env = env
        if env == 'pendulum':
            self.true_dx = pendulum.PendulumDx()
        elif env == 'cartpole':
            self.true_dx = cartpole.CartpoleDx()
        elif env == 'pendulum-complex':
            raise NotImplementedError("pendulum-complex (5-parameter damped model, "
                                      "il_env.py:40-42) has no device tables")
        else:
            assert False
        self.lqr_iter = lqr_iter
        self.mpc_T = mpc_T
        self.slew_rate_penalty = slew_rate_penalty     # unused by the reference too (il_env.py:185)
        self.grad_method = GradMethods.ANALYTIC
        self.dtype = dtype or torch.get_default_dtype()
        self.device = torch.device(device if device is not None else "cuda")
        self.tile = tile
        self.train_data = None
        self.val_data = None
        self.test_data = None

    # ------------------------------------------------------------------ sampling
    def sample_xinit(self, n_batch=1):
        """Initial states of il_env.py:57-79, returned on the host.  The reference draws
        from the global CPU generator -- pendulum: angle then angular velocity; cartpole:
        four draws that it multiplies by zero (every cartpole problem starts at rest with
        the pole at pi/1.05) -- so the same number of draws is consumed here, in the same
        order, and a seed reproduces both the states and the generator state afterwards."""
        draw = lambda lo, hi: lo + (hi - lo) * torch.rand(n_batch)
        if self.env in ('pendulum', 'pendulum-complex'):
            angle = draw(-0.5 * np.pi, 0.5 * np.pi)
            rate = draw(-1., 1.)
            return torch.stack((angle.cos(), angle.sin(), rate), dim=1)
        assert self.env == 'cartpole'
        for _ in range(4):                       # x, dx, th, dth: drawn, then discarded
            torch.rand(n_batch)
        rest = torch.zeros(n_batch)
        angle = torch.full((n_batch,), 3.1415926 / 1.05)
        return torch.stack((rest, rest, angle.cos(), angle.sin(), rest), dim=1)

    def _dev(self, t):
        return t.detach().to(device=self.device, dtype=self.dtype)

    def _cost(self, q, p, T, n_batch):
        q, p = self._dev(q), self._dev(p)
        if self.tile:                                    # il_env.py:159-162
            return QuadCost(*TileCost.apply(q, p, T, n_batch))
        return QuadCost(torch.diag(q), p)

    # ------------------------------------------------------------------ open loop
    def populate_data(self, n_train, n_val, n_test, seed=0):
        """Expert data by ONE batched open-loop solve with the true model and the true
        cost (il_env.py:81-94): rows are [T, ns+nc] trajectories, split train / val / test."""
        torch.manual_seed(seed)
        sizes = (n_train, n_val, n_test)
        x0 = self.sample_xinit(n_batch=sum(sizes))
        q_true, p_true = self.true_dx.get_true_obj()
        with torch.no_grad():     # expert data: no gradient is ever asked for
            xs, us = self.mpc(self.true_dx, x0, q_true, p_true)
        rows = torch.cat((xs, us), dim=2).transpose(0, 1)          # [n_data, T, ns+nc]
        self.train_data = rows[:n_train]
        self.val_data = rows[n_train:n_train + n_val]
        self.test_data = rows[-n_test:]                            # il_env.py:94 (tail slice)

    # ------------------------------------------------------------------ closed loop
    def closed_loop(self, x_init, n_steps=None):
        """Receding-horizon rollout of il_env.py:112-143 for a batch of initial states
        [B,ns]: returns tau [B, n_steps, ns+nc] (states before each applied control)."""
        T = self.mpc_T
        n_steps = T if n_steps is None else n_steps
        dx = self.true_dx
        q, p = dx.get_true_obj()
        x = self._dev(x_init)
        B = x.shape[0]
        cost = self._cost(q, p, T, B)
        ctrl = mpc_explicit.MPC(
            dx.n_state, dx.n_ctrl, T, u_lower=dx.lower, u_upper=dx.upper,
            lqr_iter=self.lqr_iter, verbose=-1, exit_unconverged=False,
            detach_unconverged=False, linesearch_decay=dx.linesearch_decay,
            max_linesearch_iter=dx.max_linesearch_iter, grad_method=self.grad_method,
            eps=dx.mpc_eps, n_batch=B, solo=2)
        xs = torch.empty(n_steps, B, dx.n_state, dtype=self.dtype, device=self.device)
        us = torch.empty(n_steps, B, dx.n_ctrl, dtype=self.dtype, device=self.device)
        u_init = None
        with torch.no_grad():
            for t in range(n_steps):
                ctrl.u_init = u_init
                _, nominal_actions, _ = ctrl(x, cost, dx)
                xs[t] = x
                us[t] = nominal_actions[0]
                x = dx(x, nominal_actions[0])                      # plant step (il_env.py:138)
                # warm start: shift by one, pad (il_env.py:141-142)
                u_init = torch.cat((nominal_actions[1:], torch.zeros_like(nominal_actions[:1])), 0)
                u_init[-2] = u_init[-3]
        return torch.cat((xs, us), 2).transpose(0, 1).contiguous()

    def populate_data2(self, n_train, n_val, n_test, seed=0):
        """il_env.py:96-151."""
        torch.manual_seed(seed)
        n_total = n_train + n_val + n_test
        x_init_all = self.sample_xinit(n_batch=n_total)
        all_tau = self.closed_loop(x_init_all)
        self.train_data = all_tau[:n_train]
        self.val_data = all_tau[n_train:n_train + n_val]
        self.test_data = all_tau[n_train + n_val:]

    # ------------------------------------------------------------------ training-loop helper
    def mpc(self, dx, xinit, q, p, u_init=None, eps_override=None, lqr_iter_override=None):
        """il_env.py:153-188 (differentiable in dx.params, q, p)."""
        xinit = xinit.to(device=self.device, dtype=self.dtype)
        n_batch = xinit.shape[0]
        if self.tile:
            q_, p_ = q.to(self.device, self.dtype), p.to(self.device, self.dtype)
            cost = QuadCost(*TileCost.apply(q_, p_, self.mpc_T, n_batch))
        else:
            cost = QuadCost(torch.diag(q.to(self.device, self.dtype)), p.to(self.device, self.dtype))
        eps = eps_override if eps_override else self.true_dx.mpc_eps
        lqr_iter = lqr_iter_override if lqr_iter_override else self.lqr_iter
        if u_init is not None:
            u_init = u_init.to(device=self.device, dtype=self.dtype)
        x_mpc, u_mpc, objs_mpc = mpc_explicit.MPC(
            self.true_dx.n_state, self.true_dx.n_ctrl, self.mpc_T,
            u_lower=self.true_dx.lower, u_upper=self.true_dx.upper, u_init=u_init,
            lqr_iter=lqr_iter, verbose=0, exit_unconverged=False, detach_unconverged=True,
            linesearch_decay=self.true_dx.linesearch_decay,
            max_linesearch_iter=self.true_dx.max_linesearch_iter,
            grad_method=self.grad_method, eps=eps, n_batch=n_batch,
        )(xinit, cost, dx)
        return x_mpc, u_mpc
