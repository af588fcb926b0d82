"""``lqr_step_explicit.LQRStep`` -- drop-in for the reference's DiLQR variant of the single
LQR step (lqr_step_explicit.py:24-44, 599-712): same factory arguments (plus ``theta``), the
returned function is called as ``(x_init, C, c, F, f, theta)``.

* ``no_op_forward=True`` (how ``mpc_explicit.MPC`` ends its forward, mpc_explicit.py:325-340):
  hands back ``(current_x, current_u)``; the backward is the DiLQR implicit gradient
  (lqr_step_explicit.py:652-712) wrt ``C``, ``c`` and the dynamics parameters ``theta`` --
  ``F``, ``f``, ``x_init`` get no gradient, exactly like the reference (it returns ``dC, dc,
  dtheta`` only).
* otherwise one box-constrained LQR step with line search over the env_dx dynamics
  (lqr_step_explicit.py:620-650), returning ``(new_x, new_u, n_total_qp_iter, costs,
  full_du_norm, mean_alphas)``.
"""
import torch
from torch.autograd import Function

from . import _lib, _solver
from .definitions import QuadCost, LinDx  # noqa: F401
from .mpc import _dyn_spec

RICHARDSON_PASSES = 30
RICHARDSON_TOL = 1e-14


def LQRStep(n_state, n_ctrl, T, u_lower=None, u_upper=None, u_zero_I=None, delta_u=None,
            linesearch_decay=0.2, max_linesearch_iter=10, true_cost=None, true_dynamics=None,
            delta_space=True, current_x=None, current_u=None, verbose=0, back_eps=1e-3,
            no_op_forward=False, theta=None, solo=False):
    assert delta_space                                       # lqr_step_explicit.py:630-644
    assert current_x is not None and current_u is not None
    if true_cost is not None and not isinstance(true_cost, QuadCost):
        raise NotImplementedError("lqr_step_explicit.LQRStep: QuadCost only (SURVEY 8a-2)")
    if getattr(true_dynamics, "_dilqr_kind", None) is None or isinstance(true_dynamics, LinDx):
        raise AttributeError("lqr_step_explicit.LQRStep needs an env_dx model as true_dynamics "
                             "(its backward differentiates the dynamics parameters)")

    class LQRStepFn(Function):
        @staticmethod
        def forward(ctx, x_init, C, c, F, f=None, theta_=None, if_converge=False):
            ctx.set_materialize_grads(False)
            if no_op_forward:                                # lqr_step_explicit.py:604-618
                ctx.save_for_backward(x_init, C, c, current_x, current_u)
                return current_x.clone(), current_u.clone()
            x, u, costs, info = _solver.solve_mpc(
                x_init, C, c, _dyn_spec(true_dynamics), n_state, n_ctrl, T, u_lower=u_lower,
                u_upper=u_upper, u_zero_I=u_zero_I, u_init=current_u, x_cur=current_x,
                linesearch_decay=linesearch_decay, max_linesearch_iter=max_linesearch_iter,
                solo=solo, verbose=verbose, delta_u=delta_u)
            ctx.save_for_backward(x_init, C, c, x, u)
            n_qp = torch.Tensor([info.qp_iters[0]])          # float32, lqr_step_explicit.py:648
            outs = (x, u, n_qp, costs, info.full_du_norm, info.alphas.mean())
            ctx.mark_non_differentiable(*outs[2:])
            return outs

        @staticmethod
        def backward(ctx, dl_dx, dl_du, *unused):
            x_init, C, c, x, u = ctx.saved_tensors
            dC, dc, dtheta = _solver.dilqr_backward(
                dl_dx, dl_du, x_init, C, c, x, u, true_dynamics, n_state, n_ctrl, u_lower, u_upper,
                n_passes=RICHARDSON_PASSES, tol=RICHARDSON_TOL, back_eps=back_eps, solo=solo)
            p = true_dynamics.params
            dth = dtheta.sum(0).to(device=p.device, dtype=p.dtype)
            # the reference returns (dC, dc, dtheta) for (C, c, theta) and nothing else
            return None, dC.reshape(C.shape), dc.reshape(c.shape), None, None, dth, None

    def call(x_init, C, c, F=None, f=None, theta_=None, if_converge=False):
        th = theta_ if theta_ is not None else (theta if theta is not None else true_dynamics.params)
        return LQRStepFn.apply(x_init, C, c, F, f, th, if_converge)

    return call
