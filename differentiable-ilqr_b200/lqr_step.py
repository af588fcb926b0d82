"""``LQRStep`` -- drop-in for the reference's single box-constrained LQR step
(lqr_step.py:22-409): ``LQRStep(n_state, n_ctrl, T, ...)(x_init, C, c, F, f)`` returns
``(new_x, new_u, n_total_qp_iter, costs, full_du_norm, mean_alphas)`` (or
``(current_x, current_u)`` with ``no_op_forward=True``) and differentiates through the
KKT adjoint (lqr_step.py:312-407).

The linearisation ``F`` handed to the call is the one the Riccati sweep uses for
LinDx problems; for env_dx dynamics the fused kernel re-derives the same analytic
Jacobian from (current_x, current_u), which is what ``MPC.forward`` passes anyway."""
import torch
from torch.autograd import Function

from . import _lib, _solver
from .definitions import QuadCost, LinDx
from .mpc import _dyn_spec


def LQRStep(n_state, n_ctrl, T, u_lower=None, u_upper=None, u_zero_I=None, delta_u=None,
            linesearch_decay=0.2, max_linesearch_iter=10, true_cost=None, true_dynamics=None,
            delta_space=True, current_x=None, current_u=None, verbose=0, back_eps=1e-3,
            no_op_forward=False, gain_solve=_lib.GAIN_PLAIN, solo=False):
    assert delta_space                                       # lqr_step.py:297-298
    if true_cost is not None and not isinstance(true_cost, QuadCost):
        raise NotImplementedError("only QuadCost is supported (SURVEY 8a-2)")

    class LQRStepFn(Function):
        @staticmethod
        def forward(ctx, x_init, C, c, F, f=None):
            if f is not None and f.nelement() == 0:
                f = None
            ctx.has_f = f is not None
            if no_op_forward:                                # lqr_step.py:278-282
                ctx.save_for_backward(x_init, C, c, F, f if f is not None else x_init.new_empty(0),
                                      current_x, current_u)
                return current_x, current_u
            assert current_x is not None and current_u is not None
            if isinstance(true_dynamics, LinDx) or true_dynamics is None:
                dyn = _solver.DynSpec(_lib.DYN_LINDX, F=F, f=f)
            else:
                dyn = _dyn_spec(true_dynamics)
            x, u, costs, info = _solver.solve_mpc(
                x_init, C, c, dyn, n_state, n_ctrl, T, u_lower=u_lower, u_upper=u_upper,
                u_zero_I=u_zero_I, u_init=current_u, x_cur=current_x,
                linesearch_decay=linesearch_decay, max_linesearch_iter=max_linesearch_iter,
                gain_solve=gain_solve, solo=solo, verbose=verbose, delta_u=delta_u)
            ctx.save_for_backward(x_init, C, c, F, f if f is not None else x_init.new_empty(0),
                                  x, u)
            n_qp = torch.Tensor([info.qp_iters[0]])          # float32, lqr_step.py:308
            outs = (x, u, n_qp, costs, info.full_du_norm, info.alphas.mean())
            ctx.mark_non_differentiable(*outs[2:])
            return outs

        @staticmethod
        def backward(ctx, dl_dx, dl_du, *unused):
            x_init, C, c, F, f, x, u = ctx.saved_tensors
            dx0, dC, dc, dF, df = _solver.kkt_backward(
                dl_dx.contiguous(), dl_du.contiguous(), x_init, C, c, F, f, x, u, n_state, n_ctrl,
                u_lower, u_upper, gain_solve=gain_solve, back_eps=back_eps)
            if not ctx.has_f:
                df = None
            return dx0, dC, dc, dF, df

    return LQRStepFn.apply
