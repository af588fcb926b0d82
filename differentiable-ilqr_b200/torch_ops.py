"""``torch.ops.dilqr.*`` -- the hot path as dispatcher-visible torch custom ops
(``TORCH_LIBRARY(dilqr, ...)`` in csrc/torch_ops.cpp, a thin layer over the C ABI of
include/dilqr.h) with their autograd formulas:

  torch.ops.dilqr.mpc_solve(x_init, C, c, F, f, u_init, theta, dynamics, T, u_lower, u_upper,
                            lqr_iter, eps, linesearch_decay, max_linesearch_iter,
                            not_improved_lim, best_cost_eps, solo)
        -> (x, u, costs, full_du_norm, n_total_qp_iter[iterations])        mpc.py:184-337
     differentiable wrt C, c, theta (env_dx models: DiLQR implicit gradient,
     lqr_step_explicit.py:652-712) or x_init, C, c, F, f (LinDx: KKT, lqr_step.py:312-407)
  torch.ops.dilqr.dilqr_backward(...), torch.ops.dilqr.lqr_kkt_backward(...)

The op library is built in-tree (``build()``; __graft_entry__.build does it) into
torch_ops/dilqr_torch_ops.so next to libdilqr.so, which it links."""
import os

import torch

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "torch_ops")
SO_PATH = os.path.join(BUILD_DIR, "dilqr_torch_ops.so")
RICHARDSON_PASSES = 30      # adjoint passes of the registered backward of mpc_solve

_loaded = False


def build(verbose=False):
    """Compile csrc/torch_ops.cpp against this torch (g++, no GPU needed) and load it."""
    from torch.utils import cpp_extension
    os.makedirs(BUILD_DIR, exist_ok=True)
    _lib.lib()      # libdilqr.so must exist: the op library links it
    cpp_extension.load(
        name="dilqr_torch_ops", sources=[os.path.join(_lib.CSRC, "torch_ops.cpp")],
        build_directory=BUILD_DIR, extra_cflags=["-O2", "-std=c++17"],
        extra_ldflags=["-L" + _HERE, "-ldilqr", "-Wl,-rpath," + _HERE], with_cuda=True,
        is_python_module=False, verbose=verbose)
    _register()
    return SO_PATH


def load():
    """Load the op library (fails loudly when it has not been built)."""
    global _loaded
    if _loaded:
        return
    if not hasattr(torch.ops.dilqr, "mpc_solve"):
        if not os.path.isfile(SO_PATH):
            raise _lib.DilqrLibraryError(
                "%s not found -- run `python -c 'import __graft_entry__ as g; g.build()'`" % SO_PATH)
        _lib.lib()
        torch.ops.load_library(SO_PATH)
    _register()


def _register():
    global _loaded
    if _loaded:
        return
    _loaded = True

    @torch.library.register_fake("dilqr::mpc_solve")
    def _(x_init, C, c, F, f, u_init, theta, dynamics, T, u_lower, u_upper, lqr_iter, eps,
          linesearch_decay, max_linesearch_iter, not_improved_lim, best_cost_eps, solo):
        B, ns = x_init.shape
        nc = C.shape[-1] - ns
        return (x_init.new_empty(T, B, ns), x_init.new_empty(T, B, nc), x_init.new_empty(B),
                x_init.new_empty(B), torch.empty(lqr_iter, dtype=torch.long))

    def setup(ctx, inputs, output):
        (x_init, C, c, F, f, u_init, theta, dynamics, T, u_lower, u_upper, lqr_iter, eps, decay,
         max_ls, nil, bce, solo) = inputs
        x, u = output[0], output[1]
        ctx.save_for_backward(x_init, C, c, F, f, theta, x, u)
        ctx.cfg = (dynamics, u_lower, u_upper, max_ls, decay)

    def backward(ctx, gx, gu, g_costs, g_du, g_qp):
        x_init, C, c, F, f, theta, x, u = ctx.saved_tensors
        dynamics, lo, hi, max_ls, decay = ctx.cfg
        if gu is None:
            gu = torch.zeros_like(u)
        out = [None] * 18
        if dynamics == _lib.DYN_LINDX:
            dx0, dC, dc, dF, df = torch.ops.dilqr.lqr_kkt_backward(gx, gu.contiguous(), x_init, C, c,
                                                                   F, x, u, lo, hi)
            out[0] = dx0
            out[1], out[2] = dC.sum_to_size(C.shape), dc.sum_to_size(c.shape)
            out[3] = dF
            out[4] = df if f is not None else None
        else:
            dC, dc, dth = torch.ops.dilqr.dilqr_backward(
                gx, gu.contiguous(), x_init, C, c, x, u, theta, dynamics, lo, hi, RICHARDSON_PASSES,
                max_ls, decay)
            out[1], out[2] = dC.reshape(C.shape), dc.reshape(c.shape)
            out[6] = dth.sum(0).to(device=theta.device, dtype=theta.dtype)
        return tuple(out)

    torch.library.register_autograd("dilqr::mpc_solve", backward, setup_context=setup)
