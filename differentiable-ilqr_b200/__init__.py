"""differentiable-ilqr_b200 -- B200-native (sm_100a) batched differentiable
iLQR / MPC solver; drop-in for the hot path of josef-w/Differentiable-iLQR.

The directory name is not a valid Python identifier; import it with
``importlib.import_module("differentiable-ilqr_b200")`` or through the
``dilqr_b200`` alias module at the repo root.
"""
from . import _lib
from .definitions import QuadCost, LinDx
from .mpc import MPC, GradMethods
from . import (mpc, mpc_explicit, env_dx, il, il_env, parallel, lqr_step, lqr_step_explicit,  # noqa: F401
               util, dynamics, dropin)
from .dynamics import AffineDynamics, CtrlPassthroughDynamics, NNDynamics
from . import pnqp as _pnqp_mod  # noqa: F401
from .pnqp import pnqp
from .lqr_step import LQRStep

__all__ = ["MPC", "GradMethods", "QuadCost", "LinDx", "LQRStep", "pnqp", "build",
           "AffineDynamics", "CtrlPassthroughDynamics", "NNDynamics"]


def build(verbose=False):
    return _lib.build(verbose=verbose)
