"""PendulumDx -- drop-in for env_dx/pendulum.py:28-125,444-475 of the reference
(simple=True model).  State (cos th, sin th, dth), torque, theta = (g, m, l)."""
import torch

from .. import _lib
from ._base import EnvDx


class PendulumDx(EnvDx):
    _dilqr_kind = _lib.DYN_PENDULUM

    def __init__(self, params=None, simple=True):
        super().__init__()
        if not simple:
            raise NotImplementedError("pendulum-complex is out of scope (SURVEY App. A)")
        self.simple = simple
        self.max_torque = 2.0
        self.dt = 0.05
        self.n_state, self.n_ctrl = 3, 1
        if params is None:
            self.params = torch.tensor((10., 1., 1.), requires_grad=True)   # pendulum.py:41
        else:
            self.params = params
        assert len(self.params) == 3
        self.goal_state = torch.tensor([1., 0., 0.])
        self.goal_weights = torch.tensor([1., 1., 0.1])
        self.ctrl_penalty = 0.001
        self.lower, self.upper = -2., 2.
        self.mpc_eps = 1e-3                                         # pendulum.py:56-58
        self.linesearch_decay = 0.2
        self.max_linesearch_iter = 5
