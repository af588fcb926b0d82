"""CartpoleDx -- drop-in for env_dx/cartpole.py:29-97,790-867 of the reference.
State (x, dx, cos th, sin th, dth), control force, theta = (g, m_c, m_p, l)."""
import numpy as np
import torch

from .. import _lib
from ._base import EnvDx


class CartpoleDx(EnvDx):
    _dilqr_kind = _lib.DYN_CARTPOLE

    def __init__(self, params=None):
        super().__init__()
        self.n_state, self.n_ctrl = 5, 1
        if params is None:
            self.params = torch.tensor((9.8, 1.0, 0.1, 0.5))      # cartpole.py:39
        else:
            self.params = params
        assert len(self.params) == 4
        self.force_mag = 100.
        self.theta_threshold_radians = np.pi
        self.x_threshold = 2.4
        self.max_velocity = 10
        self.dt = 0.05
        self.lower, self.upper = -self.force_mag, self.force_mag
        self.goal_state = torch.tensor([0., 0., 1., 0., 0.])
        self.goal_weights = torch.tensor([0.1, 0.1, 1., 1., 0.1])
        self.ctrl_penalty = 0.001
        self.mpc_eps = 1e-4                                         # cartpole.py:60-62
        self.linesearch_decay = 0.5
        self.max_linesearch_iter = 2
