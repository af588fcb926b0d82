"""RocketDx -- drop-in for env_dx/rocket.py:14-164,212-232,324-426 of the reference:
6-DoF rocket with quaternion attitude.  State r(3), v(3), q(4), w(3); control thrust
(3); theta = (Jx, Jy, Jz, mass, l).  Mirrors the reference's quirks: the step returns
the UN-normalised quaternion (rocket.py:158-164) and the tilt penalty enters the cost
twice (rocket.py:74-77,225)."""
import torch

from .. import _lib
from ._base import EnvDx


class RocketDx(EnvDx):
    _dilqr_kind = _lib.DYN_ROCKET

    def __init__(self, params=None):
        super().__init__()
        self.n_state, self.n_ctrl = 13, 3
        self.dt = 0.1
        if params is None:
            self.params = torch.tensor((0.5, 1.0, 1.0, 1.0, 1.0), requires_grad=True)
        else:
            self.params = params
        assert len(self.params) == 5
        self.goal_state = torch.zeros(13)
        self.goal_state[6] = 1.0
        self.goal_weights = torch.ones(13)
        self.goal_weights[0:3] = 10.0
        self.goal_weights[6:10] = 0.1
        self.side_penalty, self.thrust_penalty = 1, 0.4
        self.ctrl_penalty = torch.tensor([1.0, 1.0, 0.4])
        self.tilt_penalty = 50.0
        self.max_thrust = 20 ** 2
        self.max_tilt_angle = 0.3
        self.mpc_eps = 1e-3
        self.linesearch_decay = 0.2
        self.max_linesearch_iter = 5
        self.tilt_Q = self.tilt_penalty * torch.tensor([0., 0., 4., 4.])
        self.tilt_p = self.tilt_penalty * torch.tensor([0., 0., 0., 0.])
        # the reference stores tensor bounds (rocket.py:80) that only its float-bound
        # code paths can use; floats here
        self.lower, self.upper = -20.0, 20.0

    def get_true_obj(self):                            # rocket.py:212-232
        Q = torch.cat((self.goal_weights, self.ctrl_penalty))
        Q[6:10] = self.tilt_Q * self.tilt_penalty
        px = -torch.sqrt(self.goal_weights) * self.goal_state
        px[6:10] = -self.tilt_p * self.tilt_penalty
        p = torch.cat((px, torch.zeros(self.n_ctrl)))
        return Q, p
