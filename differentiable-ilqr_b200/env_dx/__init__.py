from . import cartpole, pendulum  # noqa: F401
from .cartpole import CartpoleDx
from .pendulum import PendulumDx
