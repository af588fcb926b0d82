from . import cartpole, pendulum, rocket  # noqa: F401
from .cartpole import CartpoleDx
from .pendulum import PendulumDx
from .rocket import RocketDx
