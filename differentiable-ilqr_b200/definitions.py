"""API containers (reference definitions.py:3-4)."""
from collections import namedtuple

QuadCost = namedtuple("QuadCost", "C c")
LinDx = namedtuple("LinDx", "F f")
QuadCost.__new__.__defaults__ = (None,) * len(QuadCost._fields)   # mpc.py:25
LinDx.__new__.__defaults__ = (None,) * len(LinDx._fields)         # mpc.py:26
