"""Reference harness -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Imports the *unmodified* reference (josef-w/Differentiable-iLQR) from
``/root/reference`` in-process so that (a) the oracle port in
``oracle/port.py`` can be validated against the real thing and (b) golden
vectors can be generated (``tests/golden/make_golden.py``).

``/root/reference`` only exists in the build container, never on the GPU box,
so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this
module.  Recipe follows SURVEY.md section 8(c):

  1. stub the optional plotting / gym / casadi imports the reference does at
     module import time (matplotlib, casadi, gym, IPython, setproctitle);
  2. put /root/reference on sys.path, never write bytecode there;
  3. import ``mpc`` before ``lqr_step`` (circular import in the reference);
  4. harness-side monkeypatch: ``mpc_explicit.MPC.linearize_dynamics`` returns
     detached F, f when ``diff=True`` (torch>=2 refuses the leaf-moved-into-
     graph pattern; the reference's backward returns None for F, f anyway).
"""
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("DILQR_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
    "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d",
    "mpl_toolkits.mplot3d.art3d", "casadi", "gym", "IPython", "IPython.core",
    "IPython.core.ultratb", "setproctitle", "ipdb",
]

_loaded = None


def available():
    return os.path.isdir(REF_ROOT) and os.path.isfile(os.path.join(REF_ROOT, "mpc.py"))


def load():
    """Import the reference modules; returns a namespace of modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    for m in _STUBS:
        if m not in sys.modules:
            s = MagicMock(name=m)
            s.__all__ = []
            sys.modules[m] = s
    # The product package ships same-named modules (mpc, util, ...) for drop-in
    # use; make sure the reference copies win inside this harness.
    saved = {}
    names = ["definitions", "util", "pnqp", "dynamics", "mpc", "lqr_step",
             "mpc_backup", "lqr_step_backup", "lqr_step_explicit",
             "mpc_explicit", "env_dx", "env_dx.pendulum", "env_dx.cartpole",
             "env_dx.rocket", "il_env"]
    for n in names:
        if n in sys.modules:
            saved[n] = sys.modules.pop(n)
    sys.path.insert(0, REF_ROOT)
    try:
        import mpc  # noqa: F401  (must precede lqr_step)
        import lqr_step
        import mpc_backup
        import lqr_step_backup
        import lqr_step_explicit
        import mpc_explicit
        import pnqp
        import util
        import definitions
        from env_dx import pendulum, cartpole, rocket
        import il_env
    finally:
        sys.path.remove(REF_ROOT)

    _orig = mpc_explicit.MPC.linearize_dynamics

    def _lin(self, x, u, dyn, diff):
        F, f = _orig(self, x, u, dyn, diff)
        return (F.detach(), f.detach()) if diff else (F, f)

    mpc_explicit.MPC.linearize_dynamics = _lin

    ns = types.SimpleNamespace(
        mpc=mpc, lqr_step=lqr_step, mpc_backup=mpc_backup,
        lqr_step_backup=lqr_step_backup, lqr_step_explicit=lqr_step_explicit,
        mpc_explicit=mpc_explicit, pnqp=pnqp, util=util,
        definitions=definitions, pendulum=pendulum, cartpole=cartpole,
        rocket=rocket, il_env=il_env)
    # Move the reference modules out of the global module table under a private
    # prefix so a later ``import mpc`` by the product's drop-in layer does not
    # pick them up (and vice versa).
    ns._modules = {}
    for n in names:
        if n in sys.modules:
            ns._modules[n] = sys.modules[n]
    _loaded = ns
    return ns
