"""Drop-in installation: make the reference's own top-level module names resolve to this
package, so that UNMODIFIED reference callers -- ``il_env.py`` (``from mpc_explicit import
MPC``, ``from env_dx import pendulum, cartpole``, il_env.py:5-9), ``il_exp.py``,
``test_mpc.py`` (``from mpc import mpc`` style imports) -- run on the CUDA path.

    import importlib
    importlib.import_module("differentiable-ilqr_b200.dropin").install()
    import il_env            # the reference's file, untouched

Only names of the hot path are aliased; the reference's plotting /
experiment scripts stay what they are."""
import importlib
import sys

_PKG = __name__.rsplit(".", 1)[0]

# reference module name -> module of this package
ALIASES = {
    "definitions": "definitions",
    "mpc": "mpc",
    "mpc_backup": "mpc",                  # the backup copies differ in the gain solve only
    "mpc_explicit": "mpc_explicit",
    "mpc_explicit_backup": "mpc_explicit",
    "lqr_step": "lqr_step",
    "lqr_step_explicit": "lqr_step_explicit",
    "pnqp": "pnqp",
    "util": "util",
    "dynamics": "dynamics",
    "env_dx": "env_dx",
    "env_dx.pendulum": "env_dx.pendulum",
    "env_dx.cartpole": "env_dx.cartpole",
    "env_dx.rocket": "env_dx.rocket",
}


def install(force=False):
    """Register the aliases in ``sys.modules`` (existing entries are kept unless ``force``).
    Returns the dict of names installed."""
    done = {}
    for ref_name, ours in ALIASES.items():
        if ref_name in sys.modules and not force:
            continue
        mod = importlib.import_module(_PKG + "." + ours)
        sys.modules[ref_name] = mod
        done[ref_name] = mod
    return done


def uninstall():
    for ref_name, ours in ALIASES.items():
        mod = sys.modules.get(ref_name)
        if mod is not None and getattr(mod, "__name__", "").startswith(_PKG + "."):
            del sys.modules[ref_name]
