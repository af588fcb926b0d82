#!/usr/bin/env python
"""Generate csrc/env_tables_gen.cuh: the second-order dynamics tables of the
reference's ``get_matrices`` (env_dx/cartpole.py:105-716, env_dx/pendulum.py:152-382).

Why a generator: SURVEY 8a-10 -- several entries of those tables are NOT the
true derivatives of the dynamics (typos in the reference), and the DiLQR
gradient of the reference is defined *by* them, so they cannot be re-derived;
they have to be reproduced entry by entry.  This tool executes the reference's
own ``get_matrices`` on SYMBOLIC inputs (a tiny object-array stand-in for the
handful of torch calls that function makes), flattens every table entry to a
sympy expression, runs common-subexpression elimination over all of them and
prints plain CUDA device code.  Nothing of the reference's source text is
copied; the output is straight-line arithmetic over CSE temporaries.

Run in the build container only (needs /root/reference):
    python tools/gen_env_tables.py
The generated header is committed; tests/test_env_tables.py (gpu) checks it
against golden outputs of the reference's get_matrices at random (x,u,theta).
"""
import os
import sys

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter
from sympy.printing.precedence import PRECEDENCE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402


class SymT(np.ndarray):
    """Object ndarray of sympy expressions with the few Tensor methods used."""

    def unsqueeze(self, d):
        return np.expand_dims(self, d).view(SymT)

    def squeeze(self, d=None, axis=None):
        if d is None:
            d = axis
        return np.squeeze(np.asarray(self, dtype=object), axis=d).view(SymT)

    def permute(self, *dims):
        return np.transpose(self, dims).view(SymT)

    def detach(self):
        return self

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return np.ndarray.reshape(self, shape).view(SymT)

    @property
    def device(self):
        return "sym"


def sym(a):
    return np.asarray(a, dtype=object).view(SymT)


def _vec(fn):
    return lambda a, *r: sym(np.vectorize(fn, otypes=[object])(a, *r))


class TorchShim:
    """Stand-in for the ``torch`` module inside the reference's env files."""
    sin = staticmethod(_vec(sp.sin))
    cos = staticmethod(_vec(sp.cos))
    atan2 = staticmethod(_vec(sp.atan2))

    @staticmethod
    def ones_like(a):
        return sym(np.full(np.shape(a), sp.Integer(1), dtype=object))

    @staticmethod
    def zeros_like(a):
        return sym(np.full(np.shape(a), sp.Integer(0), dtype=object))

    @staticmethod
    def zeros(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return sym(np.full(shape, sp.Integer(0), dtype=object))

    @staticmethod
    def ones(*shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return sym(np.full(shape, sp.Integer(1), dtype=object))

    @staticmethod
    def stack(seq, dim=0):
        seq = [np.asarray(s, dtype=object) for s in seq]
        return sym(np.stack(seq, axis=dim))

    @staticmethod
    def cat(seq, dim=0):
        seq = [np.asarray(s, dtype=object) for s in seq]
        return sym(np.concatenate(seq, axis=dim))

    @staticmethod
    def unbind(a, dim=0):
        a = np.asarray(a, dtype=object)
        return tuple(np.take(a, i, axis=dim) if a.ndim > 1 else a[i] for i in range(a.shape[dim]))

    @staticmethod
    def tensor(data, **kw):
        return sym(np.array(data, dtype=object))

    Tensor = SymT


class ParamVec(SymT):
    pass


def symbolic_tables(modname, cls, state_syms, ctrl_syms, param_syms):
    R = ref_harness.load()
    mod = getattr(R, modname)
    dx = getattr(mod, cls)()
    x = sym([state_syms])          # [1, ns]
    u = sym([ctrl_syms])           # [1, nc]
    params = sym(param_syms)

    class P(SymT):
        pass
    dx.params = params.view(P)
    P.detach = lambda self: self
    P.__iter__ = lambda self: iter(np.asarray(self, dtype=object).tolist())
    real_torch = mod.torch
    mod.torch = TorchShim
    try:
        out = dx.get_matrices(x, u)
    finally:
        mod.torch = real_torch
    names = ["D", "D_theta", "D_x", "D_u", "x_theta", "x_x", "x_u"]
    tables = {}
    for nm, t in zip(names, out):
        a = np.asarray(t, dtype=object)
        assert a.shape[0] == 1, (nm, a.shape)
        tables[nm] = a[0]
    return tables


class CPrinter(C99CodePrinter):
    def _print_Pow(self, e):
        b, ex = e.as_base_exp()
        if ex.is_Integer and 2 <= int(ex) <= 4:
            return "(" + "*".join([self.parenthesize(b, PRECEDENCE["Mul"])] * int(ex)) + ")"
        if ex.is_Integer and -4 <= int(ex) <= -1:
            inner = "*".join([self.parenthesize(b, PRECEDENCE["Mul"])] * (-int(ex)))
            return "(S(1)/(" + inner + "))"
        return super()._print_Pow(e)

    def _print_Float(self, e):
        return "S(%s)" % repr(float(e))

    def _print_Integer(self, e):
        return "S(%d)" % int(e)

    def _print_Rational(self, e):
        return "S(%d.0/%d.0)" % (e.p, e.q)

    def _print_Function(self, e):
        nm = e.func.__name__
        if nm in ("sin", "cos", "atan2"):
            return "%sS<S>(%s)" % (nm, ", ".join(self._print(a) for a in e.args))
        return super()._print_Function(e)


def emit(name, tables, state_syms, ctrl_syms, param_syms, ns, nc, nth, out):
    n = ns + nc
    flat = []
    index = []
    for nm in ["D", "D_theta", "D_x", "D_u", "x_theta", "x_x", "x_u"]:
        a = tables[nm]
        for idx in np.ndindex(a.shape):
            e = sp.sympify(a[idx])
            flat.append(e)
            index.append((nm, idx))
    print("  [%s] %d entries, %d non-zero; running CSE ..." % (
        name, len(flat), sum(1 for e in flat if e != 0)), file=sys.stderr)
    repl, red = sp.cse(flat, symbols=sp.numbered_symbols("t"), optimizations=None, order="none")
    pr = CPrinter()
    w = out.write
    w("// ---- %s: generated by tools/gen_env_tables.py -- do not edit ----\n" % name)
    w("template <class S>\nstruct EnvTables<S, DYN_%s> {\n" % name.upper())
    w("  static constexpr int NS = %d, NC = %d, N = %d, NTH = %d;\n" % (ns, nc, n, nth))
    w("  // x: state, u: control, P.p: theta.  Outputs (row-major):\n")
    w("  //  D[NS][N], Dth[NS][N][NTH], Dx[NS][N][NS], Du[NS][N][NC], xth[NS][NTH], xx[NS][NS], xu[NS][NC]\n")
    w("  DILQR_DEVICE static void eval(const DynParams<S>& P, const S* xs_, const S* us_,\n")
    w("      S (*D)[N], S (*Dth)[N][NTH], S (*Dx)[N][NS], S (*Du)[N][NC], S (*xth)[NTH],\n")
    w("      S (*xx)[NS], S (*xu)[NC]) {\n")
    used = set().union(*[e.free_symbols for e in flat]) if flat else set()
    for i, s_ in enumerate(state_syms):
        if s_ in used:
            w("    const S %s = xs_[%d];\n" % (s_, i))
    for i, s_ in enumerate(ctrl_syms):
        if s_ in used:
            w("    const S %s = us_[%d];\n" % (s_, i))
    for i, s_ in enumerate(param_syms):
        if s_ in used:
            w("    const S %s = P.p[%d];\n" % (s_, i))
    for sym_, e in repl:
        w("    const S %s = %s;\n" % (sym_, pr.doprint(e)))
    cname = {"D": "D", "D_theta": "Dth", "D_x": "Dx", "D_u": "Du", "x_theta": "xth",
             "x_x": "xx", "x_u": "xu"}
    for (nm, idx), e in zip(index, red):
        w("    %s%s = %s;\n" % (cname[nm], "".join("[%d]" % i for i in idx), pr.doprint(e)))
    w("  }\n")
    # first-order Jacobian only (the iLQR linearisation)
    Dflat = [sp.sympify(tables["D"][idx]) for idx in np.ndindex(tables["D"].shape)]
    replD, redD = sp.cse(Dflat, symbols=sp.numbered_symbols("d"), optimizations=None, order="none")
    w("  DILQR_DEVICE static void eval_D(const DynParams<S>& P, const S* xs_, const S* us_, S (*D)[N]) {\n")
    usedD = set().union(*[e.free_symbols for e in Dflat])
    for i, s_ in enumerate(state_syms):
        if s_ in usedD:
            w("    const S %s = xs_[%d];\n" % (s_, i))
    for i, s_ in enumerate(ctrl_syms):
        if s_ in usedD:
            w("    const S %s = us_[%d];\n" % (s_, i))
    for i, s_ in enumerate(param_syms):
        if s_ in usedD:
            w("    const S %s = P.p[%d];\n" % (s_, i))
    for sym_, e in replD:
        w("    const S %s = %s;\n" % (sym_, pr.doprint(e)))
    for idx, e in zip(np.ndindex(tables["D"].shape), redD):
        w("    D%s = %s;\n" % ("".join("[%d]" % i for i in idx), pr.doprint(e)))
    w("  }\n")
    # structural non-zero masks (compile-time pruning of the contractions)
    for nm in ["D", "D_theta", "D_x", "D_u", "x_theta", "x_x", "x_u"]:
        a = tables[nm]
        args = ", ".join("int i%d" % d for d in range(a.ndim))
        terms = []
        for idx in np.ndindex(a.shape):
            if sp.sympify(a[idx]) != 0:
                terms.append("(" + " && ".join("i%d == %d" % (d, v) for d, v in enumerate(idx)) + ")")
        w("  __host__ __device__ static constexpr bool nz_%s(%s) {\n    return %s;\n  }\n" % (
            cname[nm], args, " ||\n           ".join(terms) if terms else "false"))
    w("};\n\n")


def emit_py(name, tables, state_syms, ctrl_syms, param_syms, out):
    """torch (CPU) version of the same tables for oracle/port.py."""
    from sympy.printing.pycode import PythonCodePrinter

    class PP(PythonCodePrinter):
        def _print_Function(self, e):
            nm = e.func.__name__
            if nm in ("sin", "cos", "atan2"):
                return "torch.%s(%s)" % (nm, ", ".join(self._print(a) for a in e.args))
            return super()._print_Function(e)
        _print_sin = _print_cos = _print_atan2 = _print_Function

    flat, index = [], []
    order = ["D", "D_theta", "D_x", "D_u", "x_theta", "x_x", "x_u"]
    for nm in order:
        a = tables[nm]
        for idx in np.ndindex(a.shape):
            flat.append(sp.sympify(a[idx]))
            index.append((nm, idx))
    repl, red = sp.cse(flat, symbols=sp.numbered_symbols("t"), optimizations=None, order="none")
    pr = PP()
    w = out.write
    w("def %s_tables(x, u, theta):\n" % name)
    w('    """x[N,ns], u[N,nc], theta[nth] -> D, D_theta, D_x, D_u, x_theta, x_x, x_u (batched)."""\n')
    used = set().union(*[e.free_symbols for e in flat])
    for i, s_ in enumerate(state_syms):
        if s_ in used:
            w("    %s = x[:, %d]\n" % (s_, i))
    for i, s_ in enumerate(ctrl_syms):
        if s_ in used:
            w("    %s = u[:, %d]\n" % (s_, i))
    for i, s_ in enumerate(param_syms):
        if s_ in used:
            w("    %s = theta[%d]\n" % (s_, i))
    w("    N = x.shape[0]\n")
    for sym_, e in repl:
        w("    %s = %s\n" % (sym_, pr.doprint(e)))
    for nm in order:
        w("    %s = torch.zeros((N,) + %r, dtype=x.dtype)\n" % (nm, tuple(tables[nm].shape)))
    for (nm, idx), e in zip(index, red):
        if e != 0:
            w("    %s[(slice(None),) + %r] = %s\n" % (nm, tuple(idx), pr.doprint(e)))
    w("    return %s\n\n\n" % ", ".join(order))


def main():
    out_path = os.path.join(ROOT, "differentiable-ilqr_b200", "csrc", "env_tables_gen.cuh")
    with open(out_path, "w") as out:
        out.write("// env_tables_gen.cuh -- GENERATED by tools/gen_env_tables.py (do not edit).\n"
                  "// Second-order dynamics tables of the reference's get_matrices\n"
                  "// (env_dx/cartpole.py:105-716, env_dx/pendulum.py:152-382), reproduced entry by\n"
                  "// entry because the DiLQR gradient is defined by them (SURVEY 8a-10).\n"
                  "#pragma once\n#include \"dynamics.cuh\"\n\nnamespace dilqr {\n\n"
                  "template <class S, int DYN>\nstruct EnvTables;\n\n")
        c, s, w, u = sp.symbols("c s w u", real=True)
        xx_, xd_ = sp.symbols("px pv", real=True)
        g, mc, mp, l, m = sp.symbols("g mc mp l m", real=True, positive=True)
        tabs = symbolic_tables("cartpole", "CartpoleDx", [xx_, xd_, c, s, w], [u], [g, mc, mp, l])
        emit("cartpole", tabs, [xx_, xd_, c, s, w], [u], [g, mc, mp, l], 5, 1, 4, out)
        tabs = symbolic_tables("pendulum", "PendulumDx", [c, s, w], [u], [g, m, l])
        emit("pendulum", tabs, [c, s, w], [u], [g, m, l], 3, 1, 3, out)
        rx = list(sp.symbols("r0 r1 r2 v0 v1 v2 q0 q1 q2 q3 wx wy wz", real=True))
        ru = list(sp.symbols("ux uy uz", real=True))
        rp = list(sp.symbols("Jx Jy Jz mass l", real=True, positive=True))
        tabs = symbolic_tables("rocket", "RocketDx", rx, ru, rp)
        emit("rocket", tabs, rx, ru, rp, 13, 3, 5, out)
        out.write("}  // namespace dilqr\n")
    print("wrote", out_path)
    py_path = os.path.join(ROOT, "oracle", "env_tables_gen.py")
    with open(py_path, "w") as out:
        out.write('"""GENERATED by tools/gen_env_tables.py (do not edit): the reference\'s\n'
                  'get_matrices tables (env_dx/cartpole.py:105-716, env_dx/pendulum.py:152-382) as\n'
                  'straight-line torch code for the CPU oracle.  TEST INFRASTRUCTURE ONLY."""\n'
                  "import torch\n\n\n")
        tabs = symbolic_tables("cartpole", "CartpoleDx", [xx_, xd_, c, s, w], [u], [g, mc, mp, l])
        emit_py("cartpole", tabs, [xx_, xd_, c, s, w], [u], [g, mc, mp, l], out)
        tabs = symbolic_tables("pendulum", "PendulumDx", [c, s, w], [u], [g, m, l])
        emit_py("pendulum", tabs, [c, s, w], [u], [g, m, l], out)
        tabs = symbolic_tables("rocket", "RocketDx", rx, ru, rp)
        emit_py("rocket", tabs, rx, ru, rp, out)
    print("wrote", py_path)


if __name__ == "__main__":
    main()
