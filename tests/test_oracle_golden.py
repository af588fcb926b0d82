"""The CPU oracle (oracle/port.py) against the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py) -- this is what pins it."""
import pytest
import torch

from common import golden, rel


def _mpc_kw(pdx, lqr_iter, eps=None):
    return dict(u_lower=pdx.lower, u_upper=pdx.upper, lqr_iter=lqr_iter,
                eps=pdx.mpc_eps if eps is None else eps, linesearch_decay=pdx.linesearch_decay,
                max_linesearch_iter=pdx.max_linesearch_iter)


def test_cartpole_fixture_bit_exact(port):
    """data/cartpole.pkl (reference's own fixture): reproduced with max diff 0.0."""
    g = golden("fixture_cartpole.npz")
    tau = g["tau"]
    xs, us = tau[:, :, :5].transpose(0, 1), tau[:, :, 5:].transpose(0, 1)
    T, B = int(g["mpc_T"]), tau.shape[0]
    pdx = port.CartpoleDx()
    q, p = pdx.get_true_obj()
    C = torch.diag(q)[None, None].repeat(T, B, 1, 1)
    c = p[None, None].repeat(T, B, 1)
    o = port.mpc_forward(xs[0].clone(), port.QuadCost(C, c), pdx, 5, 1, T, final_pass=False,
                         **_mpc_kw(pdx, 40))
    # 40 of the fixture's 100 iterations keep this test fast; the best iterate is
    # reached well before (full 100 iterations give exactly the same tensors).
    o_full_equal = float((o.u - us).abs().max()) == 0.0 and float((o.x - xs).abs().max()) == 0.0
    if not o_full_equal:
        o = port.mpc_forward(xs[0].clone(), port.QuadCost(C, c), pdx, 5, 1, T, final_pass=False,
                             **_mpc_kw(pdx, int(g["lqr_iter"])))
    assert float((o.u - us).abs().max()) == 0.0
    assert float((o.x - xs).abs().max()) == 0.0


def test_pendulum_fixture(port):
    """data/pendulum.pkl: the reference itself reproduces it to 7e-4 (fp32)."""
    g = golden("fixture_pendulum.npz")
    tau = g["tau"]
    xs, us = tau[:, :, :3].transpose(0, 1), tau[:, :, 3:].transpose(0, 1)
    T, B = int(g["mpc_T"]), tau.shape[0]
    pdx = port.PendulumDx()
    q, p = pdx.get_true_obj()
    C = torch.diag(q)[None, None].repeat(T, B, 1, 1)
    c = p[None, None].repeat(T, B, 1)
    o = port.mpc_forward(xs[0].clone(), port.QuadCost(C, c), pdx, 3, 1, T, final_pass=False,
                         **_mpc_kw(pdx, int(g["lqr_iter"])))
    assert float((o.u - us).abs().max()) < 1e-3
    assert float((o.x - xs).abs().max()) < 2e-4
    # stored trajectories are dynamics-consistent
    nx = pdx(xs[:-1].reshape(-1, 3), us[:-1].reshape(-1, 1)).reshape(T - 1, B, 3)
    assert float((nx - xs[1:]).abs().max()) < 1e-5


@pytest.mark.parametrize("tag", ["free", "boxed"])
def test_lindx_forward_and_kkt_backward(port, tag):
    g = golden("ref_lindx_%s.npz" % tag)
    kw = dict(u_lower=-1.0, u_upper=1.0) if tag == "boxed" else {}
    o = port.mpc_forward(g["x0"], port.QuadCost(g["C"], g["c"]), port.LinDx(g["F"], g["f"]), 4, 2,
                         12, lqr_iter=20, **kw)
    assert float((o.x - g["x"]).abs().max()) == 0.0
    assert float((o.u - g["u"]).abs().max()) == 0.0
    assert float((o.costs - g["costs"]).abs().max()) == 0.0
    k = port.kkt_backward(g["gx"], g["gu"], g["x0"], g["C"], g["c"], g["F"], g["f"], o.x, o.u, 4,
                          2, **kw)
    for nm, a in (("dx0", k.dx_init), ("dC", k.dC), ("dc", k.dc), ("dF", k.dF), ("df", k.df)):
        assert rel(a, g[nm]) < 1e-13, nm


def test_dilqr_backward_at_the_headline_horizon(port):
    """Cartpole T=50 (the benchmark horizon), warm-started converged regime:
    the oracle's matrix-free backward against the reference's dense fix_point_equ solve --
    and how few Richardson passes that regime needs."""
    g = golden("ref_dilqr_cartpole_T50.npz")
    T, B = int(g["T"]), g["x0"].shape[0]
    pdx = port.CartpoleDx(params=g["theta"], dtype=torch.float64)
    C = torch.diag(g["q"])[None, None].repeat(T, B, 1, 1)
    c = g["p"][None, None].repeat(T, B, 1)
    o = port.mpc_forward(g["x0"], port.QuadCost(C, c), pdx, 5, 1, T, u_init=g["u_init"],
                         final_pass=False, **_mpc_kw(pdx, int(g["lqr_iter"]), eps=1e-9))
    assert o.n_iters == 1                    # the reference stops after one iteration too
    assert rel(o.x, g["x"]) < 1e-13 and rel(o.u, g["u"]) < 1e-13
    for n_passes in (4, 12):
        r = port.dilqr_backward(g["gx"], g["gu"], g["x0"], C, c, o.x, o.u, pdx, 5, 1, pdx.lower,
                                pdx.upper, n_passes=n_passes)
        assert rel(r.dtheta.sum(0), g["dtheta"]) < 1e-10
        assert rel(r.dC, g["dC"]) < 1e-12 and rel(r.dc, g["dc"]) < 1e-12


@pytest.mark.parametrize("env", ["pendulum", "cartpole", "rocket"])
def test_dilqr_backward_matches_reference_dense_solve(port, env):
    """Matrix-free DiLQR gradient == the reference's fix_point_equ (dense solve)."""
    g = golden("ref_dilqr_%s.npz" % env)
    dt = torch.float64
    pdx = {"pendulum": port.PendulumDx, "cartpole": port.CartpoleDx,
           "rocket": port.RocketDx}[env](params=g["theta"], dtype=dt)
    T, B = int(g["T"]), g["x0"].shape[0]
    C = torch.diag(g["q"])[None, None].repeat(T, B, 1, 1)
    c = g["p"][None, None].repeat(T, B, 1)
    o = port.mpc_forward(g["x0"], port.QuadCost(C, c), pdx, pdx.n_state, pdx.n_ctrl, T,
                         final_pass=False, **_mpc_kw(pdx, int(g["lqr_iter"]), eps=1e-9))
    # pendulum / cartpole: bit-exact; rocket's Jacobian is the generated (CSE-reordered)
    # form of the reference's expressions, equal to round-off
    assert float((o.x - g["x"]).abs().max()) <= (1e-14 if env == "rocket" else 0.0)
    d = port.dilqr_backward(g["gx"], g["gu"], g["x0"], C, c, o.x, o.u, pdx, pdx.n_state,
                            pdx.n_ctrl, pdx.lower, pdx.upper, n_passes=80, tol=1e-15)
    assert rel(d.dtheta.sum(0), g["dtheta"]) < 1e-10
    assert rel(d.dC, g["dC"]) < 1e-10
    assert rel(d.dc, g["dc"]) < 1e-10


@pytest.mark.parametrize("env", ["pendulum", "cartpole", "rocket"])
def test_tables_and_first_order(port, env):
    """Generated get_matrices tables + step + analytic Jacobian vs the reference."""
    import env_tables_gen as G
    g = golden("ref_tables.npz")
    x, u, th = g[env + "_x"], g[env + "_u"], g[env + "_theta"]
    out = getattr(G, env + "_tables")(x, u, th)
    for nm, a in zip(["D", "D_theta", "D_x", "D_u", "x_theta", "x_x", "x_u"], out):
        assert rel(a, g[env + "_" + nm]) < 1e-13, nm
    pdx = {"pendulum": port.PendulumDx, "cartpole": port.CartpoleDx,
           "rocket": port.RocketDx}[env](params=th, dtype=torch.float64)
    assert rel(pdx(x, u), g[env + "_step"]) < 1e-15
    assert rel(pdx.get_linear_dyn(x, u), g[env + "_lin"]) < 1e-13


@pytest.mark.parametrize("name", ["ref_fwd_cartpole_f64", "ref_fwd_pendulum_f64",
                                  "ref_fwd_cartpole_f32"])
def test_env_forward(port, name):
    g = golden(name + ".npz")
    dtype = g["x0"].dtype
    env = "cartpole" if "cartpole" in name else "pendulum"
    pdx = (port.PendulumDx if env == "pendulum" else port.CartpoleDx)(dtype=dtype)
    T, B = int(g["T"]), g["x0"].shape[0]
    C = torch.diag(g["q"])[None, None].repeat(T, B, 1, 1)
    c = g["p"][None, None].repeat(T, B, 1)
    o = port.mpc_forward(g["x0"], port.QuadCost(C, c), pdx, pdx.n_state, pdx.n_ctrl, T,
                         final_pass=False, **_mpc_kw(pdx, int(g["lqr_iter"])))
    assert float((o.x - g["x"]).abs().max()) == 0.0
    assert float((o.u - g["u"]).abs().max()) == 0.0


@pytest.mark.parametrize("name,n_samples", [("pendulum", 2), ("cartpole", 1)])
def test_closed_loop_bit_exact(port, name, n_samples):
    """IL_Env.populate_data2 (il_env.py:96-151) run through the unmodified reference:
    sample_xinit and the receding-horizon closed loop are reproduced with max diff 0.0."""
    g = golden("ref_closed_loop_%s.npz" % name)
    cls = port.PendulumDx if name == "pendulum" else port.CartpoleDx
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float64)
    try:
        x0 = port.sample_xinit(name, g["x0"].shape[0])
    finally:
        torch.set_default_dtype(torch.float32)
    assert float((x0 - g["x0"]).abs().max()) == 0.0
    ref = torch.cat((g["train"], g["val"], g["test"]))[:n_samples]
    tau = port.closed_loop(cls(dtype=torch.float64), g["x0"][:n_samples], int(g["mpc_T"]),
                           int(g["lqr_iter"]))
    assert float((tau - ref).abs().max()) == 0.0


@pytest.mark.parametrize("act", ["sigmoid", "relu"])
def test_nn_dynamics_bit_exact(port, act):
    """mpc.MPC with dynamics.NNDynamics run through the unmodified reference
    (tests/golden/make_golden.py nn_dynamics): the oracle's network dynamics reproduce
    the trajectories with max diff 0.0."""
    g = golden("ref_nn_dynamics.npz")
    t = lambda k: g[act + "_" + k]
    dyn = port.NNDynamics(t("W1"), t("b1"), t("W2"), t("b2"), activation=act)
    o = port.mpc_forward(t("x0"), port.QuadCost(t("C"), t("c")), dyn, 3, 1, 10, u_lower=-1.0,
                         u_upper=1.0, lqr_iter=30, final_pass=False)
    assert float((o.x - t("x")).abs().max()) == 0.0
    assert float((o.u - t("u")).abs().max()) == 0.0


def test_delta_u_bit_exact(port):
    """mpc.MPC(delta_u=0.25) of the unmodified reference on a boxed LinDx problem
    (lqr_step.py:132-134, 204-211): reproduced with max diff 0.0."""
    g = golden("ref_delta_u.npz")
    for L in (1, 3, 25):
        o = port.mpc_forward(g["x0"], port.QuadCost(g["C"], g["c"]), port.LinDx(g["F"], g["f"]),
                             4, 2, 10, u_lower=-1.0, u_upper=1.0, lqr_iter=L, final_pass=False,
                             delta_u=0.25)
        assert float((o.u - g["L%d_u" % L]).abs().max()) == 0.0
        assert float((o.x - g["L%d_x" % L]).abs().max()) == 0.0
