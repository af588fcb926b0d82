"""Host-side mirror of the reference interface: argument handling that needs no GPU."""
import importlib

import pytest
import torch


def test_constructor_defaults_match_reference(dilqr):
    m = dilqr.MPC(3, 1, 20)
    ref = dict(u_lower=None, u_upper=None, u_zero_I=None, u_init=None, lqr_iter=10, delta_u=None,
               verbose=0, eps=1e-7, back_eps=1e-7, n_batch=None, linesearch_decay=0.2,
               max_linesearch_iter=10, exit_unconverged=True, detach_unconverged=True,
               backprop=True, slew_rate_penalty=None, prev_ctrl=None, not_improved_lim=5,
               best_cost_eps=1e-4)                        # mpc.py:123-144
    for k, v in ref.items():
        assert getattr(m, k) == v, k
    assert m.grad_method == dilqr.GradMethods.ANALYTIC


def test_bounds_must_come_in_pairs(dilqr):
    with pytest.raises(AssertionError):
        dilqr.MPC(3, 1, 20, u_lower=-1.0)                # mpc.py:146
    with pytest.raises(AssertionError):
        dilqr.MPC(3, 1, 20, max_linesearch_iter=0)       # mpc.py:147


def test_cost_expansion_shapes(dilqr):
    m = dilqr.MPC(3, 1, 5)
    C, c = m._expand_cost(dilqr.QuadCost(torch.eye(4), torch.zeros(4)), 7)     # mpc.py:205-226
    assert tuple(C.shape) == (5, 7, 4, 4) and tuple(c.shape) == (5, 7, 4)
    C, c = m._expand_cost(dilqr.QuadCost(torch.eye(4).repeat(5, 1, 1), torch.zeros(5, 4)), 7)
    assert tuple(C.shape) == (5, 7, 4, 4) and tuple(c.shape) == (5, 7, 4)
    with pytest.raises(SystemExit):
        m._expand_cost(dilqr.QuadCost(torch.zeros(4), torch.zeros(4)), 7)


def test_batch_size_inference_error(dilqr):
    m = dilqr.MPC(3, 1, 5)
    with pytest.raises(SystemExit):                      # mpc.py:198-199
        m(torch.zeros(2, 3), dilqr.QuadCost(torch.eye(4), torch.zeros(4)),
          dilqr.LinDx(torch.zeros(4, 2, 3, 4), None))


def test_explicit_rejects_lindx_like_reference(dilqr):
    m = dilqr.mpc_explicit.MPC(3, 1, 5, n_batch=2)
    with pytest.raises(AttributeError):                  # mpc_explicit.py:325
        m(torch.zeros(2, 3), dilqr.QuadCost(torch.eye(4), torch.zeros(4)),
          dilqr.LinDx(torch.zeros(4, 2, 3, 4), None))


def test_env_models_expose_reference_attributes():
    env = importlib.import_module("differentiable-ilqr_b200.env_dx")
    cp, pd = env.CartpoleDx(), env.PendulumDx()
    assert (cp.n_state, cp.n_ctrl, cp.lower, cp.upper) == (5, 1, -100.0, 100.0)
    assert (cp.mpc_eps, cp.linesearch_decay, cp.max_linesearch_iter, cp.dt) == (1e-4, 0.5, 2, 0.05)
    assert (pd.n_state, pd.n_ctrl, pd.lower, pd.upper) == (3, 1, -2.0, 2.0)
    assert (pd.mpc_eps, pd.linesearch_decay, pd.max_linesearch_iter, pd.dt) == (1e-3, 0.2, 5, 0.05)
    q, p = cp.get_true_obj()                             # cartpole.py:859-867
    assert torch.allclose(q, torch.tensor([0.1, 0.1, 1., 1., 0.1, 0.001]))
    assert torch.allclose(p, torch.tensor([0., 0., -1., 0., 0., 0.]))
    q, p = pd.get_true_obj()                             # pendulum.py:117-125
    assert torch.allclose(q, torch.tensor([1., 1., 0.1, 0.001]))
    assert torch.allclose(p, torch.tensor([-1., 0., 0., 0.]))


def test_unsupported_features_fail_loudly(dilqr):
    m = dilqr.mpc_explicit.MPC(3, 1, 5, slew_rate_penalty=1.0)
    with pytest.raises(NotImplementedError):       # slew rate: KKT path (mpc.MPC) only
        m(torch.zeros(2, 3), dilqr.QuadCost(torch.eye(4), torch.zeros(4)), dilqr.env_dx.PendulumDx())
    m = dilqr.MPC(3, 1, 5, u_lower=-1.0, u_upper=1.0, delta_u=0.1)   # trust region: supported
    assert m.delta_u == 0.1


def test_bytes_per_solve_model():
    """cartpole T=50 L=10 -> 235.0 KB (fp64) / 117.5 KB (fp32)."""
    import bench
    it, bwd, tot = bench.bytes_per_solve(8)
    assert it == 2457 * 8 and bwd == 4809 * 8 and tot == 235032
    assert bench.bytes_per_solve(4)[2] == 117516


def test_affine_dynamics_container(dilqr):
    """dynamics.py:159-202: forward / grad_input / the LinDx view handed to the solver."""
    A, B, c = torch.randn(3, 3), torch.randn(3, 2), torch.randn(3)
    d = dilqr.AffineDynamics(A, B, c)
    x, u = torch.randn(5, 3), torch.randn(5, 2)
    assert torch.allclose(d(x, u), x @ A.t() + u @ B.t() + c)
    assert d(x[0], u[0]).shape == (3,)
    R, S = d.grad_input(x, u)
    assert R.shape == (5, 3, 3) and S.shape == (5, 3, 2) and torch.equal(R[2], A)
    F, f = d.as_lindx(4, 5, torch.float32, "cpu")
    assert F.shape == (3, 5, 3, 5) and f.shape == (3, 5, 3)
    assert torch.equal(F[1, 2], torch.cat((A, B), 1)) and torch.equal(f[0, 0], c)
    p = dilqr.CtrlPassthroughDynamics(d)
    xt = torch.cat((torch.randn(5, 2), x), 1)
    assert torch.allclose(p(xt, u), torch.cat((u, d(x, u)), 1))


@pytest.mark.parametrize("act", ["sigmoid", "relu"])
def test_nn_dynamics_container(dilqr, act):
    """dynamics.py:15-130: grad_input is the Jacobian of forward (checked with autograd);
    the device packing is W1 b1 W2 b2."""
    torch.manual_seed(0)
    d = dilqr.NNDynamics(3, 2, hidden_sizes=[7], activation=act).double()
    x, u = torch.randn(4, 3, dtype=torch.float64), torch.randn(4, 2, dtype=torch.float64)
    y = d(x, u)
    assert y.shape == (4, 3) and d(x[0], u[0]).shape == (3,)
    d(x, u)
    R, S = d.grad_input(x, u)
    J = torch.autograd.functional.jacobian(lambda a, b: d(a, b), (x, u))
    for b in range(4):
        assert torch.allclose(R[b], J[0][b, :, b, :], atol=1e-12)
        assert torch.allclose(S[b], J[1][b, :, b, :], atol=1e-12)
    buf, ints = d._dilqr_pack(torch.float64, "cpu")
    assert buf.numel() == 7 * 5 + 7 + 3 * 7 + 3 and ints[0] == 7 and ints[2] == 1
    assert torch.equal(buf[:35].view(7, 5), d.fcs[0].weight.detach())
    with pytest.raises(NotImplementedError):
        dilqr.NNDynamics(3, 2, hidden_sizes=[4, 4, 4])._dilqr_pack(torch.float64, "cpu")
    b2, i2 = dilqr.NNDynamics(3, 2, hidden_sizes=[9, 4])._dilqr_pack(torch.float64, "cpu")
    assert i2[0] == (9 | (4 << 16)) and b2.numel() == 9 * 5 + 9 + 4 * 9 + 4 + 3 * 4 + 3
