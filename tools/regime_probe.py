"""Regime 2b diagnostics (bench.py --regime 2b): after the pre-solve, how many problems sit
at a fixed point (||du|| < eps), and for how many does the 4-pass Richardson gradient agree
with a 30-pass one."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
solver = importlib.import_module("differentiable-ilqr_b200._solver")
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
x0, uexp = bench.make_inputs(torch, B, torch.float64, 0, sigma)
x0, uexp = x0.to(dev), uexp.to(dev)
proto = env.CartpoleDx()
q, p = [t.double().to(dev) for t in proto.get_true_obj()]
theta = torch.tensor(bench.THETA, dtype=torch.float64, device=dev)
kw = dict(u_lower=proto.lower, u_upper=proto.upper, verbose=-1, exit_unconverged=False,
          detach_unconverged=False, linesearch_decay=proto.linesearch_decay,
          max_linesearch_iter=proto.max_linesearch_iter, eps=proto.mpc_eps, n_batch=B)
pre = d.mpc_explicit.MPC(5, 1, 50, lqr_iter=250, **kw)
with torch.no_grad():
    _, u_warm, _ = pre(x0, d.QuadCost(torch.diag(q), p), env.CartpoleDx(theta))
du = pre.last_info.full_du_norm
print("B=%d sigma=%g: presolve iterations %d, max du %.3e, fraction du<eps %.5f, quantiles %s" % (
    B, sigma, pre.last_info.n_iters, float(du.max()), float((du < 1e-4).float().mean()),
    [float(v) for v in torch.quantile(du, torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.float64, device=dev))]))
res = {}
for passes in (4, 30):
    th = theta.clone().requires_grad_()
    m = d.mpc_explicit.MPC(5, 1, 50, lqr_iter=10, u_init=u_warm, richardson_passes=passes,
                           richardson_tol=None, **kw)
    x, u, _ = m(x0, d.QuadCost(torch.diag(q), p), env.CartpoleDx(th))
    # per-problem dtheta: call the backward directly
    gu = 2 * (u.detach() - uexp) / u.numel()
    dC, dc, dth = solver.dilqr_backward(None, gu, x0, torch.diag(q), p, x.detach(), u.detach(),
                                        env.CartpoleDx(theta), 5, 1, proto.lower, proto.upper,
                                        n_passes=passes)
    res[passes] = dth
    print("passes %d: solve iterations %d, du max %.2e" % (passes, m.last_info.n_iters,
                                                          float(m.last_info.full_du_norm.max())))
e = (res[4] - res[30]).abs().amax(1) / (res[30].abs().amax(1) + 1e-300)
print("per-problem |dtheta(4 passes) - dtheta(30)| / |dtheta(30)|: median %.1e, 90%% %.1e, 99%% %.1e, max %.1e; "
      "fraction < 1e-8: %.4f" % (float(e.median()), float(torch.quantile(e, 0.9)), float(torch.quantile(e, 0.99)),
                                 float(e.max()), float((e < 1e-8).float().mean())))
print("any nan:", bool(torch.isnan(res[30]).any()), "max |dtheta| %.2e" % float(res[30].abs().max()))
