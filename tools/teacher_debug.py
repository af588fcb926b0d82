"""Per-problem look at one teacher-forced iteration (tests/test_teacher_forced.py): which
problem carries the worst error, what its line-search step / pnqp state was."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import port
from common import env_problem
from test_teacher_forced import _oracle_iterates
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
dev = torch.device("cuda:0")
name, T, B, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dtype = torch.float64 if len(sys.argv) < 6 or sys.argv[5] == "f64" else torch.float32
pdx, x0, C, c, kw = env_problem(port, name, T, B, dtype, sigma=0.5)
its = _oracle_iterates(port, pdx, x0, C, c, T, L, kw)
gdx = {"cartpole": env.CartpoleDx, "pendulum": env.PendulumDx}[name](pdx.params.to(dev))
for k, (u_k, o) in enumerate(its):
    m = d.MPC(pdx.n_state, pdx.n_ctrl, T, lqr_iter=1, u_init=u_k.to(dev), verbose=-1,
              exit_unconverged=False, detach_unconverged=False, **kw)
    with torch.no_grad():
        x, u, costs = m(x0.to(dev), d.QuadCost(C.to(dev), c.to(dev)), gdx)
    eu = (u.cpu() - o.u).abs().amax((0, 2))
    ex = (x.cpu() - o.x).abs().amax((0, 2))
    b = int(torch.maximum(eu, ex).argmax())
    al = o.alphas
    sat = ((o.u.abs() - pdx.upper).abs() < 1e-9).float().mean(0).squeeze(-1)
    med = float(torch.maximum(eu, ex).median())
    print("it %d: worst problem %d err u %.2e x %.2e (median over problems %.1e) | its alpha %.3g "
          "frac saturated %.2f |u|max %.3g cost %.6g | n(alpha<1) %d, qp %s vs %s" % (
              k, b, float(eu[b]), float(ex[b]), med, float(al[b]), float(sat[b]),
              float(o.u[:, b].abs().max()), float(o.costs[b]), int((al < 1).sum()),
              m.last_info.qp_iters, o.n_total_qp_iter))
