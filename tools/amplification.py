"""How fp64 round-off differences between the CUDA path and the CPU oracle grow
with the number of (unconverged, cold-start) iLQR iterations -- documents why
multi-iteration parity is asserted at 1e-6 while single steps hold 1e-13."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import port
from common import env_problem, golden, rel
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
dev = torch.device("cuda:0")
g = golden("ref_fwd_cartpole_f64.npz")
T, B = int(g["T"]), g["x0"].shape[0]
pdx = port.CartpoleDx(dtype=torch.float64)
q, p = pdx.get_true_obj()
C = torch.diag(q)[None, None].repeat(T, B, 1, 1); c = p[None, None].repeat(T, B, 1)
kw = dict(u_lower=pdx.lower, u_upper=pdx.upper, eps=pdx.mpc_eps, linesearch_decay=pdx.linesearch_decay, max_linesearch_iter=pdx.max_linesearch_iter)
gdx = env.CartpoleDx(pdx.params.to(dev))
for L in range(1, 9):
    o = port.mpc_forward(g["x0"], port.QuadCost(C, c), pdx, 5, 1, T, lqr_iter=L, final_pass=False, **kw)
    m = d.MPC(5, 1, T, lqr_iter=L, verbose=-1, exit_unconverged=False, **kw)
    with torch.no_grad():
        x, u, costs = m(g["x0"].to(dev), d.QuadCost(C.to(dev), c.to(dev)), gdx)
    per = (u.cpu() - o.u).abs().amax((0, 2))
    print(f"L={L}: rel x {rel(x,o.x):.2e} u {rel(u,o.u):.2e} cost {rel(costs,o.costs):.2e}  worst problem {int(per.argmax())} |du| {float(per.max()):.2e} median {float(per.median()):.2e} alphas {m.last_info.mean_alpha:.3f}")
