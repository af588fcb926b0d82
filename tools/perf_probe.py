"""Quick timing probe of the forward solve (cartpole T=50)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
dev = torch.device("cuda:0")

def make(B, T, dtype, sigma=0.5):
    g = torch.Generator().manual_seed(0)
    r = (torch.rand(B, 4, generator=g, dtype=torch.float64) * 2 - 1) * sigma
    x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1).to(dtype).to(dev)
    dx = env.CartpoleDx(torch.tensor((9.8, 1.0, 0.1, 0.5), dtype=dtype, device=dev))
    q, p = dx.get_true_obj()
    C = torch.diag(q).to(dtype).to(dev)[None, None].repeat(T, B, 1, 1)
    c = p.to(dtype).to(dev)[None, None].repeat(T, B, 1)
    return x0, C, c, dx

for dtype in (torch.float64, torch.float32):
    B, T = int(os.environ.get("B", 65536)), 50
    x0, C, c, dx = make(B, T, dtype)
    m = d.MPC(5, 1, T, u_lower=dx.lower, u_upper=dx.upper, lqr_iter=10, verbose=-1,
              exit_unconverged=False, detach_unconverged=False, eps=dx.mpc_eps,
              linesearch_decay=dx.linesearch_decay, max_linesearch_iter=dx.max_linesearch_iter)
    for it in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        e0.record()
        with torch.no_grad():
            x, u, costs = m(x0, d.QuadCost(C, c), dx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        info = m.last_info
        print(f"{dtype} B={B}: {ms:.2f} ms  iters {info.n_iters} retries {info.retries} "
              f"-> {B/ms*1e3/1e6:.3f} M fwd-solves/s; mean alpha {info.mean_alpha:.3f} max du {info.max_full_du:.3e} cost {info.mean_best_cost:.4f}")
