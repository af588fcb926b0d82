"""Developer smoke/parity script (GPU box): CUDA path vs the CPU oracle port."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402

d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
dev = torch.device("cuda:0")


def rel(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def lindx_problem(ns, nc, T, B, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = ns + nc
    A = torch.randn(T, B, n, n, generator=g, dtype=torch.float64)
    C = A.transpose(2, 3) @ A + torch.eye(n, dtype=torch.float64)
    c = torch.randn(T, B, n, generator=g, dtype=torch.float64)
    F = torch.cat((torch.eye(ns, dtype=torch.float64).expand(T - 1, B, ns, ns)
                   + 0.2 * torch.randn(T - 1, B, ns, ns, generator=g, dtype=torch.float64) / ns ** 0.5,
                   torch.randn(T - 1, B, ns, nc, generator=g, dtype=torch.float64) / ns ** 0.5), 3)
    f = 0.1 * torch.randn(T - 1, B, ns, generator=g, dtype=torch.float64)
    x0 = torch.randn(B, ns, generator=g, dtype=torch.float64)
    return [t.to(dtype) for t in (C, c, F, f, x0)]


def run_lindx(ns, nc, T, B, dtype, boxed, lqr_iter=20):
    C, c, F, f, x0 = lindx_problem(ns, nc, T, B, dtype)
    kw = dict(u_lower=-1.0, u_upper=1.0) if boxed else {}
    o = port.mpc_forward(x0, port.QuadCost(C, c), port.LinDx(F, f), ns, nc, T,
                         lqr_iter=lqr_iter, **kw)
    m = d.MPC(ns, nc, T, lqr_iter=lqr_iter, verbose=-1, exit_unconverged=False, **kw)
    Cg, cg, Fg, fg, x0g = [t.to(dev).requires_grad_() for t in (C, c, F, f, x0)]
    x, u, costs = m(x0g, d.QuadCost(Cg, cg), d.LinDx(Fg, fg))
    info = m.last_info
    print(f"lindx ns={ns} nc={nc} T={T} B={B} {dtype} boxed={boxed}: iters {info.n_iters}/{o.n_iters} "
          f"retries {info.retries} qp {info.qp_iters} vs {o.qp_iters} "
          f"x {rel(x, o.x):.2e} u {rel(u, o.u):.2e} cost {rel(costs, o.costs):.2e}")
    g = torch.Generator().manual_seed(7)
    gx = torch.randn(o.x.shape, generator=g, dtype=torch.float64).to(dtype)
    gu = torch.randn(o.u.shape, generator=g, dtype=torch.float64).to(dtype)
    k = port.kkt_backward(gx, gu, x0, C, c, F, f, o.x, o.u, ns, nc, **kw)
    ((x * gx.to(dev)).sum() + (u * gu.to(dev)).sum()).backward()
    print("   grads: " + " ".join(
        f"{nm} {rel(a.grad, b):.2e}" for nm, a, b in
        [("dx0", x0g, k.dx_init), ("dC", Cg, k.dC), ("dc", cg, k.dc), ("dF", Fg, k.dF),
         ("df", fg, k.df)]))


def run_env(name, T, B, dtype, lqr_iter, sigma=0.5, seed=0):
    g = torch.Generator().manual_seed(seed)
    if name == "cartpole":
        pdx = port.CartpoleDx(dtype=dtype)
        gdx = env.CartpoleDx(torch.tensor((9.8, 1.0, 0.1, 0.5), dtype=dtype, device=dev))
        r = (torch.rand(B, 4, generator=g, dtype=torch.float64) * 2 - 1) * sigma
        x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1)
    else:
        pdx = port.PendulumDx(dtype=dtype)
        gdx = env.PendulumDx(torch.tensor((10., 1., 1.), dtype=dtype, device=dev))
        th = (torch.rand(B, generator=g, dtype=torch.float64) - 0.5) * 3.14159
        w = torch.rand(B, generator=g, dtype=torch.float64) * 2 - 1
        x0 = torch.stack((torch.cos(th), torch.sin(th), w), 1)
    x0 = x0.to(dtype)
    ns, nc = pdx.n_state, pdx.n_ctrl
    q, p = pdx.get_true_obj()
    C = torch.diag(q)[None, None].repeat(T, B, 1, 1)
    c = p[None, None].repeat(T, B, 1)
    kw = dict(u_lower=pdx.lower, u_upper=pdx.upper, lqr_iter=lqr_iter, eps=pdx.mpc_eps,
              linesearch_decay=pdx.linesearch_decay,
              max_linesearch_iter=pdx.max_linesearch_iter)
    t0 = time.time()
    o = port.mpc_forward(x0, port.QuadCost(C, c), pdx, ns, nc, T, final_pass=False, **kw)
    t_cpu = time.time() - t0
    m = d.MPC(ns, nc, T, verbose=-1, exit_unconverged=False, **kw)
    torch.cuda.synchronize()
    t0 = time.time()
    x, u, costs = m(x0.to(dev), d.QuadCost(C.to(dev), c.to(dev)), gdx)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    info = m.last_info
    print(f"{name} T={T} B={B} {dtype} L={lqr_iter}: iters {info.n_iters}/{o.n_iters} retries {info.retries} "
          f"qp {info.qp_iters[:4]} vs {o.qp_iters[:4]} x {rel(x, o.x):.2e} u {rel(u, o.u):.2e} "
          f"cost {rel(costs, o.costs):.2e} | cpu {t_cpu:.2f}s gpu {t_gpu:.3f}s")


if __name__ == "__main__" and os.environ.get("BASIC", "1") == "1":
    print(torch.cuda.get_device_name(0))
    for dtype in (torch.float64, torch.float32):
        run_lindx(4, 2, 12, 16, dtype, False)
        run_lindx(4, 2, 12, 16, dtype, True)
        run_lindx(5, 1, 20, 40, dtype, True)
        run_lindx(8, 2, 10, 33, dtype, True)
        run_env("pendulum", 20, 64, dtype, 1)
        run_env("pendulum", 20, 64, dtype, 10)
        run_env("cartpole", 20, 64, dtype, 1)
        run_env("cartpole", 50, 128, dtype, 10)
    run_env("cartpole", 50, 65536, torch.float64, 10) if os.environ.get("BIG") else None


def run_dilqr(name, T, B, dtype, lqr_iter, sigma=0.05, seed=0):
    g = torch.Generator().manual_seed(seed)
    if name == "cartpole":
        pdx = port.CartpoleDx(dtype=dtype)
        th0 = torch.tensor((9.8, 1.0, 0.1, 0.5), dtype=dtype)
        gdx = env.CartpoleDx(th0.to(dev).requires_grad_())
        r = (torch.rand(B, 4, generator=g, dtype=torch.float64) * 2 - 1) * sigma
        x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1)
    else:
        pdx = port.PendulumDx(dtype=dtype)
        gdx = env.PendulumDx(torch.tensor((10., 1., 1.), dtype=dtype, device=dev).requires_grad_())
        th = (torch.rand(B, generator=g, dtype=torch.float64) - 0.5) * 3.14159
        w = torch.rand(B, generator=g, dtype=torch.float64) * 2 - 1
        x0 = torch.stack((torch.cos(th), torch.sin(th), w), 1)
    x0 = x0.to(dtype)
    ns, nc = pdx.n_state, pdx.n_ctrl
    q, p = pdx.get_true_obj()
    C = torch.diag(q)[None, None].repeat(T, B, 1, 1)
    c = p[None, None].repeat(T, B, 1)
    kw = dict(u_lower=pdx.lower, u_upper=pdx.upper, lqr_iter=lqr_iter, eps=1e-9,
              linesearch_decay=pdx.linesearch_decay,
              max_linesearch_iter=pdx.max_linesearch_iter)
    o = port.mpc_forward(x0, port.QuadCost(C, c), pdx, ns, nc, T, final_pass=False, **kw)
    gg = torch.Generator().manual_seed(7)
    gx = torch.randn(o.x.shape, generator=gg, dtype=torch.float64).to(dtype)
    gu = torch.randn(o.u.shape, generator=gg, dtype=torch.float64).to(dtype)
    ref = port.dilqr_backward(gx, gu, x0, C, c, o.x, o.u, pdx, ns, nc, pdx.lower, pdx.upper,
                              n_passes=30, tol=1e-15)
    m = d.mpc_explicit.MPC(ns, nc, T, verbose=-1, exit_unconverged=False,
                           detach_unconverged=False, **kw)
    Cg, cg = C.to(dev).requires_grad_(), c.to(dev).requires_grad_()
    x, u, costs = m(x0.to(dev), d.QuadCost(Cg, cg), gdx)
    ((x * gx.to(dev)).sum() + (u * gu.to(dev)).sum()).backward()
    print(f"dilqr {name} T={T} B={B} {dtype}: fwd iters {m.last_info.n_iters}/{o.n_iters} x {rel(x, o.x):.2e} "
          f"passes {m.last_backward} (cpu {ref.n_passes}) dtheta {rel(gdx.params.grad, ref.dtheta.sum(0)):.2e} "
          f"dC {rel(Cg.grad, ref.dC):.2e} dc {rel(cg.grad, ref.dc):.2e}")


if __name__ == "__main__" and os.environ.get("DILQR", "1") == "1":
    for dtype in (torch.float64, torch.float32):
        run_dilqr("pendulum", 20, 8, dtype, 60)
        run_dilqr("cartpole", 12, 8, dtype, 80)
        run_dilqr("cartpole", 30, 40, dtype, 80, sigma=0.3)
