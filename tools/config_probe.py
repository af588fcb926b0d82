"""Timing of the non-headline BASELINE configs (forward solves, CUDA events):
config 3 (rocket T=100 B=16384, box +-20) and a slice of config 5 (synthetic LinDx)."""
import importlib, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
dev = torch.device("cuda:0")

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

def rocket(B=16384, T=100, dtype=torch.float64, bound=20.0):
    g = torch.Generator().manual_seed(0)
    f64 = torch.float64
    qv = torch.cat((torch.ones(B, 1, dtype=f64), 0.1 * torch.randn(B, 3, generator=g, dtype=f64)), 1)
    x0 = torch.cat(((torch.rand(B, 3, generator=g, dtype=f64) * 2 - 1) * 15, torch.rand(B, 3, generator=g, dtype=f64) * 2 - 1,
                    qv / qv.norm(dim=1, keepdim=True), (torch.rand(B, 3, generator=g, dtype=f64) * 2 - 1) * 0.1), 1).to(dtype).to(dev)
    dx = env.RocketDx(torch.tensor((0.5, 1., 1., 1., 1.), dtype=dtype, device=dev))
    q, p = dx.get_true_obj()
    C = torch.diag(q.to(dtype)).to(dev)[None, None].repeat(T, B, 1, 1)
    c = p.to(dtype).to(dev)[None, None].repeat(T, B, 1)
    m = d.MPC(13, 3, T, u_lower=-bound, u_upper=bound, lqr_iter=10, verbose=-1, exit_unconverged=False,
              detach_unconverged=False, eps=dx.mpc_eps, linesearch_decay=dx.linesearch_decay,
              max_linesearch_iter=dx.max_linesearch_iter)
    def run():
        with torch.no_grad():
            return m(x0, d.QuadCost(C, c), dx)
    ms = timed(run)
    lib = importlib.import_module("differentiable-ilqr_b200._lib")
    lib.profile = {}
    x, u, _ = run()
    torch.cuda.synchronize()
    prof = {k: (len(v), round(sum(a.elapsed_time(b) for a, b in v), 3)) for k, v in lib.profile.items()}
    lib.profile = None
    print("  per-call (n, total ms):", prof)
    info = m.last_info
    sat = float((u.abs() >= bound - 1e-9).double().mean())
    print(json.dumps({"config": "rocket T=%d B=%d %s bound=%g" % (T, B, str(dtype)[6:], bound), "ms": ms,
                      "fwd_solves_per_s": B / ms * 1e3, "iters": info.n_iters, "retries": info.retries,
                      "qp_iters": info.qp_iters, "saturated_frac": sat}))

def lindx(ns, nc, T, B, boxed, dtype=torch.float64):
    from common import lindx_problem
    C, c, F, f, x0 = [t.to(dev) for t in lindx_problem(ns, nc, T, B, dtype)]
    kw = dict(u_lower=-1.0, u_upper=1.0) if boxed else {}
    m = d.MPC(ns, nc, T, lqr_iter=10, verbose=-1, exit_unconverged=False, detach_unconverged=False, **kw)
    Cg = C.clone().requires_grad_()
    def run():
        Cg.grad = None
        x, u, _ = m(x0, d.QuadCost(Cg, c), d.LinDx(F, f))
        (x.sum() + u.sum()).backward()
    ms = timed(run)
    info = m.last_info
    print(json.dumps({"config": "lindx ns=%d nc=%d T=%d B=%d boxed=%s %s fwd+KKT bwd" % (ns, nc, T, B, boxed, str(dtype)[6:]),
                      "ms": ms, "solves_per_s": B / ms * 1e3, "iters": info.n_iters, "retries": info.retries}))

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "rocket"):
        rocket(16384, 100, torch.float64, 20.0)
        rocket(16384, 100, torch.float64, 10.0)
        rocket(16384, 100, torch.float32, 20.0)
    if which in ("all", "lindx"):
        for ns, nc in ((4, 1), (4, 2), (8, 2), (8, 4), (16, 4)):
            for boxed in (False, True):
                lindx(ns, nc, 50, 65536 if ns <= 8 else 8192, boxed)
