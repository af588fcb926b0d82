"""Per-C-ABI-call CUDA-event times of one synthetic LinDx solve + KKT backward (config 5)."""
import importlib, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = importlib.import_module("differentiable-ilqr_b200")
_lib = importlib.import_module("differentiable-ilqr_b200._lib")
ns, nc, T, B = [int(v) for v in sys.argv[1:5]]
boxed = len(sys.argv) > 5 and sys.argv[5] == "boxed"
dev = torch.device("cuda:0"); dtype = torch.float64
gd = torch.Generator(device=dev).manual_seed(0)
n = ns + nc
rn = lambda *sh: torch.randn(*sh, generator=gd, dtype=dtype, device=dev)
A = rn(T, B, n, n)
Cc = (A.transpose(2, 3) @ A + torch.eye(n, dtype=dtype, device=dev)).contiguous(); del A
cc = rn(T, B, n)
F = torch.cat((torch.eye(ns, dtype=dtype, device=dev).expand(T - 1, B, ns, ns) + 0.2 * rn(T - 1, B, ns, ns) / ns ** 0.5,
               rn(T - 1, B, ns, nc) / ns ** 0.5), 3).contiguous()
f = 0.1 * rn(T - 1, B, ns); x0 = rn(B, ns)
kw = dict(u_lower=-1.0, u_upper=1.0) if boxed else {}
m = d.MPC(ns, nc, T, lqr_iter=10, verbose=-1, exit_unconverged=False, detach_unconverged=False, n_batch=B, **kw)
Cg, cg = Cc.requires_grad_(), cc.requires_grad_()
def fn():
    Cg.grad = cg.grad = None
    x, u, _ = m(x0, d.QuadCost(Cg, cg), d.LinDx(F, f))
    (x.sum() + u.sum()).backward()
fn(); torch.cuda.synchronize()
_lib.profile = {}
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
out = {k: (len(v), round(sum(a.elapsed_time(b) for a, b in v), 3)) for k, v in _lib.profile.items()}
print(json.dumps({"shape": [ns, nc, T, B, boxed], "total_ms": round(e0.elapsed_time(e1), 3), "iters": m.last_info.n_iters, "calls": out}))
