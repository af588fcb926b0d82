"""Where a bench step spends its time: CUDA-event timing of the pieces of
il.ImitationStep (tile, forward solve, loss, backward) + per C-ABI-call times."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
d = importlib.import_module("differentiable-ilqr_b200")
lib = importlib.import_module("differentiable-ilqr_b200._lib")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
il = importlib.import_module("differentiable-ilqr_b200.il")
dev = torch.device("cuda:0")
dtype = torch.float64
B = int(os.environ.get("BD_B", "65536"))
x0, uexp = [t.to(dev) for t in bench.make_inputs(torch, B, dtype, 0)]
step = il.ImitationStep(env.CartpoleDx, T=50, lqr_iter=10, dtype=dtype, device=dev, n_richardson=4)
q, p = [t.to(dtype).to(dev).requires_grad_() for t in env.CartpoleDx().get_true_obj()]
theta = torch.tensor(bench.THETA, dtype=dtype, device=dev).requires_grad_()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(4):
    for t in (q, p, theta):
        t.grad = None
    e0 = ev()
    C, c = step.tile_cost(q, p, B)
    e1 = ev()
    dx = env.CartpoleDx(theta)
    x, u, _ = step.mpc(x0, d.QuadCost(C, c), dx)
    e2 = ev()
    loss = (u - uexp).pow(2).mean()
    e3 = ev()
    if it == 3:
        lib.profile = {}
    loss.backward()
    e4 = ev()
    torch.cuda.synchronize()
    print("tile %.3f  forward %.3f  loss %.3f  backward %.3f  total %.3f ms" % (
        e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4), e0.elapsed_time(e4)))
prof = {k: [a.elapsed_time(b) for a, b in v] for k, v in lib.profile.items()}
tot = 0
for k, v in prof.items():
    print("  %-28s n=%2d  sum %.3f ms" % (k, len(v), sum(v))); tot += sum(v)
print("  backward C-ABI calls sum %.3f ms" % tot)
