"""cProfile of the host side of a bench step (where the CPU time between launches goes)."""
import cProfile, importlib, os, pstats, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
il = importlib.import_module("differentiable-ilqr_b200.il")
dev = torch.device("cuda:0")
dtype = torch.float64
B = int(os.environ.get("BD_B", "64"))     # tiny batch: the device is never the bottleneck
x0, uexp = [t.to(dev) for t in bench.make_inputs(torch, B, dtype, 0)]
step = il.ImitationStep(env.CartpoleDx, T=50, lqr_iter=10, dtype=dtype, device=dev, n_richardson=4)
q, p = [t.to(dtype).to(dev) for t in env.CartpoleDx().get_true_obj()]
theta = torch.tensor(bench.THETA, dtype=dtype, device=dev)
res = step.prepare(x0, q, p, theta)
for _ in range(5):
    step.run_resident(res, uexp)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step.run_resident(res, uexp)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
