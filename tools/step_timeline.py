"""GPU timeline of one bench step (torch.profiler / CUPTI): kernels in launch order with the
idle gap in front of each, to find where the device waits for the host."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
il = importlib.import_module("differentiable-ilqr_b200.il")
dev = torch.device("cuda:0")
dtype = torch.float64
B = int(os.environ.get("BD_B", "65536"))
x0, uexp = [t.to(dev) for t in bench.make_inputs(torch, B, dtype, 0)]
step = il.ImitationStep(env.CartpoleDx, T=50, lqr_iter=10, dtype=dtype, device=dev, n_richardson=4)
q, p = [t.to(dtype).to(dev) for t in env.CartpoleDx().get_true_obj()]
theta = torch.tensor(bench.THETA, dtype=dtype, device=dev)
res = step.prepare(x0, q, p, theta)
for _ in range(3):
    step.run_resident(res, uexp)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step.run_resident(res, uexp)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
busy = 0.0
print("%9s %8s %8s  %s" % ("start_us", "dur_us", "gap_us", "kernel"))
for e in ev:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = s - prev_end
    busy += d
    if gap > 15 or d > 150:
        print("%9.0f %8.0f %8.0f  %s" % (s - t0, d, gap, e.name[:70]))
    prev_end = max(prev_end, e.time_range.end)
print("span %.0f us, busy %.0f us, idle %.0f us, kernels %d" % (prev_end - t0, busy, prev_end - t0 - busy, len(ev)))
