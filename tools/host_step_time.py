"""Host cost of one bench step: wall time per step at a batch so small that the device never
limits (the step is then purely enqueue + the one sync), next to the headline batch."""
import importlib, os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
il = importlib.import_module("differentiable-ilqr_b200.il")
dev = torch.device("cuda:0"); dtype = torch.float64
out = {}
for B in (64, 65536):
    x0, uexp = [t.to(dev) for t in bench.make_inputs(torch, B, dtype, 0)]
    step = il.ImitationStep(env.CartpoleDx, T=50, lqr_iter=10, dtype=dtype, device=dev, n_richardson=4)
    q, p = [t.to(dtype).to(dev) for t in env.CartpoleDx().get_true_obj()]
    res = step.prepare(x0, q, p, torch.tensor(bench.THETA, dtype=dtype, device=dev))
    for _ in range(5):
        step.run_resident(res, uexp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        step.run_resident(res, uexp)
    torch.cuda.synchronize()
    out["B=%d" % B] = round(1e3 * (time.perf_counter() - t0) / n, 3)
print(json.dumps({"wall_ms_per_step": out, "cpus": os.cpu_count()}))
