"""BASELINE config 4 timed: the `il_exp --mode empc` training loop (il.ImitationLearner.step:
warm-started solve, imitation loss, DiLQR backward, NCCL all-reduce of the parameter gradient,
RMSprop update -- theta changes every step) on a global batch of 65536 problems per GPU,
batch-sharded over the ranks.  Prints one JSON line on rank 0."""
import importlib, json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, T = int(os.environ.get("B", 65536)), 50
g = torch.Generator().manual_seed(rank)
r = (torch.rand(B, 4, generator=g, dtype=torch.float64) * 2 - 1) * 0.05
x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1).to(dev)
L = d.il.ImitationLearner(env.CartpoleDx, (9.8, 1.2, 0.1, 0.6), T, lqr_iter=10, device=dev,
                          richardson_passes=5, richardson_tol=None)
L.mpc.verbose = -1
u_exp = L.expert((9.8, 1.0, 0.1, 0.5), x0)
losses = [L.step(x0, u_exp, n_global=B * world) for _ in range(3)]
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
steps = 10
e0.record()
for _ in range(steps):
    losses.append(L.step(x0, u_exp, n_global=B * world))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
if rank == 0:
    print(json.dumps({"what": "ImitationLearner.step (il_exp empc loop), cartpole T=50, B=%d per GPU, "
                              "lqr_iter=10 warm-started, 5 Richardson passes, RMSprop" % B,
                      "n_gpus": world, "ms_per_step": ms, "solves_per_s": B * world / (ms * 1e-3),
                      "loss_first": losses[0], "loss_last": losses[-1],
                      "theta": [round(v, 5) for v in L.theta.tolist()]}))
if world > 1:
    dist.destroy_process_group()
