"""il_exp-shaped demo (BASELINE config 4): learn cartpole dynamics parameters by
differentiating through the MPC.  Single GPU, or `torchrun --nproc-per-node N`
(batch-sharded, NCCL all-reduce of the parameter gradient)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = importlib.import_module("differentiable-ilqr_b200")
env = importlib.import_module("differentiable-ilqr_b200.env_dx")
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B_glob, T = int(os.environ.get("B", 4096)), 20
lo, hi = d.parallel.shard_range(B_glob, rank, world)
g = torch.Generator().manual_seed(0)
r = (torch.rand(B_glob, 4, generator=g, dtype=torch.float64) * 2 - 1) * 0.2
x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1)[lo:hi].to(dev)
true_theta, init_theta = (9.8, 1.0, 0.1, 0.5), (9.8, 3.0, 0.1, 1.0)     # il_exp.py:140-142
L = d.il.ImitationLearner(env.CartpoleDx, init_theta, T, lqr_iter=60, device=dev)
u_exp = L.expert(true_theta, x0)
for it in range(int(os.environ.get("STEPS", 40))):
    loss = L.step(x0, u_exp, n_global=B_glob)
    if rank == 0 and it % 5 == 0:
        print("step %3d  im_loss %.4e  theta %s" % (it, loss, [round(v, 4) for v in L.theta.tolist()]))
if rank == 0:
    print("final theta", L.theta.tolist(), "true", true_theta)
if world > 1:
    dist.destroy_process_group()
