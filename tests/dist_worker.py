"""Worker of tests/test_dist_gloo.py::test_two_rank_step_equals_full_batch (and of the
`--gpus 2` NCCL variant): every rank runs the il_exp-shaped step on its shard of ONE global
batch, the gradient scalars are all-reduced, and rank 0 also runs the full batch alone.

    python -m torch.distributed.run --nproc-per-node 2 tests/dist_worker.py OUT.pt [gloo|nccl]

With `gloo` both ranks may share one GPU (the single-GPU box of the round-end test run);
with `nccl` each rank owns its device."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out, backend = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "gloo")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", 0)) if backend == "nccl" else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    d = importlib.import_module("differentiable-ilqr_b200")
    env = importlib.import_module("differentiable-ilqr_b200.env_dx")
    il = d.il
    T, B = 30, 192
    g = torch.Generator().manual_seed(5)
    r = (torch.rand(B, 4, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
    x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1)
    uexp = torch.randn(T, B, 1, generator=g, dtype=torch.float64)
    q, p = [t.double() for t in env.CartpoleDx().get_true_obj()]
    theta = torch.tensor((9.8, 1.2, 0.1, 0.6), dtype=torch.float64)
    lo, hi = d.parallel.shard_range(B, rank, world)

    def run(lo, hi, frac, group):
        step = il.ImitationStep(env.CartpoleDx, T=T, lqr_iter=40, dtype=torch.float64, device=dev,
                                n_richardson=8, group=group)
        step.mpc.solo = True      # per-problem pnqp flags: see the note in the test
        res = step.prepare(x0[lo:hi].to(dev), q.to(dev), p.to(dev), theta.to(dev))
        flat, dfr = step._step(res["x0"], uexp[:, lo:hi].contiguous().to(dev), res["q"], res["p"],
                               res["theta"], frac, res["theta_host"], False)
        return flat.cpu()

    # sharded: local mean * (local share) summed over ranks == global mean
    sharded = run(lo, hi, (hi - lo) / B, None)
    result = {"sharded": sharded, "world": world}
    # ImitationLearner.step on the shard (RMSprop update from the all-reduced gradient)
    L = il.ImitationLearner(env.CartpoleDx, (9.8, 3.0, 0.1, 1.0), T, lqr_iter=30, device=dev,
                            richardson_passes=12, richardson_tol=None)
    L.mpc.solo = True
    loss = L.step(x0[lo:hi].to(dev), uexp[:, lo:hi].contiguous().to(dev), n_global=B)
    result.update(learner_loss=loss, learner_theta=L.theta.detach().cpu())
    # the same step on the whole batch, alone: a private single-rank group (new_group is
    # collective), over which parallel.allreduce_sum_ has nothing to exchange
    solo_group = dist.new_group([0])
    if rank == 0:
        result["full"] = run(0, B, 1.0, solo_group)
        L1 = il.ImitationLearner(env.CartpoleDx, (9.8, 3.0, 0.1, 1.0), T, lqr_iter=30, device=dev,
                                 richardson_passes=12, richardson_tol=None, group=solo_group)
        L1.mpc.solo = True
        result["full_loss"] = L1.step(x0.to(dev), uexp.to(dev), n_global=B)
        result["full_theta"] = L1.theta.detach().cpu()
        torch.save(result, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
