#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched differentiable iLQR/MPC path.

Metric (BASELINE.json): MPC solves/sec (forward + implicit backward), cartpole
T=50, B=65536 per GPU, FP64; % of the HBM roofline.

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA)
  python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host CPU

A "step" is one pass of the hot path over one batch of synthetic problems: an
iLQR solve (lqr_iter = 10 fixed, cold start: SURVEY 8d config 2a) followed by
one backward pass.  `value` is measured with every input resident in HBM;
`e2e` is the same metric through the reference-facing API with HOST inputs
(x_init, expert controls, q, p, theta in pinned memory; the cost tensors are
tiled on the device exactly as il_env.IL_Env.mpc does; loss + gradients are
copied back).  Problems shard across GPUs by batch index with no data-path
collective (weak scaling: B per GPU fixed).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_H = 50
NS, NC = 5, 1
N = NS + NC
LQR_ITER = 10
SIGMA = 0.5
THETA = (9.8, 1.0, 0.1, 0.5)


def bytes_per_solve(s, L=LQR_ITER, T=T_H, n=N, ns=NS, nc=NC, ntheta=4):
    """SURVEY 8d ALGORITHMIC bytes: L fused iterations + one backward pass."""
    it = s * (T * n * n + 2 * T * n + T * nc + ns + 2)
    bwd = s * (2 * T * n * n + 4 * T * n + ns + ntheta)
    return it, bwd, L * it + bwd


def ncu_traffic(dtype):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r1_iter_traffic.json")
    if dtype == "f64" and os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], capture_output=True, text=True,
                    timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def make_inputs(torch, B, dtype, seed):
    """x,dx,dth ~ U(-s,s), th ~ U(-s,s) rad, seeded on CPU in fp64."""
    g = torch.Generator().manual_seed(seed)
    r = (torch.rand(B, 4, generator=g, dtype=torch.float64) * 2 - 1) * SIGMA
    x0 = torch.stack((r[:, 0], r[:, 1], torch.cos(r[:, 2]), torch.sin(r[:, 2]), r[:, 3]), 1)
    uexp = torch.randn(T_H, B, NC, generator=g, dtype=torch.float64)
    return x0.to(dtype), uexp.to(dtype)


# ------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    d = importlib.import_module("differentiable-ilqr_b200")
    lib = importlib.import_module("differentiable-ilqr_b200._lib")
    env = importlib.import_module("differentiable-ilqr_b200.env_dx")
    il = importlib.import_module("differentiable-ilqr_b200.il")

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    s = 8 if dtype == torch.float64 else 4
    B = args.batch

    x0_h, uexp_h = make_inputs(torch, B, dtype, seed=rank)
    x0_h, uexp_h = x0_h.pin_memory(), uexp_h.pin_memory()
    theta_h = torch.tensor(THETA, dtype=dtype).pin_memory()
    step = il.ImitationStep(env.CartpoleDx, T=T_H, lqr_iter=LQR_ITER, dtype=dtype, device=dev,
                            n_richardson=args.richardson, tile=not args.broadcast_cost)
    q_h, p_h = [t.to(dtype).pin_memory() for t in env.CartpoleDx().get_true_obj()]

    # resident inputs for the device-timed `value`
    x0 = x0_h.to(dev)
    uexp = uexp_h.to(dev)
    res = step.prepare(x0, q_h.to(dev), p_h.to(dev), theta_h.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    # --- value: inputs resident in HBM ------------------------------------
    lib.profile = None
    n0 = lib.launch_count
    ms = timed(lambda: step.run_resident(res, uexp), args.steps, args.warmup)
    launches = (lib.launch_count - n0) // (args.steps + args.warmup) * args.steps

    # --- per-kernel timing of the dominant kernel (CUDA events, same stream) -
    lib.profile = {}
    barrier()
    for _ in range(max(1, min(args.steps, 3))):
        step.run_resident(res, uexp)
    torch.cuda.synchronize()
    prof = {k: [a.elapsed_time(b) for a, b in v] for k, v in lib.profile.items()}
    lib.profile = None
    it_ms = sorted(prof.get("dilqr_mpc_iterate", [0.0]))
    it_avg = sum(it_ms) / max(1, len(it_ms))

    # --- e2e: host inputs, H2D / D2H inside the timed region ---------------
    def e2e_step():
        out = step.run_host(x0_h, uexp_h, q_h, p_h, theta_h)
        return out
    ms_e2e = timed(e2e_step, max(1, args.steps), max(3, args.warmup))
    h2d = (x0_h.numel() + uexp_h.numel() + q_h.numel() + p_h.numel() + theta_h.numel()) * s
    d2h = step.d2h_bytes

    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank == 0:
        hbm, which = peaks()
        it_b, bwd_b, tot_b = bytes_per_solve(s)
        achieved = it_b * B / (it_avg * 1e-3) / 1e9 if it_avg > 0 else 0.0
        value = world * B / (ms * 1e-3)
        line = {
            "metric": "MPC solves/sec (fwd+implicit bwd), cartpole T=50 B=64k",
            "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": "cartpole env_dx MPC T=50 B=%d/GPU forward(lqr_iter=10, cold start "
                            "sigma=0.5)+%s backward, per B200" % (B, step.backward_name),
                "batch_per_gpu": B, "T": T_H, "lqr_iter": LQR_ITER,
                "l2": "inputs larger than L2 (C alone is %.2f GB)" % (T_H * B * N * N * s / 1e9),
                "parallelism": "batch-sharded x%d, no data-path collective" % world,
                "roofline_frac_of_solve": value / world * tot_b / (hbm * 1e9),
                "bytes_per_solve": tot_b, "richardson_passes": args.richardson,
                "cost_layout": "broadcast [n,n]" if args.broadcast_cost else "dense [T,B,n,n]",
                "trace_retries": step.retries,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm,
                         "traffic": None if args.broadcast_cost else ncu_traffic(args.dtype),
                         "peak_source": which,
                         "kernel": "ilqr_iter_kernel<double,5,1,CARTPOLE>",
                         "algorithmic_bytes_per_launch": it_b * B,
                         "avg_launch_ms": it_avg},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "solves/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": sampler.summary() if sampler else None,
            "kernel_ms": {k: sum(v) / len(v) for k, v in prof.items()},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(dtype, sample_B=args.cpu_batch,
                                                n_passes=args.richardson)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------- reference arm
def cpu_solve_once(torch, port, dtype, B, seed=0, n_passes=4):
    """One step of the same workload with the oracle port (== the reference's
    algorithm, validated bit-exact against it) on the host CPU."""
    x0, uexp = make_inputs(torch, B, dtype, seed)
    dx = port.CartpoleDx(dtype=dtype)
    q, p = dx.get_true_obj()
    C = torch.diag(q)[None, None].repeat(T_H, B, 1, 1)
    c = p[None, None].repeat(T_H, B, 1)
    t0 = time.time()
    o = port.mpc_forward(x0, port.QuadCost(C, c), dx, NS, NC, T_H, u_lower=dx.lower,
                         u_upper=dx.upper, lqr_iter=LQR_ITER, eps=dx.mpc_eps,
                         linesearch_decay=dx.linesearch_decay,
                         max_linesearch_iter=dx.max_linesearch_iter, final_pass=True)
    gu = 2.0 * (o.u - uexp) / o.u.numel()
    gx = torch.zeros_like(o.x)
    port.dilqr_backward(gx, gu, x0, C, c, o.x, o.u, dx, NS, NC, dx.lower, dx.upper,
                        n_passes=n_passes)
    return time.time() - t0


def cpu_baseline(dtype, sample_B=2048, n_passes=4):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.no_grad():
        dt = cpu_solve_once(torch, port, dtype, sample_B, n_passes=n_passes)
    return {"value": sample_B / dt, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": "B=%d problems of the same workload (forward lqr_iter=10 + DiLQR backward, "
                      "%d Richardson passes) with the oracle port of the reference, %.1f s"
                      % (sample_B, n_passes, dt)}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import port
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = args.cpu_batch
    with torch.no_grad():
        for _ in range(min(args.warmup, 1)):
            cpu_solve_once(torch, port, dtype, min(Bs, 64), n_passes=args.richardson)
        t = [cpu_solve_once(torch, port, dtype, Bs, seed=i, n_passes=args.richardson)
             for i in range(max(1, min(args.steps, 3)))]
    ms = 1e3 * sum(t) / len(t)
    v = Bs / (ms * 1e-3)
    sample = ("B=%d problems per step on %d host threads, oracle port of the reference "
              "(its matrix-free backward; the reference's own dense fix_point_equ is limited "
              "to T*B < 1000)" % (Bs, cores))
    print(json.dumps({
        "impl": "reference",
        "metric": "MPC solves/sec (fwd+implicit bwd), cartpole T=50 B=64k",
        "value": v, "unit": "solves/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": len(t), "warmup": min(args.warmup, 1), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": "cartpole env_dx MPC T=50 forward(lqr_iter=10, cold start sigma=0.5)"
                               "+DiLQR implicit (%d Richardson passes) backward, oracle port of "
                               "the reference on host CPU" % args.richardson, "batch": Bs},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--cpu-batch", type=int, default=32768)
    ap.add_argument("--richardson", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--broadcast-cost", action="store_true",
                    help="hand MPC the [n,n]/[n] cost (mpc.py:205-219 broadcast) instead of the "
                         "dense [T,B,n,n] tiling of il_env.py:159-162 (not the headline config)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
