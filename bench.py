#!/usr/bin/env python
"""Headline benchmark: MC fidelity evaluations/s of the RobChar robustness sweep at nspin=7.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one synthetic batch: Philox noise -> perturbed
Hamiltonians -> fidelities [S][C][B] -> 15 statistics (sort-free streaming pass) -> per-group top-k ->
(all launches of a step issued from one C call, rc_robustness_sweep, on device-resident buffers)
clustered/ordinal ranks -> Kendall tau matrices (+ one all-gather of the statistics for N > 1).
Default workload `paper_n7` = BASELINE.json configs[2]: nspin=7 0->6, 19 controller groups x 1000
controllers (the paper's fig-5 sweep size per problem), S=11 sigma_sim levels, B=100 draws.
Controllers are sharded by controller block, per-GPU work fixed (weak scaling).

The JSON line carries: value (device-timed whole-job evals/s, inputs resident), e2e (same metric
through the host-buffer public API: H2D of controllers, D2H of statistics + tau every step),
roofline of the dominant kernel (FP64-pipe bound; algorithmic flops F_alg(N) = 24N^2 + 29N + 3 per
evaluation) against an in-run DFMA peak measurement, and cpu_baseline (the oracle's
per-sample port of the reference path on all host cores).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nspin, in, out, groups, ctrl/group, S, B, zz)
    "paper_n7": (7, 0, 6, 19, 1000, 11, 100, False),        # configs[2] (headline)
    "cfg1_n4": (4, 0, 2, 1, 100, 11, 100000, False),       # configs[0] at throughput B
    "cfg2_n5": (5, 0, 4, 34, 1000, 11, 1000, False),       # configs[1] controller count, reduced B
    "cfg2_n6": (6, 0, 3, 34, 1000, 11, 1000, False),
    "cfg4_n16": (16, 0, 15, 10, 125, 1, 100000, True),     # configs[3] per-GPU slice, reduced B
    "cfg5_n32": (32, 0, 31, 10, 1250, 11, 100, False),     # configs[4] per-GPU slice, reduced B
}


def f_alg(n):
    return 24 * n * n + 29 * n + 3


def synthetic_controllers(C, nspin, seed=20221):
    """biases U(-10,10), time U(1,30), RandomState(20221)."""
    rs = np.random.RandomState(seed)
    ctrl = np.empty((C, nspin + 1))
    ctrl[:, :nspin] = rs.uniform(-10, 10, (C, nspin))
    ctrl[:, nspin] = rs.uniform(1, 30, C)
    return ctrl


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's faithful per-sample port of the reference path, on all host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    nspin, inspin, outspin, sigma, nevals, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle.robchar_oracle import ReferencePathPort, synthetic_controllers as sc
    np.random.seed(seed)
    port = ReferencePathPort(nspin, inspin, outspin, sigma)
    ctrl = sc(8, nspin, seed=seed)
    for k in range(200):
        port.evaluate_noisy_fidelity(ctrl[k % 8], True)
    t0 = time.perf_counter()
    acc = 0.0
    for k in range(nevals):
        acc += port.evaluate_noisy_fidelity(ctrl[k % 8], True)
    return time.perf_counter() - t0, acc


def cpu_reference_throughput(nspin, inspin, outspin, evals_per_worker, procs=None, pool=None):
    procs = procs or os.cpu_count() or 1
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(procs)
    try:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(nspin, inspin, outspin, 0.05, evals_per_worker, 1000 + p) for p in range(procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close(); pool.join()
    busy = max(r[0] for r in res)
    return procs * evals_per_worker / busy, procs, wall


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle's per-sample
    port; the Python reference itself cannot travel to the GPU box), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nspin, inspin, outspin, groups, cg, S, B, zz = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    # bounded sample: the whole --steps/--warmup run stays within ~2 minutes of CPU wall time
    per_worker = max(500, min(args.cpu_evals, int(1.2e6 * (7.0 / nspin) ** 2) // (args.steps + args.warmup)))
    pool = mp.get_context("spawn").Pool(procs)
    try:
        for _ in range(args.warmup):
            cpu_reference_throughput(nspin, inspin, outspin, max(200, per_worker // 10), procs, pool)
        t0 = time.perf_counter()
        vals = []
        for _ in range(args.steps):
            v, procs, _ = cpu_reference_throughput(nspin, inspin, outspin, per_worker, procs, pool)
            vals.append(v)
        wall = time.perf_counter() - t0
    finally:
        pool.close(); pool.join()
    value = float(np.mean(vals))
    sample = f"{procs} procs x {per_worker} evals of evaluate_noisy_fidelity(x, True) per step, nspin={nspin}"
    print(json.dumps({
        "impl": "reference", "metric": "mc_fidelity_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(name, gpus):
    nspin, inspin, outspin, groups, cg, S, B, zz = WORKLOADS[name]
    return {"workload": f"{name}: nspin={nspin} {inspin}->{outspin}, {groups} controller groups x {cg} controllers per GPU, "
                        f"S={S} sigma_sim levels linspace(0,0.1), B={B} draws, complex 3-draw noise model"
                        f"{', ZZ term on' if zz else ''}; fidelities + 15 statistics + top-100 + Kendall tau + ARIM with bootstrap error bars per group",
            "nspin": nspin, "controllers_per_gpu": groups * cg, "sigma_levels": S, "draws": B,
            "evals_per_step": S * groups * cg * B * gpus, "noise": "in-kernel Philox4x32-10 + 1024-layer ziggurat (fp64)",
            "l2": "160 MiB memset (> 126 MB L2) between steps (inside the timed region); per-step fidelity tensor "
                  f"{S * groups * cg * B * 8 / 2**20:.0f} MiB", "parallelism": f"controller-sharded x{gpus}"}


def run_ours(args):
    import torch
    import torch.distributed as td
    import robchar_b200 as rb
    eng = rb.engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    nspin, inspin, outspin, groups, cg, S, B, zz = WORKLOADS[args.workload]
    if args.draws:
        B = args.draws
    C_local = groups * cg
    C_total = C_local * world
    sig_np = np.linspace(0, 0.1, S) if S > 1 else np.array([0.05])
    ctrl_all = synthetic_controllers(C_total, nspin)
    lo, hi = rb.dist.shard_bounds(C_total, world, rank)
    ctrl_np = np.ascontiguousarray(ctrl_all[lo:hi])
    dev = torch.device("cuda", local)
    ctrl = torch.as_tensor(ctrl_np).to(dev)
    sig = torch.as_tensor(sig_np).to(dev)
    eps = float(eng.compute_dkw_error(0.05, B))
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)   # > the 126 MB L2
    fids = torch.empty((S, C_local, B), dtype=torch.float64, device=dev)
    topk = min(100, cg)
    fused = B > 512   # long segments: streaming statistics (no fidelity tensor); short ones: materialise + sort-free statistics pass
    fid_events = []

    # the device-resident step: ONE C call issues every launch (rc_robustness_sweep); the evolution kernel is
    # timed by a pair of events recorded inside that call, on the launching stream, within the timed region
    plan = eng.RobustnessSweepPlan(C_local, S, B, nspin, inspin, outspin, groups=groups, topk=topk, dkw_eps=eps, zz=zz,
                                   fused=fused, fids=None if fused else fids)

    def step(seed, timed=False):
        ev = None
        if timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(); ev[1].record()          # creates the CUDA handles; re-recorded inside the call
            fid_events.append(ev)
        st, tau = plan.run(ctrl, sig, seed=seed, c_offset=lo, evolution_events=ev)
        if world > 1:
            st = rb.dist.all_gather_stats(st, C_total)
        return st, tau

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(w)
        flush.zero_()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.LAUNCHES
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        st, tau = step(1000 + k, timed=True)
        flush.zero_()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches = (eng.LAUNCHES - launches0)
    # the kernel-only time of the dominant (evolution) kernel, events on the launching stream
    fid_ms = float(np.mean([a.elapsed_time(b) for a, b in fid_events])) if fid_events else None
    # ---- e2e: host buffers through the public API, copies inside the timed region --------------------
    ctrl_pinned = torch.as_tensor(ctrl_np).pin_memory()
    sig_host = np.ascontiguousarray(sig_np)
    h2d = ctrl_np.nbytes + sig_host.nbytes
    d2h = 15 * S * C_local * 8 + groups * S * S * 8 + groups * topk * 8 + 2 * groups * S * 8

    def e2e_step(seed):
        return rb.rim_analysis.robustness_sweep(ctrl_pinned.numpy(), sig_host, B, nspin, inspin, outspin, groups=groups,
                                                topk=topk, seed=seed, fused=fused, zz=zz)

    for w in range(max(1, args.warmup // 2)):
        e2e_step(w)
    barrier()
    te0 = time.perf_counter()
    for k in range(args.steps):
        out = e2e_step(2000 + k)
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    clocks = sampler.stop() if rank == 0 else None   # sampled across the device-timed and the end-to-end loops
    e2e_ms = torch.tensor([te * 1e3], dtype=torch.float64, device=dev)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(ms_t, op=td.ReduceOp.MAX)
        td.all_reduce(e2e_ms, op=td.ReduceOp.MAX)
    ms = float(ms_t.item()); e2e_total_ms = float(e2e_ms.item())

    evals_step_total = S * C_total * B
    value = evals_step_total * args.steps / (ms * 1e-3)
    e2e_value = evals_step_total * args.steps / (e2e_total_ms * 1e-3)
    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    fp64_peak = eng.fp64_peak_tflops()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    evals_launch = S * C_local * B
    flops_launch = evals_launch * f_alg(nspin)
    achieved = flops_launch / (fid_ms * 1e-3) / 1e12
    bytes_launch = evals_launch * 8 if not fused else 0
    roofline = {
        "bound": "fp64", "kernel": "fidelity_stats_reg_kernel" if fused else "fidelity_reg_kernel" if nspin <= 16 else "fidelity_smem_kernel",
        "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
        "peak_source": "in-run DFMA-chain microbenchmark (rc_fp64_peak_tflops); MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flops_per_eval": f_alg(nspin), "evals_per_launch": evals_launch, "kernel_ms": fid_ms,
        "kernel_evals_per_sec": evals_launch / (fid_ms * 1e-3),
        "hbm": {"achieved": bytes_launch / (fid_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_launch / (fid_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"},
        "traffic": None,
    }
    try:
        roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass

    cpu_value, procs, cpu_wall = cpu_reference_throughput(nspin, inspin, outspin, args.cpu_evals)
    line = {
        "metric": "mc_fidelity_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, world),
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_total_ms / args.steps,
                "api": "robchar_b200.rim_analysis.robustness_sweep -> rc_robustness_sweep_host (one C call, host buffers in/out)"},
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_value, "unit": "evals/s", "cores": procs, "kind": "port",
                         "sample": f"{procs} procs x {args.cpu_evals} evals of the oracle's per-sample port of "
                                   f"evaluate_noisy_fidelity(x, True), nspin={nspin}, {cpu_wall:.1f} s wall"},
        "sweep_wall_s": {"device": ms / args.steps * 1e-3, "e2e": e2e_total_ms / args.steps * 1e-3},
    }
    if world > 1:
        td.destroy_process_group()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="paper_n7", choices=sorted(WORKLOADS))
    ap.add_argument("--draws", type=int, default=0, help="override B (draws per sigma level and controller)")
    ap.add_argument("--cpu-evals", type=int, default=30000, help="evaluations per CPU worker in the baseline sample")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
