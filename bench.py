#!/usr/bin/env python
"""Headline benchmark: MC fidelity evaluations/s of the RobChar robustness sweep at nspin=7.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one synthetic batch: Philox noise -> perturbed Hamiltonians ->
fidelities [S][C][B] -> 15 statistics -> per-group top-k -> clustered / ordinal ranks -> Kendall tau matrices ->
ARIM with bootstrap error bars (every launch of a step issued from one C call, rc_robustness_sweep, on
device-resident buffers).  For N > 1 the controllers are sharded by block, one process per GPU, and every rank
pushes its [15][S][C_local] statistics block into every peer's [15][S][C_total] tensor over NVLink with the copy
engines on a side stream (dist.PeerStatsExchange, csrc/rc_peer.cu): step k's exchange runs underneath step k+1's
evolution; the timed region ends after the last exchange has landed on every rank.

Default workload `paper_n7` = BASELINE.json configs[2]: nspin=7 0->6, 19 controller groups x 1000 controllers per
GPU (the paper's fig-5 sweep size per problem), S=11 sigma_sim levels, B=100 draws; per-GPU work fixed (weak
scaling).  `--workload` selects the other BASELINE configurations (per-GPU slices, or `*_full` = the full size
quoted in BASELINE.json: those are strong-scaling workloads whose total size is fixed).

The JSON line carries: value (device-timed whole-job evals/s, inputs resident), per-step device times (min / median
/ max per rank), e2e (the same metric through the host-buffer public API incl. the exchange: H2D of controllers, D2H
of statistics + tau + selection + ARIM every step), roofline of the dominant kernel (FP64-pipe bound; algorithmic
flops F_alg(N) = 24N^2 + 29N + 3 per evaluation) against an in-run DFMA peak measurement, cpu_baseline
(the UNMODIFIED reference's evaluate_noisy_fidelity on all host cores, staged under oracle/_ref by build(); the
oracle's port when that is absent), and at N=1 the wall time of the user-facing MCDataSim.get_metrics_dict for
both arms at BASELINE configs[0].
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: nspin, in, out, groups, controllers per group, S, B, zz, scaling ("weak": counts are per GPU; "strong": total)
WORKLOADS = {
    "paper_n7": dict(n=7, i=0, o=6, G=19, cg=1000, S=11, B=100, zz=False, scaling="weak"),            # configs[2] (headline)
    "cfg1_n4": dict(n=4, i=0, o=2, G=1, cg=100, S=11, B=100000, zz=False, scaling="weak"),             # configs[0], throughput B
    "cfg1_full_n4": dict(n=4, i=0, o=2, G=1, cg=100, S=11, B=1000000, zz=False, scaling="weak"),       # configs[0] at B = 1e6
    "cfg2_n5": dict(n=5, i=0, o=4, G=34, cg=1000, S=11, B=1000, zz=False, scaling="weak"),             # configs[1] controllers, reduced B
    "cfg2_n6": dict(n=6, i=0, o=3, G=34, cg=1000, S=11, B=1000, zz=False, scaling="weak"),
    "cfg2_full_n5": dict(n=5, i=0, o=4, G=34, cg=1000, S=11, B=1000000, zz=False, scaling="strong"),   # configs[1]: 34 000 x 11 x 1e6
    "cfg2_full_n6": dict(n=6, i=0, o=3, G=34, cg=1000, S=11, B=1000000, zz=False, scaling="strong"),
    "cfg4_n16": dict(n=16, i=0, o=15, G=10, cg=125, S=1, B=100000, zz=True, scaling="weak"),           # configs[3] per-GPU slice, reduced B
    "cfg4_full": dict(n=16, i=0, o=15, G=80, cg=125, S=1, B=1000000, zz=True, scaling="strong"),       # configs[3]: 1e4 x 1e6, Z term
    "cfg5_n32": dict(n=32, i=0, o=31, G=10, cg=1250, S=11, B=100, zz=False, scaling="weak"),           # configs[4] per-GPU slice, reduced B
    "cfg5_full": dict(n=32, i=0, o=31, G=80, cg=1250, S=11, B=100000, zz=False, scaling="strong"),     # configs[4]: 1e5 x 1e5 x 11
    "n4_paper": dict(n=4, i=0, o=2, G=19, cg=1000, S=11, B=100, zz=False, scaling="weak"),             # paper sweep shape at the other chain lengths
    "n5_paper": dict(n=5, i=0, o=4, G=19, cg=1000, S=11, B=100, zz=False, scaling="weak"),
    "n6_paper": dict(n=6, i=0, o=3, G=19, cg=1000, S=11, B=100, zz=False, scaling="weak"),
    "n16_paper": dict(n=16, i=0, o=15, G=10, cg=500, S=11, B=100, zz=True, scaling="weak"),
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


def shard_shape(w, world):
    """(groups on this rank, controllers per rank, total controllers) of a workload at `world` GPUs."""
    if w["scaling"] == "weak":
        return w["G"], w["G"] * w["cg"], w["G"] * w["cg"] * world
    if w["G"] % world:
        raise SystemExit(f"workload needs a GPU count that divides {w['G']} groups")
    return w["G"] // world, w["G"] // world * w["cg"], w["G"] * w["cg"]


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref, staged by build()), else the oracle's per-sample port
# ------------------------------------------------------------------------------------------------
def _port_worker(args):
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


def cpu_arm_kind():
    from oracle import ref_runner
    return "reference" if ref_runner.available() else "port"


def cpu_reference_throughput(nspin, inspin, outspin, evals_per_worker, procs=None, pool=None):
    from oracle import ref_runner
    worker = ref_runner.eval_worker if ref_runner.available() else _port_worker
    procs = procs or os.cpu_count() or 1
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(procs)
    try:
        t0 = time.perf_counter()
        res = pool.map(worker, [(nspin, inspin, outspin, 0.05, evals_per_worker, 1000 + p) for p in range(procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close(); pool.join()
    busy = max(r[0] for r in res)
    return procs * evals_per_worker / busy, procs, wall


def cpu_sample_text(kind, procs, evals, nspin):
    what = ("the unmodified reference's structured_perturbation.evaluate_noisy_fidelity(x, True) (noise_model.py:98-147, "
            "oracle/_ref)") if kind == "reference" else "the oracle's per-sample port of evaluate_noisy_fidelity(x, True)"
    return f"{procs} procs x {evals} evals of {what}, nspin={nspin}, sigma=0.05, one Python process per core"


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


def workload_config(name, gpus):
    w = WORKLOADS[name]
    G, C_local, C_total = shard_shape(w, gpus)
    per = "per GPU" if w["scaling"] == "weak" else f"in total ({G} groups per GPU)"
    return {"workload": f"{name}: nspin={w['n']} {w['i']}->{w['o']}, {w['G']} controller groups x {w['cg']} controllers {per}, "
                        f"S={w['S']} sigma_sim levels {'linspace(0,0.1)' if w['S'] > 1 else '(0.05)'}, B={w['B']} draws, complex 3-draw noise model"
                        f"{', ZZ term on' if w['zz'] else ''}; fidelities + 15 statistics + top-100 + Kendall tau + ARIM with bootstrap error bars per group",
            "nspin": w["n"], "controllers_per_gpu": C_local, "controllers_total": C_total, "sigma_levels": w["S"], "draws": w["B"],
            "evals_per_step": w["S"] * C_total * w["B"], "noise": "in-kernel Philox4x32-10 + 1024-layer ziggurat (fp64)",
            "l2": "160 MiB memset (> 126 MB L2) between steps (inside the timed region)"
                  + (f"; per-step fidelity tensor {w['S'] * C_local * w['B'] * 8 / 2**20:.0f} MiB" if w["B"] <= 512 else "; fused streaming statistics (no fidelity tensor)"),
            "parallelism": f"controller-sharded x{gpus}" + ("" if gpus == 1 else ", statistics blocks pushed peer-to-peer over NVLink (copy engines, side stream)")}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores — the unmodified
    noise_model.py staged in oracle/_ref (kind "reference"), or the oracle's port when it is absent (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    nspin, inspin, outspin = w["n"], w["i"], w["o"]
    procs = os.cpu_count() or 1
    kind = cpu_arm_kind()
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
    print(json.dumps({
        "impl": "reference", "metric": "mc_fidelity_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": procs, "kind": kind,
                         "sample": cpu_sample_text(kind, procs, per_worker, nspin) + " per step"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def mcdatasim_walls(rb, paper_group=False):
    """Wall time of the user-facing cached API, MCDataSim.get_metrics_dict from scratch (mcsim.py:463-510), for both
    arms: BASELINE configs[0] (N=4 0->2, the 100 LBFGS controllers, S=11, B=100) and, with --paper-group, one
    paper-size group (11 x 1000 x 100 at N=7)."""
    from oracle import ref_runner
    out = {}
    cases = [("config1_n4_100x11x100", 4, 0, 2, None, 100)]
    if paper_group:
        cases.append(("paper_group_n7_1000x11x100", 7, 0, 6, 1000, 100))
    noises = np.linspace(0, 0.1, 11)
    for name, n, i, o, C, B in cases:
        if C is None and ref_runner.available():
            conts = ref_runner.lbfgs_controllers(n, i, o)
        else:
            conts = synthetic_controllers(C or 100, n).tolist()
        C = len(conts)
        rec = {"controllers": C, "sigma_levels": 11, "draws": B}
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(f"{td}/experiments/bench")
            json.dump({"lbfgs": {str(n): {"controller": [list(map(float, c)) for c in conts]}}},
                      open(f"{td}/experiments/bench/ppo_spin_{n}_{i}-{o}_c_{C}", "w"))
            os.chdir(td)
            try:
                for rep in range(2):   # second run: steady state (library loaded, allocator warm); caches removed in between
                    for fn in os.listdir(f"{td}/experiments/bench"):
                        if fn.endswith(".mc") or fn.endswith(".mcm"):
                            os.remove(f"{td}/experiments/bench/{fn}")
                    sim = rb.MCDataSim(experiment_name="bench", Nspin=n, inspin=i, outspin=o, noises=noises, bootreps=B,
                                       numcontrollers=C, topk=min(100, C), seed=rep)
                    t0 = time.perf_counter()
                    md = sim.get_metrics_dict(None, noises, algoname="lbfgs")
                    rec["ours_s" if rep else "ours_first_call_s"] = time.perf_counter() - t0
                assert len(md["lbfgs"]) == 15
            finally:
                os.chdir(cwd)
        if ref_runner.available():
            devnull = open(os.devnull, "w")
            so, se = sys.stdout, sys.stderr
            sys.stdout = sys.stderr = devnull          # the reference prints per sigma level and a tqdm bar
            try:
                rec["reference_s"], _ = ref_runner.mcdatasim_wall(n, i, o, conts, noises, B)
            finally:
                sys.stdout, sys.stderr = so, se
            rec["speedup"] = rec["reference_s"] / rec["ours_s"]
        out[name] = rec
    if ref_runner.available():
        out["reference_stage_us"] = ref_runner.stage_timings()
    return out


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
    w = dict(WORKLOADS[args.workload])
    if args.draws:
        w["B"] = args.draws
    nspin, inspin, outspin, S, B, zz = w["n"], w["i"], w["o"], w["S"], w["B"], w["zz"]
    groups, C_local, C_total = shard_shape(w, world)
    cg = w["cg"]
    sig_np = np.linspace(0, 0.1, S) if S > 1 else np.array([0.05])
    ctrl_all = synthetic_controllers(C_total, nspin)
    lo, hi = rb.dist.shard_bounds(C_total, world, rank)
    assert hi - lo == C_local
    ctrl_np = np.ascontiguousarray(ctrl_all[lo:hi])
    dev = torch.device("cuda", local)
    ctrl = torch.as_tensor(ctrl_np).to(dev)
    sig = torch.as_tensor(sig_np).to(dev)
    eps = float(eng.compute_dkw_error(0.05, B))
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)   # > the 126 MB L2
    topk = min(100, cg)
    fused = B > 512   # long segments: streaming statistics (no fidelity tensor); short ones: materialise + sort-free statistics pass
    sweep = rb.dist.ShardedRobustnessSweep(C_total, S, B, nspin, inspin, outspin, groups_per_rank=groups, topk=topk, dkw_eps=eps,
                                           zz=zz, fused=fused, exchange=args.exchange)
    # full-size workloads: warm up on a short-B twin of the sweep (same kernels, same shapes otherwise)
    big = S * C_local * B > 2e9
    warm = sweep if not big else rb.dist.ShardedRobustnessSweep(C_total, S, max(1024, B // 1000), nspin, inspin, outspin,
                                                              groups_per_rank=groups, topk=topk, dkw_eps=eps, zz=zz, fused=True,
                                                              exchange="nccl" if world > 1 else "peer")
    fid_events = []

    def step(seed, timed=False, obj=sweep):
        ev = None
        if timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(); ev[1].record()          # creates the CUDA handles; re-recorded inside the C call
            fid_events.append(ev)
        return obj.step(ctrl, sig, seed=seed, evolution_events=ev)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    t_warm = time.perf_counter()
    for k in range(args.warmup):
        step(k, obj=warm)
        flush.zero_()
    warm.finish()
    warm_ms = (time.perf_counter() - t_warm) * 1e3 / max(1, args.warmup)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # keep the GPU busy (untimed steps) while the clock sampler spins up: an idle gap here lets the SM clock drop and the
    # first timed step pays the ramp (seen as one 7.0 ms step among 5.3 ms ones)
    n_spin = torch.tensor([min(200, max(2, int(300.0 / max(warm_ms, 0.05)) + 1))], device=dev)
    if world > 1:
        td.broadcast(n_spin, 0)        # every rank runs the same number of steps (the exchange is step-synchronous)
    for k_spin in range(int(n_spin.item())):
        step(500 + k_spin, obj=warm)
        flush.zero_()
        if k_spin % 8 == 7:
            torch.cuda.synchronize()
    warm.finish()
    launches0 = eng.launch_count()
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_end = torch.cuda.Event(enable_timing=True)
    for k in range(args.steps):
        marks[k].record()
        step(1000 + k, timed=True)
        flush.zero_()
    marks[args.steps].record()
    sweep.gathered()                         # the last exchange has landed on this rank: end of the timed region
    t_end.record()
    barrier()
    sweep.finish()                           # raises on non-convergence / illegal samples / exchange timeout
    ms = marks[0].elapsed_time(t_end)
    step_ms = [marks[k].elapsed_time(marks[k + 1]) for k in range(args.steps)]
    tail_ms = marks[args.steps].elapsed_time(t_end)
    launches = eng.launch_count() - launches0
    fid_ms = float(np.mean([a.elapsed_time(b) for a, b in fid_events])) if fid_events else None

    # ---- e2e: host buffers through the public API, copies and the exchange inside the timed region ------------
    ctrl_pinned = torch.as_tensor(ctrl_np).pin_memory().numpy()
    sig_host = np.ascontiguousarray(sig_np)
    h2d = ctrl_np.nbytes + sig_host.nbytes
    d2h = 15 * S * C_local * 8 + groups * S * S * 8 + groups * topk * 8 + 2 * groups * S * 8
    e2e_steps = args.steps if not big else max(1, args.steps // 2)

    def e2e_step(seed, obj=sweep):
        if world == 1 and not args.e2e_sharded_api:
            return rb.rim_analysis.robustness_sweep(ctrl_pinned, sig_host, B, nspin, inspin, outspin, groups=groups, topk=topk,
                                                    seed=seed, fused=fused, zz=zz, copy=False)
        return obj.step_host(ctrl_pinned, sig_host, seed=seed)

    if args.skip_e2e:
        e2e_steps = 0
    if not big and e2e_steps:
        for k in range(max(1, args.warmup // 2)):
            e2e_step(k)
    sweep.finish()
    barrier()
    e2e_step_ms = []
    te0 = time.perf_counter()
    for k in range(e2e_steps):
        ts = time.perf_counter()
        out = e2e_step(2000 + k)
        e2e_step_ms.append((time.perf_counter() - ts) * 1e3)
    sweep.finish()                           # waits for the last exchange
    te = time.perf_counter() - te0
    clocks = sampler.stop() if rank == 0 else None   # sampled across the device-timed and the end-to-end loops
    red = torch.tensor([ms, te * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(red, op=td.ReduceOp.MAX)
    ms, e2e_total_ms = float(red[0].item()), float(red[1].item())
    per_rank = [None] * world
    mine = {"min": float(np.min(step_ms)), "median": float(np.median(step_ms)), "max": float(np.max(step_ms)),
            "exchange_tail": tail_ms, "evolution_kernel": fid_ms}
    if world > 1:
        td.all_gather_object(per_rank, mine)
    else:
        per_rank[0] = mine

    evals_step_total = S * C_total * B
    value = evals_step_total * args.steps / (ms * 1e-3)
    e2e_value = evals_step_total * e2e_steps / (e2e_total_ms * 1e-3) if e2e_steps else None
    if rank != 0:
        sweep.close()
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
        "bound": "fp64", "kernel": eng.evolution_kernel_name(nspin, replay=False, fused=fused) + (" + fused_finalize_kernel" if fused else ""),
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

    kind = cpu_arm_kind()
    cpu_value, procs, cpu_wall = cpu_reference_throughput(nspin, inspin, outspin, args.cpu_evals)
    line = {
        "metric": "mc_fidelity_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, world),
        "clocks": clocks, "gpu_launches": launches,
        "per_step_ms": {"min": min(p["min"] for p in per_rank), "median": float(np.median([p["median"] for p in per_rank])),
                        "max": max(p["max"] for p in per_rank), "per_rank": per_rank,
                        "note": "CUDA events at step boundaries on each rank's compute stream; exchange_tail = wait for the "
                                "last step's peer blocks after the loop (the exchange of the other steps is overlapped)"},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_total_ms / e2e_steps if e2e_steps else None, "steps": e2e_steps,
                "step_ms_min_median_max": [float(np.min(e2e_step_ms)), float(np.median(e2e_step_ms)), float(np.max(e2e_step_ms))] if e2e_steps else None,
                "api": ("robchar_b200.rim_analysis.robustness_sweep -> rc_robustness_sweep_host (one C call, host buffers in/out)"
                        if world == 1 and not args.e2e_sharded_api else
                        "robchar_b200.dist.ShardedRobustnessSweep.step_host -> rc_robustness_sweep_host_keep (one C call per rank, "
                        "host buffers in/out) + peer push of the statistics blocks; the timed region ends when the last exchange has landed")},
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_value, "unit": "evals/s", "cores": procs, "kind": kind,
                         "sample": cpu_sample_text(kind, procs, args.cpu_evals, nspin) + f", {cpu_wall:.1f} s wall"},
        "sweep_wall_s": {"device": ms / args.steps * 1e-3, "e2e": e2e_total_ms / e2e_steps * 1e-3 if e2e_steps else None},
        "spectral_fallbacks": eng.spectral_fallbacks() if nspin >= 11 else None,
    }
    if world == 1 and args.workload == "paper_n7" and not args.no_mcdatasim:
        try:
            line["mcdatasim_get_metrics_dict"] = mcdatasim_walls(rb, paper_group=args.paper_group)
        except Exception as e:
            line["mcdatasim_get_metrics_dict"] = {"error": repr(e)}
    sweep.close()
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
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU exchange of the statistics blocks")
    ap.add_argument("--e2e-sharded-api", action="store_true", help="time ShardedRobustnessSweep.step_host at N=1 as well")
    ap.add_argument("--skip-e2e", action="store_true", help="table cells of the full-size workloads: skip the second (host-buffer) pass; e2e.value is null")
    ap.add_argument("--no-mcdatasim", action="store_true", help="skip the MCDataSim.get_metrics_dict wall-time leg")
    ap.add_argument("--paper-group", action="store_true", help="also time one paper-size group (11x1000x100) through MCDataSim (~80 s of CPU)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
