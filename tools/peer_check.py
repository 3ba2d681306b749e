"""Multi-rank check of the peer exchange (run under torchrun on N GPUs): gathered statistics == NCCL all-gather ==
the unsharded sweep, over several pipelined steps; draw-sharded mode == single-GPU statistics."""
import os, sys
import numpy as np, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from bench import synthetic_controllers

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
n, S, B, G, cg = 7, 11, 100, 3, 200
C_total = G * cg * world
ctrl_all = synthetic_controllers(C_total, n)
sig = torch.linspace(0, 0.1, S, dtype=torch.float64).cuda()
eps = float(rb.engine.compute_dkw_error(0.05, B))
sw = rb.dist.ShardedRobustnessSweep(C_total, S, B, n, 0, 6, groups_per_rank=G, topk=100, dkw_eps=eps)
ctrl = torch.as_tensor(np.ascontiguousarray(ctrl_all[sw.lo:sw.hi])).cuda()
ok = True
for k in range(5):
    st, tau = sw.step(ctrl, sig, seed=10 + k)
    ref = rb.dist.all_gather_stats(st.clone(), C_total)
    got = sw.gathered()
    same = torch.equal(got, ref)
    # unsharded single-GPU reference of the same sweep (global Philox counters)
    whole = rb.engine.fidelity_mc_stats(torch.as_tensor(ctrl_all).cuda(), sig, B, n, 0, 6, dkw_eps=eps, seed=10 + k)[1]
    same2 = torch.equal(got, whole)
    ok = ok and same and same2
    if rank == 0:
        print(f"step {k}: peer == nccl {same}, == unsharded {same2}")
# host-buffer path
pinned = torch.as_tensor(np.ascontiguousarray(ctrl_all[sw.lo:sw.hi])).pin_memory().numpy()
for k in range(3):
    sth, tauh, sel, ar, ars = sw.step_host(pinned, sig.cpu().numpy(), seed=50 + k)
    got = sw.gathered()
    whole = rb.engine.fidelity_mc_stats(torch.as_tensor(ctrl_all).cuda(), sig, B, n, 0, 6, dkw_eps=eps, seed=50 + k)[1]
    same = torch.equal(got, whole) and np.array_equal(sth, whole[:, :, sw.lo:sw.hi].cpu().numpy())
    ok = ok and same
    if rank == 0:
        print(f"host step {k}: gathered == unsharded and host block == slice: {same}")
sw.finish()
# draw-sharded: one controller, many draws
if 8 % world == 0:
    c1 = ctrl_all[:1]
    Bd = 200000
    epsd = float(rb.engine.compute_dkw_error(0.05, Bd))
    got = rb.dist.sharded_rim_sweep(c1, sig, Bd, n, 0, 6, dkw_eps=epsd, seed=3)
    whole = rb.engine.fidelity_stats(c1, sig, Bd, n, 0, 6, dkw_eps=epsd, seed=3)
    same = torch.equal(got, whole)
    ok = ok and same
    if rank == 0:
        print("draw-sharded == single GPU:", same, "W row:", got[0, :, 0].tolist()[:3])
# uneven shards through the exchange itself: 7 columns over the ranks, 3 pipelined steps
C7 = 4 * world + 3
x = rb.dist.PeerStatsExchange(5, C7)
for k in range(1, 4):
    blk = x.local_block(k)
    blk.copy_(torch.arange(5, dtype=torch.float64, device="cuda")[:, None] * 1000 + torch.arange(x.lo, x.hi, dtype=torch.float64, device="cuda")[None, :] + k * 0.5)
    x.push(k)
    want = torch.arange(5, dtype=torch.float64, device="cuda")[:, None] * 1000 + torch.arange(C7, dtype=torch.float64, device="cuda")[None, :] + k * 0.5
    same = torch.equal(x.gathered(k), want)
    ok = ok and same
    if rank == 0:
        print(f"uneven exchange step {k}: {same}")
x.raise_if_timed_out()
x.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
td.all_reduce(flag, op=td.ReduceOp.MIN)
sw.close()
td.destroy_process_group()
if rank == 0:
    print("PEER CHECK", "PASSED" if int(flag.item()) else "FAILED")
sys.exit(0 if int(flag.item()) else 1)
