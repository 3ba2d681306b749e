"""World-size-2 gloo test of the sharded sweep's host logic (controller blocks, uneven shards,
all-gather assembly).  The per-block compute is supplied by the CPU oracle here; on the GPU box the
same code path runs the CUDA sweep under NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as tmp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, C, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    import robchar_b200 as rb
    from oracle import robchar_oracle as orc
    n, S, B = 4, 3, 5
    ctrl = orc.synthetic_controllers(C, n)
    sig = np.linspace(0, 0.1, S)

    def compute(block, c_offset):
        # deterministic "noise" that depends on the GLOBAL controller index, like the Philox counters
        rs_all = np.random.RandomState(7).standard_normal((S, C, B, 3 * n))
        z = rs_all[:, c_offset:c_offset + block.shape[0]]
        f = orc.fidelity_mc_replay(block, sig, z, n, 0, 2)
        m = orc.metrics(f, 0.05)
        return torch.as_tensor(np.stack([m[k] for k in rb.engine.STAT_KEYS]))

    st = rb.dist.sharded_rim_sweep(ctrl, sig, B, n, 0, 2, compute_fn=compute)
    lo, hi = rb.dist.shard_bounds(C, world, rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), st.numpy())
    np.save(os.path.join(out_dir, f"b{rank}.npy"), np.array([lo, hi]))
    td.destroy_process_group()


@pytest.mark.parametrize("C", [7, 8])
def test_sharded_sweep_world2(tmp_path, C):
    port = _free_port()
    tmp.spawn(_worker, args=(2, port, C, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "r0.npy"); b = np.load(tmp_path / "r1.npy")
    assert a.shape == (15, 3, C) and np.array_equal(a, b, equal_nan=True)   # every rank holds the full tensor
    b0 = np.load(tmp_path / "b0.npy"); b1 = np.load(tmp_path / "b1.npy")
    assert b0[0] == 0 and b0[1] == b1[0] and b1[1] == C
    # identical to the unsharded computation
    sys.path.insert(0, ROOT)
    from oracle import robchar_oracle as orc
    import robchar_b200 as rb
    n, S, B = 4, 3, 5
    ctrl = orc.synthetic_controllers(C, n)
    z = np.random.RandomState(7).standard_normal((S, C, B, 3 * n))
    f = orc.fidelity_mc_replay(ctrl, np.linspace(0, 0.1, S), z, n, 0, 2)
    m = orc.metrics(f, 0.05)
    want = np.stack([m[k] for k in rb.engine.STAT_KEYS])
    assert np.array_equal(a, want)
