"""CPU check of the device eigensolver source (csrc/rc_ql.cuh compiles for the host): the pinned-end register
solver the kernels use must give the SAME BITS as the block-at-0 form it replaced and the
oracle's expm value, on the reference's controllers under noise and on matrices that split (zero / negligible
couplings, which the pinned-end form hands to the strided solver out of line).  Builds tools/host_sim.cpp twice
with g++; no GPU, no product code path involved."""
import os, shutil, struct, subprocess, sys
import numpy as np
import pytest
import scipy.linalg

from oracle import robchar_oracle as orc
from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "code-robchar_b200", "csrc")


@pytest.fixture(scope="module")
def sims(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    d = tmp_path_factory.mktemp("host_sim")
    exe = {}
    for name, flags in (("pinned", ["-DRC_QL_PINNED_END=1"]), ("block0", ["-DRC_QL_PINNED_END=0"])):
        exe[name] = str(d / name)
        subprocess.run(["g++", "-O2", "-DRC_QL_STATS", *flags, "-I", CSRC, os.path.join(ROOT, "tools", "host_sim.cpp"),
                        "-o", exe[name]], check=True)
    return exe, d


def _run(exe, d, n, i, o, dd, ee, T, mode=2):
    rec = np.concatenate([dd, ee, T[:, None]], axis=1)
    fin, fout = str(d / "in.bin"), str(d / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("4i", n, i, o, rec.shape[0]))
        f.write(np.ascontiguousarray(rec, dtype=np.float64).tobytes())
    r = subprocess.run([exe, fin, fout, str(mode)], check=True, capture_output=True, text=True)
    return np.fromfile(fout), r.stderr


def _irregular(stderr):
    return int(stderr.split("irregular")[1].split()[0])


@pytest.mark.parametrize("name", ["replay_n4_0_2", "replay_n7_0_6"])
def test_pinned_end_solver_is_bit_identical_on_reference_controllers(sims, name):
    exe, d = sims
    g = load_golden(name + ".npz")
    n, i, o = map(int, g["nio"])
    ctrl = g["ctrl"]
    ctrl = ctrl[np.isfinite(ctrl).all(axis=1)]       # the stored sets are NaN-padded (mcsim.py:436-443)
    rs = np.random.RandomState(5)
    z = rs.standard_normal((ctrl.shape[0], 300, n, 3)) * 0.05
    dd = (ctrl[:, None, :n] + z[..., 0]).reshape(-1, n)
    ee = np.hypot(1.0 + z[..., 1:, 1], z[..., 1:, 2]).reshape(-1, n - 1)
    T = np.abs(np.repeat(ctrl[:, n], 300))
    fp, sp = _run(exe["pinned"], d, n, i, o, dd, ee, T)
    fb, _ = _run(exe["block0"], d, n, i, o, dd, ee, T)
    assert np.array_equal(fp.view(np.int64), fb.view(np.int64))
    assert _irregular(sp) <= 3                       # interior splits are a few in 1e5 at the paper's noise levels
    k = np.arange(0, len(dd), 97)
    ref = np.array([abs(scipy.linalg.expm(-1j * T[j] * (np.diag(dd[j]) + np.diag(ee[j], 1) + np.diag(ee[j], -1)))[o, i]) ** 2 for j in k])
    assert np.abs(fp[k] - ref).max() < 1e-10


@pytest.mark.parametrize("n", [2, 3, 5, 8])
def test_pinned_end_solver_on_matrices_that_split(sims, n):
    exe, d = sims
    rs = np.random.RandomState(n)
    D, E = [], []
    for t in range(1500):
        dd = rs.standard_normal(n) * rs.choice([0.0, 1.0, 10.0])
        ee = 1 + 0.1 * rs.standard_normal(n - 1)
        kind = t % 6
        if kind == 0: ee[rs.randint(n - 1)] = 0.0
        if kind == 1: ee[rs.randint(n - 1)] = 1e-20
        if kind == 2: ee[:] = 0.0
        if kind == 3: dd = (dd + dd[::-1]) / 2; ee = (ee + ee[::-1]) / 2     # mirror symmetric: coincident eigenvalue pairs
        if kind == 4: ee[rs.randint(n - 1)] = 3e-16
        D.append(dd); E.append(ee)
    D, E = np.array(D), np.array(E)
    T = np.abs(rs.standard_normal(len(D))) * 10
    i, o = 0, n - 1
    fp, sp = _run(exe["pinned"], d, n, i, o, D, E, T)
    fb, _ = _run(exe["block0"], d, n, i, o, D, E, T)
    fs, _ = _run(exe["block0"], d, n, i, o, D, E, T, mode=1)      # the strided (shared-memory) solver
    assert np.array_equal(fp.view(np.int64), fb.view(np.int64))
    assert np.abs(fp - fs).max() < 1e-13
    if n > 2:
        assert _irregular(sp) > 500                  # the out-of-line continuation did run
    k = np.arange(0, len(D), 11)
    ref = np.array([abs(scipy.linalg.expm(-1j * T[j] * (np.diag(D[j]) + np.diag(E[j], 1) + np.diag(E[j], -1)))[o, i]) ** 2 for j in k])
    assert np.abs(fp[k] - ref).max() < 1e-10
