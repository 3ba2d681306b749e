"""Fidelity evaluator with the interface of upstream's ``qnewton.LBFGS`` (the optimisers' objective).

Mirrors the evaluator half of ``qnewton.py``: constructor arguments (:28-42), ``sys_hamiltonian``
(:140-151, incl. the Heisenberg/Z diagonal), ``controls`` (:153-159), ``structured_perturabation``
(:366-379, real symmetric, two draws per site), ``randHset_constructor`` (:122-137, seed 4),
``fidelity_ss`` (:383-423 incl. binomial shot noise), ``fidelity_ss_av`` (:425-444) and ``wass_cost``
(:447-455).  All propagators run on the GPU (RC_MODEL_REAL2 replay rows); random numbers are drawn on
the host from ``np.random`` in upstream's order so a seeded run reproduces upstream's values.
The optimiser loops themselves (``run``) are CPU control flow and out of scope: an optimiser can be
pointed at this class as a drop-in objective (ppo.py:179 builds ``LBFGS(nspin, In, Out, noise=noise)``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import MODEL_REAL2


class LBFGS(object):
    def __init__(self, nspin, in_spin, out_spin, bmin=-10, bmax=10, max_time=30, repeats=1000000, fid_threshold=0.98,
                 topo="linear", noisy=False, fid_noisy=False, draws=10, ham_noisy=False, verbose=False, adp_tol=0.05,
                 adaptive=False, noise=0.05, use_wass_cost=False, heisenberg_int: bool = False,
                 use_fixed_ham: bool = False, opt_train_size: int = 100, opt_test_size: int = 10000, **_ignored):
        self.topo = topo  # "ring" closes the chain (qnewton.py:145-147): evaluated on the dense-expm device path
        self.heisenberg_int = heisenberg_int
        self.Nspin, self.In, self.Out = nspin, in_spin, out_spin
        self.Tmin, self.Tmax, self.Bmin, self.Bmax = 0, max_time, bmin, bmax
        self.repeats = repeats
        self.HH = self.sys_hamiltonian()
        self.CC = self.controls()
        self.fid_threshold = fid_threshold
        self.draws = draws
        self.ham_noisy = ham_noisy
        self.fid_noisy = fid_noisy
        self.verbose = verbose
        self.adp_tol = adp_tol
        self.adaptive = adaptive
        self.adp_func_calls_increment = self.draws
        self.noise = noise
        self.use_wass_cost = use_wass_cost
        self.val_bounds = [(self.Bmin, self.Bmax)] * self.Nspin + [(self.Tmin, self.Tmax)]
        self.use_fixed_ham = use_fixed_ham
        self.train_size = opt_train_size
        self.randH, self.randH_test = self.randHset_constructor(train_size=opt_train_size, test_size=opt_test_size)
        self._rows_train = self._rows_test = None
        self._evaluators = {}

    # ---- model -------------------------------------------------------------------------------------
    def sys_hamiltonian(self):
        n = self.Nspin
        HH = np.zeros((n, n), dtype=np.complex128)
        for l in range(1, n):
            HH[l - 1, l] = 1
            HH[l, l - 1] = 1
        if self.topo == "ring":
            HH[n - 1, 0] = 1
            HH[0, n - 1] = 1
        if self.heisenberg_int:
            t = 0.5 * np.triu(HH).sum().sum() * np.ones(n) - np.sum(HH, axis=1)
            HH += np.diag(t)
        return HH

    def controls(self):
        CC = []
        for k in range(self.Nspin):
            CM = np.zeros((self.Nspin, self.Nspin))
            CM[k, k] = 1
            CC.append(CM)
        return CC

    def structured_perturabation(self):
        """qnewton.py:366-379: (z_ii, nn_i) per site, nn_0 consumed and discarded."""
        n = self.Nspin
        z = np.zeros((n, n), dtype=np.complex128)
        for i in range(n):
            z[i][i] = np.random.normal(scale=self.noise)
            nn = np.random.normal(scale=self.noise)
            if i >= 1:
                z[i][i - 1] = nn
                z[i - 1][i] = nn
        return z

    def randHset_constructor(self, train_size=1000, test_size=10000):
        """qnewton.py:122-137 (fixed seed 4, train then test)."""
        np.random.seed(4)
        out_train = np.zeros((train_size, self.Nspin, self.Nspin), dtype="complex128")
        for i in range(train_size):
            out_train[i] = self.HH + self.structured_perturabation()
        out_test = np.zeros((test_size, self.Nspin, self.Nspin), dtype="complex128")
        for i in range(test_size):
            out_test[i] = self.HH + self.structured_perturabation()
        return out_train, out_test

    # ---- packing of explicit Hamiltonians into replay rows (sigma = 1) -----------------------------------
    def _is_real_symmetric_tridiagonal(self, H: np.ndarray) -> bool:
        """True when every matrix of the stack [m][N][N] is real, symmetric and tridiagonal — what a replay row can
        carry.  Anything else (imaginary parts, ring / next-nearest entries, non-symmetric input) must go through the
        dense path: the reference runs expm on the full complex matrix (qnewton.py:395-397)."""
        H = np.asarray(H)
        n = self.Nspin
        band = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= 1
        return (not np.any(np.imag(H) != 0)) and (not np.any(H[..., ~band] != 0)) and \
            np.array_equal(np.diagonal(H, offset=-1, axis1=-2, axis2=-1), np.diagonal(H, offset=1, axis1=-2, axis2=-1))

    def _rows_from_hamiltonians(self, H: np.ndarray) -> np.ndarray:
        """[m][N][N] real-symmetric tridiagonal Hamiltonians (HH + perturbation) -> [m][2N] replay rows."""
        H = np.asarray(H)
        n = self.Nspin
        rows = np.zeros((H.shape[0], 2 * n))
        base = np.real(np.diag(self.HH))
        rows[:, 0::2] = np.real(np.diagonal(H, axis1=1, axis2=2)) - base
        lo = np.real(np.diagonal(H, offset=-1, axis1=1, axis2=2))
        rows[:, 3::2] = lo - 1.0
        return rows

    def _noise_rows(self, reps: int) -> np.ndarray:
        """`reps` draws of structured_perturabation (qnewton.py:366-379) as replay rows [reps][2N]: the 2N normals of
        one perturbation in upstream's order (z_ii, nn_i per site) ARE the row, so one vectorised draw consumes the
        global np.random stream exactly like upstream's scalar calls (same legacy Gaussian sequence)."""
        return np.random.normal(scale=self.noise, size=(reps, 2 * self.Nspin))

    def _dense_fidelities(self, x, H: np.ndarray) -> np.ndarray:
        """|expm(-i T (H_k + diag(x)))[out, in]|^2 for explicit full Hamiltonians (ring topology, complex or
        non-tridiagonal user input): the dense device path (rc_expm_batch)."""
        n = self.Nspin
        H = np.asarray(H, dtype=np.complex128).reshape(-1, n, n) + np.diag(np.asarray(x[:n], dtype=np.float64))
        return engine.dense_fidelity(H, np.full(H.shape[0], abs(x[n])), self.In, self.Out).cpu().numpy()

    def _evaluator(self, m: int, stats: bool) -> "engine.ObjectiveEvaluator":
        """Cached call frame of the low-latency objective entry point for m perturbation rows (0: nominal)."""
        key = (m, stats)
        ev = self._evaluators.get(key)
        if ev is None:
            ev = engine.ObjectiveEvaluator(self.Nspin, self.In, self.Out, m, model=MODEL_REAL2, zz=self.heisenberg_int,
                                           want_fids=not stats, want_stats=stats)
            if len(self._evaluators) < 64:
                self._evaluators[key] = ev
        return ev

    def _eval_rows(self, x, rows: np.ndarray | None) -> np.ndarray:
        """Fidelities of x under the perturbation rows (None: nominal) as a host array."""
        if self.topo != "ring":   # tridiagonal: the low-latency objective entry point (one C call, host in/out)
            if rows is None:
                return self._evaluator(0, False)(x).fids
            return self._evaluator(rows.shape[0], False)(x, rows).fids
        n = self.Nspin
        if rows is None:
            rows = np.zeros((1, 2 * n))
        m = rows.shape[0]
        H = np.broadcast_to(self.HH, (m, n, n)).copy()     # ring: not tridiagonal, dense expm on explicit Hamiltonians
        idx = np.arange(n)
        H[:, idx, idx] += rows[:, 0::2]
        lo = np.arange(1, n)
        H[:, lo, lo - 1] += rows[:, 3::2]
        H[:, lo - 1, lo] += rows[:, 3::2]
        return self._dense_fidelities(x, H)

    def eval_static_fidelity_gradient(self, x):
        """qnewton.py:162-212: infidelity and its gradient w.r.t. biases and time.  Chain topology: from the
        eigendecomposition of the tridiagonal Hamiltonian on the device (rc_fidelity_grad, any Nspin <= 32) instead
        of upstream's N + 1 dense matrix exponentials; ham_noisy adds one structured_perturabation() draw exactly
        like upstream (:179-180).  Ring topology keeps upstream's construction on the dense device path."""
        n = self.Nspin
        if self.topo != "ring":
            rows = self._noise_rows(1) if self.ham_noisy else None
            err, grad = engine.fidelity_grad(np.asarray(x, dtype=np.float64)[None], n, self.In, self.Out, rows=rows,
                                             zz=self.heisenberg_int)
            return float(err[0]), grad[0]
        if 2 * n > 32:
            raise NotImplementedError("ring-topology gradient supports Nspin <= 16 (2N x 2N block exponentials)")
        T = abs(x[n])
        H = self.HH.copy()
        for l in range(n):
            H += x[l] * self.CC[l]
        if self.ham_noisy:
            H += self.structured_perturabation()
        TH = -1j * T * H
        A = np.zeros((n, 2 * n, 2 * n), dtype=np.complex128)
        A[:, 0:n, 0:n] = TH
        A[:, n:2 * n, n:2 * n] = TH
        for l in range(n):
            A[l, n:2 * n, 0:n] = -1j * T * self.CC[l]
        PSI = engine.expm_batch(A).cpu().numpy()
        U = PSI[0, 0:n, 0:n]
        HU = np.dot(H, U)
        grad = np.zeros(n + 1)
        phi = U[self.Out, self.In]
        err = (1 - (phi.real * phi.real + phi.imag * phi.imag))
        for l in range(n):
            z = PSI[l, n:2 * n, 0:n][self.Out, self.In] * phi.conjugate()
            grad[l] -= 2 * z.real
        z = HU[self.Out, self.In] * phi.conjugate()
        grad[n] -= 2 * z.imag
        return err, grad

    # ---- objectives -------------------------------------------------------------------------------------
    def _shot_noise(self, fid):
        if not self.adaptive:
            return np.random.binomial(self.draws, fid) / self.draws        # qnewton.py:407
        a, b = 0.5, 0.5
        mean = a / (a + b)
        var = mean * (1 - mean) / (a + b + 1)
        while np.sqrt(var) > self.adp_tol:                                 # qnewton.py:411-421
            s = np.random.binomial(self.draws, fid)
            a += s
            b += (self.draws - s)
            mean = (a + s) / (a + b + self.draws)
            var = mean * (1 - mean) / (a + b + self.draws + 1)
            self.adp_func_calls_increment += self.draws
        return mean

    def fidelity_ss(self, x, noisy=False, ham_noisy=False, use_fixed_ham=False, rH=None):
        """qnewton.py:383-423."""
        n = self.Nspin
        if use_fixed_ham:
            if rH is None:
                raise AssertionError(f"H cannot be {type(rH)}")
            rH = np.asarray(rH)
            if self.topo == "ring" or not self._is_real_symmetric_tridiagonal(rH):
                fid = float(self._dense_fidelities(x, rH[None])[0])     # full complex matrix, like upstream's expm
                return self._shot_noise(fid) if noisy else fid
            rows = self._rows_from_hamiltonians(rH[None])
        else:
            rows = self._noise_rows(1) if ham_noisy else None
        fid = float(self._eval_rows(x, rows)[0])
        return self._shot_noise(fid) if noisy else fid

    def fidelity_ss_av(self, x, noisy=False, ham_noisy=False, reps=10, test=False):
        """qnewton.py:425-444: mean over the first `reps` fixed training Hamiltonians (or the whole test set)."""
        if self._rows_train is None:
            self._rows_train = self._rows_from_hamiltonians(self.randH)
            self._rows_test = self._rows_from_hamiltonians(self.randH_test)
        rows = self._rows_test if test else self._rows_train[:reps]
        if noisy:   # per-Hamiltonian shot noise is a host binomial draw (qnewton.py:407): needs the individual values
            return float(np.mean([self._shot_noise(v) for v in self._eval_rows(x, rows)]))
        if self.topo == "ring":
            return float(np.mean(self._eval_rows(x, rows)))
        st = self._evaluator(rows.shape[0], True)(x, rows).stats
        return float(1.0 - st[0])                                  # mean fidelity = 1 - W1 (device reduction)

    def wass_cost(self, x, bootstrap_reps=5):
        """qnewton.py:447-455: W1 to delta(1) of `bootstrap_reps` noisy fidelities (host draws in upstream order)."""
        rows = self._noise_rows(bootstrap_reps)
        if self.topo == "ring":
            return float(engine.stats(torch.as_tensor(self._eval_rows(x, rows)).reshape(1, -1), 0.0)[0, 0].item())
        st = self._evaluator(bootstrap_reps, True)(x, rows).stats
        return float(st[0])                                        # W1 to the ideal distribution (wd_from_ideal)

    def infidelity(self, x):
        return 1 - self.fidelity_ss(x, noisy=self.fid_noisy, ham_noisy=self.ham_noisy)

    def run(self, *a, **k):
        raise NotImplementedError("optimiser loops are out of scope; use this class as the objective evaluator")
