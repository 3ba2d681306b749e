"""Perturbation spec and single/batched noisy-fidelity evaluator with the reference's API.

Mirrors upstream ``noise_model.py`` (noise_function :21-46, noise_model_base :49-115,
structured_perturbation :117-147, directional_perturbation :150-201): same class names,
constructor arguments, attributes and call semantics.  The propagator/fidelity arithmetic runs on
the GPU (real-tridiagonal eigensolver behind the C-ABI); the perturbation *specification* — which
random numbers are drawn, in which order, from which generator — stays on the host exactly as
upstream so that a seeded ``np.random`` stream gives the same samples.
"""
from __future__ import annotations

import numpy as np

from . import engine
from ._lib import MODEL_COMPLEX3


class noise_function:
    """noise_model.py:21-46: callable wrapping a generator and its persistent kwargs; every call
    updates the stored kwargs *and draws* (``rng(scale=s)`` is how mcsim changes the level)."""

    def __init__(self, generator, **args):
        self.generator = generator
        self.args = args

    def __call__(self, **extraargs):
        for arg in extraargs:
            self.args[arg] = extraargs[arg]
        return self.generator(**self.args)


class noise_model_base:
    """noise_model.py:49-115."""

    def __init__(self, Nspin: int = 5, inspin: int = 0, outspin: int = 2, noise: float = 0.02,
                 topo: str = "chain", rng: noise_function = None, zz: bool = False):
        self.Nspin = Nspin
        self.inspin = inspin
        self.outspin = outspin
        self.noise = noise
        self.topo = topo
        self.zz = zz  # extension: Heisenberg/Z diagonal of qnewton.py:148-150 (BASELINE config 4)
        self.rng = self.default_gaussian_noise_generator(scale=self.noise) if rng is None else rng
        self.HH = np.zeros((Nspin, Nspin), dtype=np.complex128)
        for l in range(1, self.Nspin):
            self.HH[l - 1, l] = 1
            self.HH[l, l - 1] = 1
        if topo == "ring":
            self.HH[self.Nspin - 1, 0] = 1
            self.HH[0, self.Nspin - 1] = 1
        self.CC = self.controls()

    def controls(self):
        CC = []
        for k in range(0, self.Nspin):
            CM = np.zeros((self.Nspin, self.Nspin))
            CM[k, k] = 1
            CC.append(CM)
        return CC

    # -- the drop-in single evaluation ---------------------------------------------------------
    def evaluate_noisy_fidelity(self, x, ham_noisy: bool = False):
        """noise_model.py:98-109.  The perturbation is drawn on the host by ``self.perturbation()``
        (same generator calls, same order as upstream); the evolution runs on the GPU."""
        n = self.Nspin
        z = np.asarray(self.perturbation()) if ham_noisy else None
        if self.topo == "chain" and (z is None or self._is_hermitian_tridiagonal(z)):
            # low-latency entry point (rc_objective_host): one C call, host in/out
            rows = self._replay_row_from_matrix(z).reshape(1, 3 * n) if z is not None else None
            return float(engine.objective_host(x, rows, n, self.inspin, self.outspin, model=MODEL_COMPLEX3, zz=self.zz)[0])
        # generality path (ring topology, non-tridiagonal or non-Hermitian perturbations such as
        # directional_perturbation's complex diagonal entries): dense expm on the device
        H = self.HH.copy()
        if self.zz:
            H = H + np.diag(0.5 * np.triu(self.HH).sum().sum() * np.ones(n) - np.sum(self.HH, axis=1))
        if z is not None:
            H = H + z
        H = H + np.diag(np.asarray(x[:n], dtype=np.float64))
        return float(engine.dense_fidelity(H, abs(x[n]), self.inspin, self.outspin)[0].item())

    def _is_hermitian_tridiagonal(self, z: np.ndarray) -> bool:
        n = self.Nspin
        band = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= 1
        return (not np.any(z[~band] != 0)) and np.array_equal(z, z.conj().T) and not np.any(np.diag(z).imag != 0)

    def _replay_row_from_matrix(self, z: np.ndarray) -> np.ndarray:
        """Pack a Hermitian tridiagonal perturbation matrix into one replay row (sigma = 1)."""
        n = self.Nspin
        if not self._is_hermitian_tridiagonal(z):
            raise NotImplementedError("the eigensolver fast path needs a Hermitian tridiagonal perturbation")
        row = np.zeros(3 * n)
        row[0::3] = np.diag(z).real
        lo = np.diag(z, -1)  # z[i, i-1] = nn + 1j*nn2
        row[4::3] = lo.real
        row[5::3] = lo.imag
        return row

    # -- batched evaluation (the fast path) ----------------------------------------------------
    def evaluate_noisy_fidelity_batch(self, X, draws: int = 1, ham_noisy: bool = True, noises=None, seed: int = 0,
                                      as_numpy: bool = True):
        """All (noise level, controller, draw) fidelities in one launch with in-kernel Philox noise.
        X: [C][N+1]; noises: iterable of sigma (default: the generator's current ``scale``).
        Returns [S][C][draws]."""
        if noises is None:
            noises = [self.rng.args.get("scale", self.noise) if ham_noisy else 0.0]
        sig = np.asarray(noises, dtype=np.float64) if ham_noisy else np.zeros(len(noises))
        out = engine.fidelity_mc(np.asarray(X, dtype=np.float64).reshape(-1, self.Nspin + 1), sig, draws, self.Nspin,
                                 self.inspin, self.outspin, model=MODEL_COMPLEX3, zz=self.zz, seed=seed)
        return out.cpu().numpy() if as_numpy else out

    def perturbation(self) -> np.ndarray:
        raise NotImplementedError

    def default_gaussian_noise_generator(self, **genargs):
        return noise_function(np.random.normal, **genargs)


class structured_perturbation(noise_model_base):
    """noise_model.py:117-147."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def perturbation(self) -> np.ndarray:
        # three draws per site in the order (z_ii, nn_i, nn2_i); site 0's coupling draws are
        # consumed and discarded; next-nearest terms are zero (noise_model.py:135-146)
        n = self.Nspin
        z = np.zeros((n, n), dtype=np.complex128)
        for i in range(n):
            z[i][i] = self.rng()
            nn = self.rng()
            nn2 = self.rng()
            if i >= 1:
                z[i][i - 1] = nn + 1j * nn2
                z[i - 1][i] = nn - 1j * nn2
        return z


class directional_perturbation(noise_model_base):
    """noise_model.py:150-201: one random Hermitian pair of entries."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.directions = [(0, 0), (self.Nspin - 1, self.Nspin - 1)]
        for d in range(1, self.Nspin - 1):
            for o in [-1, 0, 1]:
                self.directions.append((d, d + o))
        self.directions.append((0, 1))
        self.directions.append((1, 0))
        self.directions.append((self.Nspin - 2, self.Nspin - 1))
        self.directions.append((self.Nspin - 1, self.Nspin - 2))

    def perturbation(self) -> np.ndarray:
        pert_index = self.directions[np.random.randint(low=0, high=len(self.directions))]
        pert_index2 = (pert_index[1], pert_index[0])
        z = np.zeros((self.Nspin, self.Nspin), dtype=np.complex128)
        nval = self.rng(size=2)
        z[pert_index] = nval[0] + 1j * nval[1]
        z[pert_index2] = nval[0] - 1j * nval[1]
        return z

    def evaluate_noisy_fidelity_batch(self, X, draws: int = 1, ham_noisy: bool = True, noises=None, seed: int = 0,
                                      as_numpy: bool = True, replay=None):
        """All (noise level, controller, draw) fidelities of a directional-perturbation sweep in one call
        (rc_directional_fidelity_mc): upstream evaluates them one at a time (noise_model.py:98-109 with :165-201).
        replay: optional [S][C][draws][3] = (direction index, n0, n1), the values np.random.randint(0, 3N) and
        rng(size=2) / sigma return upstream; default: in-kernel Philox draws.  Returns [S][C][draws]."""
        if noises is None:
            noises = [self.rng.args.get("scale", self.noise) if ham_noisy else 0.0]
        sig = np.asarray(noises, dtype=np.float64) if ham_noisy else np.zeros(len(noises))
        out = engine.directional_fidelity_mc(np.asarray(X, dtype=np.float64).reshape(-1, self.Nspin + 1), sig, draws,
                                             self.Nspin, self.inspin, self.outspin, zz=self.zz, ring=self.topo == "ring",
                                             seed=seed, replay=replay)
        return out.cpu().numpy() if as_numpy else out
