"""The PPO agent's spin-chain environment with the interface of upstream's
``RLreinforceXXchain_actionedtime.Environment`` (ppo.py:152 builds it, ppo.py:338-421 drives it).

What upstream computes per ``step`` (RLreinforceXXchain_actionedtime.py:260-279): accumulate the bias action, wrap
it into the bounds, evolve ``|in>`` under ``sys + action (+ noise)`` for ``timestep`` with ``scipy.linalg.expm``
(:147-178) and return the transfer fidelity ``|<out|U|in>|^2`` as the reward (optionally through binomial shot
noise, :192-222); in fixed-Hamiltonian mode the reward is the squared modulus of the MEAN propagator element over
the training set (:153-163).  Because ``step`` resets ``in_state`` to ``|in>`` before returning, every reward is one
(or ``train_size``) transfer amplitude(s) of a real symmetric tridiagonal Hamiltonian — exactly the path's
evaluation — so the evolution runs on the GPU through ``rc_objective_host`` (explicit perturbation rows, complex
amplitudes for the propagator average).  Ring topology and evolutions from a non-basis ``in_state`` (calling
``state()`` repeatedly without ``step``) go through the dense device exponential (``rc_expm_batch``).
All random numbers are drawn on the host from ``np.random`` in upstream's order, so seeded runs reproduce upstream.
The PPO training loop itself is out of scope; this class is the drop-in environment / objective.
"""
from __future__ import annotations

import numpy as np

from . import engine
from ._lib import MODEL_REAL2


class Environment(object):
    "XX(Z) spin-chain environment, linear or ring topology (RLreinforceXXchain_actionedtime.py:14-55)"

    def __init__(self, nspin, in_spin, out_spin, action_vector=None, final_time=6, topo="linear", timestep_res=0.01,
                 max_time=30, bmin=-20, bmax=20, fid_noisy=False, ham_noisy=False, draws=20, adaptive=False,
                 adp_tol=0.05, noise=0.05, transfer_learning=False, heisenberg_int: bool = False,
                 use_fixed_ham=False, opt_train_size=100, opt_test_size=10000):
        self.Nspin, self.in_spin, self.out_spin = nspin, in_spin, out_spin
        self.topo, self.heisenberg_int = topo, heisenberg_int
        self.timestep, self.tres = 0, timestep_res
        self.action = np.zeros(nspin) if action_vector is None else np.diag(action_vector)
        self.sys = self.system_hamiltonian()
        if transfer_learning:                       # perturbed couplings, diagonal removed (:30-35)
            self.sys = self._off_diagonal_only(self.sys + self.structured_perturabation(0.1))
            print(f"old ham {self.sys}")
        self.in_state, self.out_state = self.state_vector(in_spin), self.state_vector(out_spin)
        self.maxtime = max_time
        self.final_time = max_time
        self.min, self.max = bmin, bmax
        self.noise = noise
        self.fid_noisy, self.ham_noisy = fid_noisy, ham_noisy
        self.draws, self.adaptive, self.adp_var_tol = draws, adaptive, adp_tol
        self.adp_func_calls_increment = draws
        self.tf = 0
        self.use_fixed_ham, self.train_size = use_fixed_ham, opt_train_size
        self.randH, self.randH_test = self.randHset_constructor(train_size=opt_train_size, test_size=opt_test_size)
        self._train_rows = None

    # ---- model (host side, upstream's draw order) -----------------------------------------------------------
    def system_hamiltonian(self):
        n = self.Nspin
        J = np.zeros((n, n))
        k = np.arange(1, n)
        J[k - 1, k] = J[k, k - 1] = 1
        if self.topo == "ring":
            J[n - 1, 0] = J[0, n - 1] = 1
        if self.heisenberg_int:                     # :90-92
            J += np.diag(0.5 * np.triu(J).sum().sum() * np.ones(n) - np.sum(J, axis=1))
        return J

    def _off_diagonal_only(self, H):
        return H * (np.ones_like(H) - np.eye(self.Nspin))

    def reinit_sys_hamiltonian(self):
        self.sys = self._off_diagonal_only(self.system_hamiltonian() + self.structured_perturabation(.1))
        self._train_rows = None
        print(f"new ham: {self.sys}")

    def state_vector(self, occ):
        psi = np.zeros(self.Nspin)
        psi[occ] = 1
        return psi

    def input_state(self):
        return np.outer(self.state_vector(self.in_spin), self.state_vector(self.in_spin))

    def output_state(self):
        return np.outer(self.state_vector(self.out_spin), self.state_vector(self.out_spin))

    def structured_perturabation(self, noise):
        """:122-133: per site one diagonal and one nearest-neighbour draw (site 0's coupling draw is consumed
        and discarded); next-nearest terms are hard zero."""
        n = self.Nspin
        z = np.zeros((n, n))
        for i in range(n):
            z[i][i] = np.random.normal(scale=noise)
            nn = np.random.normal(scale=noise)
            if i >= 1:
                z[i][i - 1] = z[i - 1][i] = nn
        return z

    def randHset_constructor(self, train_size=1000, test_size=10000):
        """:57-72 (fixed seed 4, train then test)."""
        np.random.seed(4)
        sets = []
        for size in (train_size, test_size):
            H = np.zeros((size, self.Nspin, self.Nspin), dtype="complex128")
            for i in range(size):
                H[i] = self.sys + self.structured_perturabation(self.noise)
            sets.append(H)
        return sets[0], sets[1]

    # ---- device evaluation --------------------------------------------------------------------------------
    def _wrapped_time(self):
        t = abs(self.timestep)
        return t % self.maxtime if t > self.maxtime else t

    def _rows(self, H):
        """[m][N][N] real symmetric tridiagonal Hamiltonians (without the action) -> [m][2N] perturbation rows
        relative to the unit chain (diagonal entry, coupling - 1), the real 2-draw replay layout."""
        H = np.real(np.asarray(H)).reshape(-1, self.Nspin, self.Nspin)
        rows = np.zeros((H.shape[0], 2 * self.Nspin))
        rows[:, 0::2] = np.diagonal(H, axis1=1, axis2=2)
        rows[:, 3::2] = np.diagonal(H, offset=-1, axis1=1, axis2=2) - 1.0
        return rows

    def _is_real_symmetric_tridiagonal(self, H) -> bool:
        """What a replay row can carry; cached per array object (the fixed training stack is checked once)."""
        key = id(H)
        hit = getattr(self, "_tridiag_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        Hs = np.asarray(H).reshape(-1, self.Nspin, self.Nspin)
        band = np.abs(np.subtract.outer(np.arange(self.Nspin), np.arange(self.Nspin))) <= 1
        ok = (not np.any(np.imag(Hs) != 0)) and (not np.any(Hs[:, ~band] != 0)) and \
            np.array_equal(np.diagonal(Hs, offset=-1, axis1=1, axis2=2), np.diagonal(Hs, offset=1, axis1=1, axis2=2))
        if isinstance(H, np.ndarray):
            self._tridiag_cache = (key, ok)
        return ok

    def _amplitudes(self, H, action, t):
        """<out| exp(-i t (H_k + action)) |in> for every Hamiltonian of the stack H [m][N][N]."""
        action = np.asarray(action, dtype=np.float64)
        bias = np.diag(action) if action.ndim == 2 else action
        tridiag = self._is_real_symmetric_tridiagonal(H)
        # replay rows carry real symmetric tridiagonal Hamiltonians only; anything else takes the dense path below
        if self.topo != "ring" and tridiag and (action.ndim != 2 or not np.any(action - np.diag(bias))):
            x = np.concatenate([bias, [t]])
            _, amps = engine.objective_host(x, self._rows(H), self.Nspin, self.in_spin, self.out_spin, model=MODEL_REAL2,
                                            want_amps=True)
            return amps
        U = self._propagators(np.asarray(H) + (action if action.ndim == 2 else np.diag(bias)), t)
        return U[:, self.out_spin, self.in_spin]

    def _propagators(self, H, t):
        """Dense exp(-i t H_k) on the device (ring topology, non-diagonal actions, non-basis in_state)."""
        H = np.asarray(H, dtype=np.complex128).reshape(-1, self.Nspin, self.Nspin)
        return engine.expm_batch(-1j * t * H).cpu().numpy()

    # ---- upstream's stepping API ----------------------------------------------------------------------------
    def state(self, action=None):
        """:147-178: evolve in_state by one application of the propagator (mean propagator in fixed-Hamiltonian
        mode).  Returns U like upstream (None in fixed-Hamiltonian mode)."""
        action = self.action if action is None else action
        self.timestep = self._wrapped_time()
        action = np.asarray(action, dtype=np.float64)
        amat = action if action.ndim == 2 else np.diag(action)
        if self.use_fixed_ham:
            U = self._propagators(self.randH[:self.train_size] + amat, self.timestep).mean(axis=0)
            self.in_state = np.matmul(U, self.in_state)
            return None
        H = self.sys + amat
        if self.ham_noisy:
            H = H + self.structured_perturabation(self.noise)
        U = self._propagators(H, self.timestep)[0]
        self.in_state = np.matmul(U, self.in_state)
        return U

    def reset(self):
        self.timestep = 0
        self.in_state = self.state_vector(self.in_spin)
        self.action = np.zeros((self.Nspin, self.Nspin))
        # upstream calls state() here (:184): with timestep 0 the propagator is the identity, but in noisy-
        # Hamiltonian mode that call consumes one perturbation's worth of draws — keep the stream aligned
        if self.ham_noisy and not self.use_fixed_ham:
            self.structured_perturabation(self.noise)
        return self.action

    def _shot_noise(self, fid):
        """:205-222: binomial estimate with `draws` shots, or the adaptive Beta-posterior scheme."""
        sample = np.random.binomial(self.draws, fid)
        if not self.adaptive:
            return sample / self.draws
        a, b = 0.5, 0.5
        mean = a / (a + b)
        var = mean * (1 - mean) / (a + b + 1)
        while np.sqrt(var) > self.adp_var_tol:
            s = np.random.binomial(self.draws, fid)
            a += s
            b += (self.draws - s)
            mean = (a + s) / (a + b + self.draws)
            var = mean * (1 - mean) / (a + b + self.draws + 1)
            self.adp_func_calls_increment += self.draws
        return mean

    def fidelity(self):
        """:192-222 on the current in_state."""
        overlap = np.matmul(np.conj(self.in_state).T, self.out_state)
        fid = np.real(np.conj(overlap) * overlap)
        return self._shot_noise(fid) if self.fid_noisy else fid

    def _true_fid_single(self, action, base_H=None, timestep_n=None):
        """:225-234.  Upstream evaluates ``self.sys + action`` whatever base_H is (the argument is unused there);
        kept, so true_fid in fixed-Hamiltonian mode is the nominal fidelity at timestep_n."""
        if base_H is None:
            timestep_n = self.timestep
        amp = self._amplitudes(self.sys[None], action, timestep_n)[0]
        fid = np.real(np.conj(amp) * amp)
        if not np.array_equal(self.in_state, self.state_vector(self.in_spin)):     # upstream evolves in_state itself
            v = np.matmul(self._propagators(self.sys + np.asarray(action), timestep_n)[0], self.in_state)
            ov = np.matmul(np.conj(v).T, self.out_state)
            fid = np.real(np.conj(ov) * ov)
        return fid

    def true_fid(self, action, timestep_n=None):
        if self.use_fixed_ham:
            return self._true_fid_single(action, base_H=self.randH_test[0], timestep_n=timestep_n)
        return self._true_fid_single(action)

    def _wrap_action(self):
        if (np.abs(self.action) > self.max).any():
            self.action = self.action % np.diag(np.sign(self.action) * self.max)

    def normalize(self):
        self._wrap_action()
        self.timestep = self._wrapped_time()

    def step(self, action):
        """:260-279.  One GPU call (two without fixed Hamiltonians: the noise-free `tf`, then the reward)."""
        if np.shape(self.action) == np.shape(action):
            self.action += action                    # in place like upstream: reset()'s return value tracks the state
        else:
            self.action = self.action + action
        self._wrap_action()
        try:
            if not self.use_fixed_ham:
                self.tf = self.true_fid(self.action)
            self.timestep = self._wrapped_time()
            if self.use_fixed_ham:
                if self._train_rows is None:
                    self._train_rows = np.ascontiguousarray(self.randH[:self.train_size])
                amp = self._amplitudes(self._train_rows, self.action, self.timestep).mean()
            else:
                H = self.sys + self.structured_perturabation(self.noise) if self.ham_noisy else self.sys
                amp = self._amplitudes(H[None], self.action, self.timestep)[0]
            fid = np.real(np.conj(amp) * amp)
            reward = self._shot_noise(fid) if self.fid_noisy else fid
            done_flag = bool(self.timestep > self.final_time)
            self.in_state = self.state_vector(self.in_spin)
            return self.action, reward, done_flag
        except ValueError as e:
            print(e)
            return np.zeros_like(self.action), 0, False
