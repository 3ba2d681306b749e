"""Cachable Monte-Carlo simulation class with the reference's API; the sweep runs on the GPU.

Mirrors upstream ``mcsim.py``: ``MCDataSim`` (:200-660) with the same constructor, attributes,
file-name grammar and JSON layouts, the metric registry (:144-183) and the small helpers ``get_cdf``,
``get_supcdf``, ``vn_test`` (:42-123, restated).  Plotting / TSNE parts and the controller-file merge
utilities (:594-649) are not part of the compute path and are omitted.

Differences that are deliberate and documented:
* noise: by default drawn in-kernel with counter-based Philox (``seed``); ``rng_mode="numpy"``
  instead consumes the global ``np.random`` stream exactly like upstream (one discarded draw per
  noise level, mcsim.py:425; 3N draws per evaluation; NaN-padded controllers consume nothing) and
  replays it on the device — bit-identical sampling for parity runs.
* JSON ``null`` controllers (present in some ``noisy_analysis/*_in`` files) become NaN rows; upstream
  would raise on them.
"""
from __future__ import annotations

import glob
import json
import os
from dataclasses import dataclass
from typing import Callable, List

import numpy as np

from . import engine
from .noise_analysis import DirectoryDoesNotExistError, ExperimentNamer
from .noise_model import structured_perturbation
from .wd_sortof_fast_implementation import compute_dkw_error, wd_from_ideal


def check_numpytype(f):
    def method(arrays, *args, **kwargs):
        if type(arrays) == np.ndarray and len(arrays.shape) == 1:
            return f(arrays, *args, **kwargs)
        raise TypeError("make sure arg is a numpy array")
    return method


@check_numpytype
def get_cdf(arrays):
    """mcsim.py:42-47."""
    sarrays = np.sort(arrays)
    return sarrays.cumsum() / sarrays.sum(), sarrays


@check_numpytype
def get_supcdf(cdf):
    """Tail averages of a cdf vector: out[i] = mean(cdf[i:]) (what mcsim.py:50-57 computes with a loop)."""
    tail_sums = np.cumsum(cdf[::-1])[::-1]
    return tail_sums / np.arange(cdf.size, 0, -1)


@check_numpytype
def vn_test(obs_v, alpha=0.95, verbose=True, bartels=True):
    """Von Neumann ratio test of serial independence (same decision rule and return shape as mcsim.py:59-123):
    statistic = mean squared successive difference / population variance, ~2 for independent samples.
    bartels=True uses the fixed cut 1.1; otherwise the normal quantile at 1 - alpha with the ratio's exact mean
    2n/(n-1) and variance 4 n^2 (n-2) / ((n+1)(n-1)^3)."""
    n = obs_v.size
    if n < 40:
        raise Exception("{} nobs are insufficient for the test.".format(n))
    ratio = np.mean(np.square(obs_v[1:] - obs_v[:-1])) / np.var(obs_v)
    if bartels:
        if verbose:
            print(ratio)
        return (bool(ratio > 1.1), ratio)
    from scipy.stats import norm
    centre = 2.0 * n / (n - 1)
    spread = np.sqrt(4.0 * n * n * (n - 2) / ((n + 1) * (n - 1) ** 3))
    cut = norm.ppf(1 - alpha, loc=centre, scale=spread)
    return (bool(ratio > cut), cut)


def ovlen(obj):
    """Number of entries of a container, 1 for a scalar (mcsim.py:133-142)."""
    if isinstance(obj, (int, float)):
        return 1
    if isinstance(obj, dict):
        return len(obj)
    if hasattr(obj, "__len__") and not isinstance(obj, str):
        return len(obj)
    raise TypeError("unknown data type encountered")


# -- metric registry (mcsim.py:144-183): same names and call shapes, device-batched bodies ----------
def _stat_rows(fids, row):
    arr = np.ascontiguousarray(np.asarray(fids, dtype=np.float64))
    return engine.stats(arr, 0.0)[row].cpu().numpy()


@check_numpytype
def Q(fid_array, threshold):
    return len(fid_array[fid_array >= threshold]) / len(fid_array)


def wc_fids(fids):
    return iter(_stat_rows(fids, 12))


def std_fids(fids):
    return iter(_stat_rows(fids, 9))


def Q_fids(fids, threshold=0.95):
    if threshold == 0.95:
        return iter(_stat_rows(fids, 3))
    if threshold == 0.98:
        return iter(_stat_rows(fids, 6))
    a = np.asarray(fids)
    return iter(-1 * (a >= threshold).sum(axis=-1) / a.shape[-1])


def wd_from_ideal_fids(fids):
    return iter(_stat_rows(fids, 0))


@dataclass
class Q_partial:
    qthres: float = 0.95

    def Q_fids(self, fids) -> Callable[[List[float]], List[float]]:
        return Q_fids(fids, self.qthres)


__metric_name_to_metric__ = {engine.METRIC_W: wd_from_ideal_fids,
                             "Q th. 0.95": Q_partial(qthres=0.95).Q_fids,
                             "Q th. 0.98": Q_partial(qthres=0.98).Q_fids,
                             "std": std_fids,
                             "worst case fid": wc_fids,
                             }


class MCDataSim:
    "A class for MC data generation with structured perturbations of XX-controllers (GPU sweep)."

    def __init__(self, experiment_name: str = "pipeline_alpha", Nspin: int = 5,
                 inspin: int = 0, outspin: int = 2,
                 noises: np.ndarray = np.linspace(0, 0.1, 11),
                 bootreps: int = 100, training_noise: float = None,
                 numcontrollers: int = 100, parallel: bool = False,
                 num_workers: int = None,
                 dkw_conflvl: float = 0.95,
                 filemarker: str = None,
                 topk: int = 100,
                 seed: int = 0, rng_mode: str = "philox", verbose: bool = False):
        self.global_experiments_directory = "experiments/"
        self.filemarker = filemarker
        self.experiment_name = experiment_name
        self.topk = topk
        self.args = dict(Nspin=Nspin, inspin=inspin, outspin=outspin)
        self.bootreps = bootreps
        self.alpha = 1 - dkw_conflvl
        self.training_noise = training_noise
        self.Nspin = Nspin
        self.inspin = inspin
        self.outspin = outspin
        self.noises = noises
        self.numcontrollers = numcontrollers
        self.seed = seed
        self._fresh_metrics = {}
        if rng_mode not in ("philox", "numpy"):
            raise ValueError("rng_mode must be 'philox' or 'numpy'")
        self.rng_mode = rng_mode
        self.verbose = verbose

        self.get_controller_name = self.get_experiment_name(experiment_name)()
        if self.filemarker is not None:
            self.get_controller_name += self.filemarker
        if self.verbose:
            print(self.get_controller_name)
        try:
            self.controllers = self.load_controllers()
            self.algos = self.ctrlnames(self.controllers)
        except FileNotFoundError as e:
            print("flagging: ", e)
            self.controllers = None
            self.algos = None

        self.noise_model = structured_perturbation(**self.args)
        self.parallel = parallel        # accepted for compatibility; the device sweep is always batched
        self.num_workers = num_workers
        self.colors = ["blue", "orange", "gold", "purple", "pink", "brown",
                       "red", "cyan", "gray", "mediumseagreen", "olive"]
        self.figlabels = ["({})".format(i) for i in "abcdefghijklmnopqrstuvwxyz"]
        self._algo_counter = 0

    # ---- controller containers -------------------------------------------------------------------
    def get_all_algo_controllers(self):
        "combine all algo controllers (mcsim.py:251-264)"
        cs = []
        for alg in list(self.controllers.keys()):
            if alg == "lbfgs":
                conts = np.array(self.controllers[alg][str(self.Nspin)]["controller"])
                if self.numcontrollers - len(conts) > 0:
                    conts = np.pad(conts, [(self.numcontrollers - len(conts), 0), (0, 0)])
                cs.append(conts)
            else:
                for noise in list(self.controllers[alg].keys()):
                    cs.append(np.array(self.controllers[alg][noise]["controller"]))
        return np.array(cs).reshape(-1, self.Nspin + 1)

    @staticmethod
    def bootstrap_resampling_std(summarystatistic, l, bootsamples):
        "mcsim.py:267-275 (host; consumes np.random like upstream)"
        bootsss = np.zeros(bootsamples)
        for i in range(bootsamples):
            randi = np.random.randint(0, len(l), size=len(l))
            bootsss[i] = summarystatistic(l[randi])
        return bootsss.std()

    def ctrlnames(self, ctrlcontainer) -> List:
        "mcsim.py:335-349"
        if isinstance(ctrlcontainer, dict):
            for key in list(ctrlcontainer.keys()):
                if ctrlcontainer[key] == {}:
                    ctrlcontainer.pop(key)
            return list(ctrlcontainer.keys())
        if isinstance(ctrlcontainer, (list, np.ndarray)):
            return ["unnamed"]
        raise TypeError("need controller container either as a list or a dict")

    def get_mcname(self, training_noise=None, noises=None) -> str:
        "mcsim.py:351-356 (the name embeds numpy's str() of the noise array)"
        if training_noise is None:
            training_noise = self.training_noise
        if noises is None:
            noises = self.noises
        return self.get_controller_name + "_tn{}_br_{}_nlvl{}.mc".format(training_noise, self.bootreps, noises)

    def load_controllers(self, controllers=None):
        "mcsim.py:358-364"
        if controllers is None:
            return json.load(open(self.get_controller_name, "rb"))
        if isinstance(controllers, str):
            return json.load(open(controllers, "rb"))
        if isinstance(controllers, (list, np.ndarray)):
            return controllers

    def loadsimdata(self, simname: str):
        return json.load(open(simname, "rb"))

    def get_controller_fid_dist_boot(self, x=None):
        "mcsim.py:369-374: one sample for self.controller (NaN controller -> NaN)"
        if self.controller is not np.nan:
            return self.noise_model.evaluate_noisy_fidelity(self.controller, ham_noisy=True)
        return np.nan

    def get_experiment_name(self, experiment_name: str) -> Callable[[str], ExperimentNamer]:
        return ExperimentNamer(experiment_name=experiment_name, numcontrollers=self.numcontrollers, **self.args)

    # ---- the sweep ---------------------------------------------------------------------------------
    def _controller_matrix(self, algoname, training_noise) -> np.ndarray:
        """[numcontrollers][N+1] with NaN rows beyond the stored list (mcsim.py:428-443)."""
        key = str(self.Nspin) if algoname == "lbfgs" else str(training_noise)
        conts = self.controllers[algoname][key]["controller"]
        n = self.Nspin
        out = np.full((self.numcontrollers, n + 1), np.nan)
        for i, c in enumerate(conts[:self.numcontrollers]):
            if c is not None and not (isinstance(c, float) and np.isnan(c)):
                out[i] = np.asarray(c, dtype=np.float64)
        return out

    def _numpy_stream_replay(self, ctrl: np.ndarray, noises: np.ndarray) -> np.ndarray:
        """Consume np.random exactly as upstream's loops do and return STANDARD normals [S][C][B][3N]."""
        S, Cn, B, K = len(noises), ctrl.shape[0], self.bootreps, 3 * self.Nspin
        valid = ~np.isnan(ctrl).any(axis=1)
        z = np.zeros((S, Cn, B, K))
        for s in range(S):
            self.noise_model.rng(scale=noises[s])  # mcsim.py:425 (draws once, discarded)
            for c in range(Cn):
                if valid[c]:
                    z[s, c] = np.random.standard_normal((B, K))
        return z

    def simulate_fid_tensor(self, algoname: str, noises, training_noise):
        """Device fidelity tensor [S][C][B] for one controller group."""
        ctrl = self._controller_matrix(algoname, training_noise)
        noises = np.asarray(noises, dtype=np.float64)
        replay = None
        if self.rng_mode == "numpy":
            replay = self._numpy_stream_replay(ctrl, noises)
        else:
            self.noise_model.rng.args["scale"] = noises[-1] if len(noises) else self.noise_model.noise
        seed = (self.seed * 1000003 + self._algo_counter) & (2**63 - 1)
        self._algo_counter += 1
        return engine.fidelity_mc(ctrl, noises, self.bootreps, self.Nspin, self.inspin, self.outspin, seed=seed,
                                  replay=replay)

    def get_algo_fid_dist(self, algoname: str, allalgoallfids: dict, noises, training_noise):
        "mcsim.py:422-460: fills allalgoallfids[algoname] with nested lists [S][C][B] and dumps the .mc"
        fids = self.simulate_fid_tensor(algoname, noises, training_noise)
        # the metric tensors of a freshly simulated group are taken from the device tensor right here (one launch)
        # instead of after the list -> JSON -> ndarray -> device round trip; get_metrics_dict picks them up
        dkw_error = float(compute_dkw_error(self.alpha, self.bootreps))
        st = engine.stats_unsorted(fids, dkw_error).cpu().numpy() if fids.numel() else None
        allalgoallfids[algoname] = fids.cpu().numpy().tolist()
        if st is not None:
            self._fresh_metrics[algoname] = (allalgoallfids[algoname], {k: st[i].tolist() for i, k in enumerate(engine.STAT_KEYS)})
        with open(self.get_mcname(training_noise, noises), "w") as fh:
            json.dump(allalgoallfids, fh)
        return allalgoallfids

    def get_fid_dists(self, training_noise: str = None, noises: np.ndarray = None, algoname=None) -> dict:
        "mcsim.py:382-419 (cache protocol incl. the lbfgs -> training_noise=None quirk)"
        if isinstance(algoname, str):
            algos = [algoname]
        elif algoname is None:
            algos = self.algos
        if noises is None:
            noises = self.noises
        if training_noise is None:
            training_noise = self.training_noise

        if os.path.exists(self.get_mcname(training_noise, noises)):
            simdict = self.loadsimdata(self.get_mcname(training_noise, noises))
            for algoname in algos:
                if algoname not in simdict:
                    self.get_algo_fid_dist(algoname, simdict, noises, training_noise)
            for algoname in simdict.keys():
                if algoname not in algos:
                    raise Exception(f"Fid distribution generation for {algoname} was unsuccessful.")
            return simdict
        allalgoallfids = {}
        for algoname in algos:
            if algoname == "lbfgs":
                training_noise = None
            self.get_algo_fid_dist(algoname, allalgoallfids, noises, training_noise)
        for algoname in allalgoallfids.keys():
            if algoname not in algos:
                raise Exception(f"Fid distribution generation for {algoname} was unsuccessful.")
        return allalgoallfids

    def metrics_from_tensor(self, dists_tensor) -> dict:
        """The 15 metric tensors of one algo (mcsim.py:480-498) from a fidelity tensor [S][C][B]:
        one device call (segmented sort + fused reductions, DKW shift applied to the values)."""
        dkw_error = float(compute_dkw_error(self.alpha, self.bootreps))
        st = engine.stats(np.ascontiguousarray(np.asarray(dists_tensor, dtype=np.float64)), dkw_error).cpu().numpy()
        return {k: st[i].tolist() for i, k in enumerate(engine.STAT_KEYS)}

    def get_metrics_dict(self, training_noise: str = None, noises: np.ndarray = None, algoname=None):
        "mcsim.py:463-510"
        if training_noise is None:
            training_noise = self.training_noise
        if noises is None:
            noises = self.noises

        def get_metric_dict_from_scratch(algos, algoname):
            algofiddists = self.get_fid_dists(training_noise, noises, algoname)
            allalgos_metrics_dict = {}
            for algo in algos:
                fresh = self._fresh_metrics.pop(algo, None)
                if fresh is not None and fresh[0] is algofiddists[algo]:      # simulated in this call: already reduced
                    allalgos_metrics_dict[algo] = fresh[1]
                else:
                    allalgos_metrics_dict[algo] = self.metrics_from_tensor(algofiddists[algo])
            with open(self.get_mcname(training_noise, noises) + "m", "w") as fh:
                json.dump(allalgos_metrics_dict, fh)
            return allalgos_metrics_dict

        if os.path.exists(self.get_mcname(training_noise, noises) + "m"):
            return self.loadsimdata(self.get_mcname(training_noise, noises) + "m")
        return get_metric_dict_from_scratch(algos=self.algos, algoname=None)

    # ---- ranking helpers ---------------------------------------------------------------------------
    @staticmethod
    def get_ranks(array):
        "mcsim.py:513-518: ordinal ranks (ascending, NaN last, ties by index) computed on the device"
        return engine.ranks(np.asarray(array, dtype=np.float64)).cpu().numpy()

    def get_best_controller_perf(self, metric_data, algo=None, contcount=None):
        "mcsim.py:520-545: rank-sum best / median controller; ranks from the device"
        metric_data = np.asarray(metric_data, dtype=np.float64)
        if contcount is None:
            contcount = self.numcontrollers
        ranks = engine.ranks(metric_data).cpu().numpy()
        assert metric_data[-1][np.argmin(ranks[-1])] == np.min(metric_data[-1]), "rank order needs to be metric ascending"
        best_across_plot_noises = ranks.sum(axis=0)
        if best_across_plot_noises.size != contcount:
            print("summation axis is incorrect!")
        bests_nranks = engine.ranks(best_across_plot_noises.astype(np.float64)).cpu().numpy()
        order = np.empty_like(bests_nranks)
        order[bests_nranks] = np.arange(len(bests_nranks))  # stable argsort from ranks
        best_controller_index = order[0]
        median_controller_index = order[metric_data.shape[-1] // 2]
        best_per_noise = np.min(metric_data, axis=1)
        best_controller_per_noise = metric_data[:, best_controller_index]
        median_controller_per_noise = metric_data[:, median_controller_index]
        diff_median = median_controller_per_noise - best_per_noise
        diff = best_controller_per_noise - best_per_noise
        return diff, diff_median, best_controller_per_noise, median_controller_per_noise, best_per_noise

    def get_top_k_by_fid_idx(self, wd_data_c, topk, idx=0):
        "mcsim.py:548-551"
        wd_data_c = np.asarray(wd_data_c)
        filmask = self.get_ranks(wd_data_c[idx]) <= topk - 1
        return np.ix_(np.ones(wd_data_c.shape[0], dtype=bool), filmask)

    def get_top_k_by_fid(self, wd_data_c, wd_data_u, wd_data_l, topk, fid_thres=0.8):
        "mcsim.py:651-660"
        wd_data_c = np.asarray(wd_data_c)
        filmask = self.get_ranks(wd_data_c[0]) <= topk - 1
        if fid_thres:
            filmask &= wd_data_c[0] <= 1 - fid_thres
        idx = np.ix_(np.ones(wd_data_c.shape[0], dtype=bool), filmask)
        return wd_data_c[idx], np.array(wd_data_u)[idx], np.array(wd_data_l)[idx]

    # ---- locating an experiment's files (mcsim.py:572-592) ---------------------------------------------
    def get_path(self, directory_exportable, of: str = "controllers"):
        """Controller file of another experiment directory, or the list of its .mc / .mcm caches."""
        if not os.path.exists(self.global_experiments_directory + directory_exportable):
            raise DirectoryDoesNotExistError(self.global_experiments_directory)
        stem = self.get_experiment_name(directory_exportable)() + (self.filemarker or "")
        if not os.path.exists(stem):
            raise DirectoryDoesNotExistError(stem)
        if of == "controllers":
            return stem
        if of in ("mc", "mcm"):
            return glob.glob(f"{stem}**.{of}")
        raise Exception("No such object type exists. Please specify a correct .description.")
