"""robchar_b200 — B200-native Monte-Carlo robustness-characterisation hot path of RobChar.

Python-facing API mirrors the reference's modules (noise_model, mcsim, wd_sortof_fast_implementation,
rim_analysis); all arithmetic on the path runs in hand-written CUDA behind the C-ABI of
include/robchar_b200.h.  Importing the package does not need a GPU; calling any compute entry
point without the built library or without a CUDA device raises (no CPU fallback).
"""
from . import _lib, engine  # noqa: F401
from . import noise_model, wd_sortof_fast_implementation, rim_analysis, noise_analysis, mcsim, kendall, dist, arim, qnewton  # noqa: F401
from . import RLreinforceXXchain_actionedtime  # noqa: F401
from .mcsim import MCDataSim  # noqa: F401
from .noise_model import noise_function, structured_perturbation, directional_perturbation  # noqa: F401
from .wd_sortof_fast_implementation import wd_from_ideal, wd_from_ideal_zero, RIM_p, compute_dkw_error  # noqa: F401

__version__ = "0.1.0"


def install_reference_module_aliases():
    """Register this package's modules under the reference's flat module names (``mcsim``,
    ``noise_model``, ``wd_sortof_fast_implementation``, ``noise_analysis``,
    ``RLreinforceXXchain_actionedtime``) so unmodified scripts that do ``from mcsim import MCDataSim`` or
    ``from RLreinforceXXchain_actionedtime import Environment`` (ppo.py:16) pick up the GPU path."""
    import sys
    for name, mod in (("mcsim", mcsim), ("noise_model", noise_model),
                      ("wd_sortof_fast_implementation", wd_sortof_fast_implementation),
                      ("noise_analysis", noise_analysis),
                      ("RLreinforceXXchain_actionedtime", RLreinforceXXchain_actionedtime)):
        sys.modules[name] = mod
