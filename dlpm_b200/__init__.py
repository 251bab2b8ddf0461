"""dlpm_b200: B200-native (sm_100a) implementation of the DLPM sampling hot path.

Keeps the reference's ``dlpm/methods`` surface (``GenerativeLevyProcess``, ``DLPM``, the ``Generator``
noise objects) on top of hand-written CUDA kernels behind a C ABI (``include/dlpm_b200.h``,
``dlpm_b200/libdlpm_b200.so``).  See DESIGN.md and INTEGRATION.md.
"""
from . import rng  # noqa: F401
from .datasets.Data import Generator  # noqa: F401
from .datasets.Distributions import gen_normal, gen_sas, gen_skewed_levy  # noqa: F401
from .methods import DLPM, GenerativeLevyProcess, LossType, ModelMeanType, ModelVarType  # noqa: F401
from .rng import manual_seed, set_sample_base  # noqa: F401
from .generation import FusedPost, GenerationManager  # noqa: F401
from .score_nets import load_checkpoint  # noqa: F401

__all__ = ["GenerativeLevyProcess", "DLPM", "Generator", "gen_skewed_levy", "gen_sas", "gen_normal", "ModelMeanType",
           "ModelVarType", "LossType", "manual_seed", "set_sample_base", "rng", "FusedPost", "GenerationManager", "load_checkpoint"]
