from .dlpm import DLPM, LossType, ModelMeanType, ModelVarType, match_last_dims  # noqa: F401
from .GenerativeLevyProcess import GenerativeLevyProcess  # noqa: F401
from .lim import LIM_sampler, VPSDE  # noqa: F401
