from . import Data, Distributions  # noqa: F401
