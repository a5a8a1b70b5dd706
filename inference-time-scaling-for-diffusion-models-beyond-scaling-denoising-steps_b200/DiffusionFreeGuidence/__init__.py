"""Mirror of the reference's `DiffusionFreeGuidence` package for the sampling path
(DiffusionCondition.py, ModelCondition.py)."""
from .DiffusionCondition import GaussianDiffusionSampler  # noqa: F401
from .ModelCondition import UNet  # noqa: F401
