"""Mirror of the reference's `Diffusion` package for the sampling path
(Diffusion/Diffusion.py, Diffusion/Model.py).  Training code is out of scope."""
from .Diffusion import GaussianDiffusionSampler, extract  # noqa: F401
from .Model import UNet  # noqa: F401
