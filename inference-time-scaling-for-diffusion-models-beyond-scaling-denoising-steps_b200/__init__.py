"""its_b200 — B200-native inference-time-scaling sampling path.

Host-side mirror of the reference's Python surface (Diffusion.Diffusion,
Diffusion.Model, DiffusionFreeGuidence.*, search.*) over libits_b200.so, a C-ABI
library of hand-written sm_100a kernels.  There is no CPU or torch-op fallback:
every compute call raises if the extension is missing or no CUDA device exists.
"""
__version__ = "0.1.0"
