"""B200-native MoDE denoising engine (sm_100a CUDA behind a C ABI) with the reference's Python surface.

Only the hot path of intuitive-robots/MoDE_Diffusion_Policy lives here: MoDeDiT / GCDenoiser / the k-diffusion samplers
(mode/models/networks/modedit.py, mode/models/edm_diffusion/{score_wrappers,gc_sampling}.py).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
