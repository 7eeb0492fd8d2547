"""B200-native (sm_100a) implementation of the MM-Diffusion denoising hot path.

Public surface mirrors the reference package (mm_diffusion/): see `mm_diffusion_b200.unet`,
`mm_diffusion_b200.gaussian_diffusion`, `mm_diffusion_b200.script_util`.  The compute path is
libmmdiff.so (hand-written CUDA behind the C-ABI in include/mmdiff.h); there is no fallback.
"""
__version__ = "0.1.0"
