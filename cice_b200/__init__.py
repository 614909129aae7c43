"""cice_b200 -- B200-native EVP sea-ice dynamics subcycling path behind a C ABI.

Only what the path needs: csrc/ (CUDA kernels + the C ABI), the host-side mirror of the
reference seam (dyn_evp), the block-decomposition data contract (decomp) and the synthetic
box2001 input generator (synth).
"""
from . import abi, decomp, synth  # noqa: F401

__all__ = ["abi", "decomp", "synth", "dyn_evp"]
