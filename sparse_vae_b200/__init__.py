"""sparse_vae_b200: B200-native (sm_100a) implementation of the data-parallel hot path of norabelrose/sparse-vae --
block-sparse causal self-attention of the TransformerVAE decoder plus the latent bottleneck -- behind the
reference's own module surface.  The CUDA library (csrc/libsvae_b200.so, C ABI in include/sparse_vae_b200.h) is
mandatory: importing this package fails if it has not been built."""
from . import _native
from .core import *  # noqa: F401,F403
from .hparam_presets import hparam_presets
from .transformer_vae import TransformerVAE, TransformerVAEHparams

__version__ = '0.1.0'
