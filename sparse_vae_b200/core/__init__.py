"""Same public names as the reference's `sparse_vae.core` (core/__init__.py:1-13) for the hot path and its callers."""
from .attention import Attention, Perceiver, TransformerLayer, encode_position_rotary
from .conditional_gaussian import ConditionalGaussian, fused_bottleneck
from .continuous_autoencoder import ContinuousVAE, ContinuousVAEHparams
from .generation import GenerationState
from .language_model import (LanguageModel, LanguageModelHparams, cosine_decay, cosine_decay_with_warmup,
                             get_cosine_decay_with_warmup_schedule, robust_cross_entropy)
from .math_utils import marginal_kl
from .padded_tensor import PaddedTensor
from .rectified_adam import RAdam
from .rotary_embedding import RotaryEmbedding
from .sparse_attention import SparseAttention
from .transformer_language_model import VOCAB_SIZE, TransformerHparams, TransformerLanguageModel
