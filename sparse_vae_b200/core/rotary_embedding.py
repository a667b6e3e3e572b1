"""Stand-in for a module the reference imports but does not ship (core/rotary_embedding.py is missing from
its tree; only `RotaryEmbedding.embedding_context` is ever used: core/perceiver.py:15,
core/transformer_language_model.py:65).  The rotation the model actually applies is
`encode_position_rotary` in the attention module."""
from contextlib import contextmanager


class RotaryEmbedding:
    current_embedding = None

    @staticmethod
    @contextmanager
    def embedding_context(d_model: int):
        yield
