"""Mirror of `model.system.SubLayers` (parameter containers, see Models.py)."""
from .Models import MultiHeadAttention, PositionwiseFeedForward, SHBlock  # noqa: F401
