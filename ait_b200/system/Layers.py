"""Mirror of `model.system.Layers` (parameter containers, see Models.py)."""
from .Models import DecoderLayer, EncoderLayer  # noqa: F401
