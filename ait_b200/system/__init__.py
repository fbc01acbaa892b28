"""Mirror of the reference's `model.system` package (≡ `transformer`): the Adaptive Image Transformer."""
from .Models import Transformer  # noqa: F401
