"""flax stand-in (see ../README.md)."""
from . import linen, serialization, training  # noqa: F401
