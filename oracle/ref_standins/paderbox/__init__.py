"""Minimal paderbox subset: only what `import padertorch` + the separation hot path touch."""
from . import array, io, transform, utils  # noqa: F401
