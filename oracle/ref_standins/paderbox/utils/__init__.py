from . import mapping, nested  # noqa: F401
