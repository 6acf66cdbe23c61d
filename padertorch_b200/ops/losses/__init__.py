"""Mirror of ``padertorch.ops.losses`` (padertorch/ops/losses/__init__.py) for the hot path."""
from . import regression
from . import source_separation
from .regression import *  # noqa: F401,F403
from .source_separation import *  # noqa: F401,F403
