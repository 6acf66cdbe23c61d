"""Mirror of the hot-path part of ``padertorch.ops`` (padertorch/ops/__init__.py:1-11): STFT and the
separation / regression losses, same names and call signatures, backed by libb200sep.so."""
from . import losses
from ._stft import STFT
from .losses import *  # noqa: F401,F403
from .losses.source_separation import (compute_pairwise_losses, pit_loss_from_loss_matrix,  # noqa: F401
                                       register_fast_loss)
