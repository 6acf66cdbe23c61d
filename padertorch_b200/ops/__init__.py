"""Mirror of the hot-path part of ``padertorch.ops`` (padertorch/ops/__init__.py:1-11): STFT, the
separation / regression losses, einsum and the packing glue, same names and call signatures, backed by libb200sep.so."""
from . import losses
from . import sequence
from ._stft import STFT
from .einsum import einsum
from .linear import FusedLinear, linear  # noqa: F401
from .sequence import pack_sequence, unpack_sequence, pad_sequence, unpad_sequence  # noqa: F401
from .losses import *  # noqa: F401,F403
from .losses.source_separation import (compute_pairwise_losses, pit_loss_from_loss_matrix,  # noqa: F401
                                       register_fast_loss)
