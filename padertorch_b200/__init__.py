"""padertorch_b200 -- B200-native (sm_100a) kernels for padertorch's speech-separation hot path:
STFT / iSTFT front-end, mask (*) spectrogram, permutation-invariant MSE, deep-clustering affinity loss
and the SI-SDR / SDR / log-MSE family, behind padertorch's own ``pt.ops`` call surface.

    import padertorch_b200 as b2s
    stft = b2s.ops.STFT(1024, 256)
    loss, perm = b2s.ops.pit_loss(est, tgt, axis=-2, return_permutation=True)
    b2s.patch_padertorch()        # make an installed padertorch (models + Trainer) use these kernels

The compute lives in ``libb200sep.so`` (C ABI: include/b200sep.h); there is no CPU fallback.
"""
from . import _lib
from . import ops
from . import review
from . import features
from .ops import STFT
from .patch import patch_padertorch, unpatch_padertorch

__version__ = '0.1.0'


def library_version():
    """Version reported by the loaded libb200sep.so (raises ImportError if it is not built)."""
    return _lib.load().b2s_version()
