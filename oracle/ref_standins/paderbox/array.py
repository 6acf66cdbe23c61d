"""paderbox.array.segment_axis restated: strided framing of numpy arrays or torch tensors."""
import numpy as np


def _tail_pad(n, length, shift):
    # same rule as the reference's STFT tail padding (padertorch/ops/_stft.py:148-154)
    if n < length:
        return length - n
    rest = (n + shift - length) % shift
    if shift != 1 and rest != 0:
        return shift - rest
    return 0


def segment_axis(x, length, shift, axis=-1, end='cut', pad_mode='constant', pad_value=0):
    try:
        import torch
        torch_input = isinstance(x, torch.Tensor)
    except ImportError:  # pragma: no cover
        torch_input = False
    axis = axis % x.ndim
    assert end in ('pad', 'cut', None), end
    if end == 'pad':
        extra = _tail_pad(x.shape[axis], length, shift)
        if extra:
            if torch_input:
                widths = [0, 0] * (x.ndim - 1 - axis) + [0, extra]
                x = torch.nn.functional.pad(x, widths, mode=pad_mode, value=pad_value)
            else:
                widths = [(0, 0)] * x.ndim
                widths[axis] = (0, extra)
                x = np.pad(x, widths, mode=pad_mode, constant_values=pad_value)
    if torch_input:
        return x.unfold(axis, length, shift).movedim(-1, axis + 1)
    n_frames = (x.shape[axis] - length) // shift + 1
    gather = shift * np.arange(n_frames)[:, None] + np.arange(length)[None, :]
    return np.take(x, gather, axis=axis)
