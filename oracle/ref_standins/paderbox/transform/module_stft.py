"""paderbox.transform.module_stft restated (numpy STFT the reference mirrors).

Pinned by: literal vector padertorch/contrib/cb/transform.py:219-232, frame counts
tests/test_ops/test_stft.py:44-70,139-165, round trip :36-42.
"""
import math

import numpy as np
from scipy import signal

from ..array import segment_axis


def _get_window(window, symmetric_window, window_length):
    if isinstance(window, str):
        window = getattr(signal.windows, window)
    if not callable(window):
        raise TypeError(window)
    if symmetric_window:
        return window(window_length)
    # DFT-even ("periodic") window, cf. scipy issue 4551
    return window(window_length + 1)[:-1]


def _biorthogonal_window_fastest(analysis_window, shift, use_amplitude=False):
    w = np.asarray(analysis_window, dtype=np.float64)
    length = len(w)
    power = w if use_amplitude else w ** 2
    denominator = np.zeros(length)
    for residue in range(min(shift, length)):
        denominator[residue::shift] = power[residue::shift].sum()
    if use_amplitude:
        return 1 / denominator
    return w / denominator


def _fading_samples(size, shift, fading):
    if fading in (None, False):
        return 0
    return (1 + (fading != 'half')) * (size - shift)


def _samples_to_stft_frames(samples, size, shift, *, pad=True, fading=None):
    samples = samples + _fading_samples(size, shift, fading)
    frames = (samples - size + shift) / shift
    rounding = np.ceil if pad else np.floor
    frames = rounding(frames)
    if isinstance(frames, np.ndarray):
        return frames.astype(int)
    return int(frames)


def _stft_frames_to_samples(frames, size, shift, fading=None):
    return frames * shift + size - shift - _fading_samples(size, shift, fading)


def sample_index_to_stft_frame_index(sample, window_length, shift, fading='full'):
    # PARITY UNPINNED: centre-of-window convention.
    if fading in (None, False):
        frame = (sample - window_length // 2) // shift
    elif fading == 'half':
        frame = (sample + (window_length - shift) // 2 - window_length // 2) // shift
    else:
        frame = (sample + (window_length - shift) - window_length // 2) // shift
    return np.maximum(frame, 0)


def _fading_widths(window_length, shift, fading):
    if fading in (None, False):
        return None
    if fading == 'half':
        return (window_length - shift) // 2, math.ceil((window_length - shift) / 2)
    return window_length - shift, window_length - shift


def stft(time_signal, size=1024, shift=256, *, axis=-1, window='blackman',
         window_length=None, fading='full', pad=True, symmetric_window=False):
    x = np.asarray(time_signal)
    assert axis % x.ndim == x.ndim - 1, 'stand-in supports the last axis only'
    if window_length is None:
        window_length = size
    widths = _fading_widths(window_length, shift, fading)
    if widths is not None:
        x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [widths])
    w = _get_window(window, symmetric_window, window_length)
    frames = segment_axis(x, window_length, shift, axis=-1, end='pad' if pad else 'cut')
    return np.fft.rfft(frames * w, n=size, axis=-1)


def istft(stft_signal, size=1024, shift=256, *, window='blackman', fading='full',
          window_length=None, symmetric_window=False, num_samples=None, pad=True,
          biorthogonal_window=None):
    if window_length is None:
        window_length = size
    if biorthogonal_window is None:
        biorthogonal_window = _biorthogonal_window_fastest(
            _get_window(window, symmetric_window, window_length), shift)
    pieces = np.fft.irfft(stft_signal, n=size, axis=-1)[..., :window_length]
    pieces = pieces * biorthogonal_window
    n_frames = pieces.shape[-2]
    out = np.zeros(pieces.shape[:-2] + (n_frames * shift + window_length - shift,),
                   dtype=pieces.dtype)
    for m in range(n_frames):
        out[..., m * shift:m * shift + window_length] += pieces[..., m, :]
    widths = _fading_widths(window_length, shift, fading)
    if widths is not None:
        cut = window_length - shift
        if fading == 'half':
            cut = cut / 2
        out = out[..., int(cut):out.shape[-1] - math.ceil(cut)]
    if num_samples is not None:
        out = out[..., :num_samples]
    return out
