"""CPU oracle: STFT / iSTFT of the reference (TEST INFRASTRUCTURE, see oracle/__init__.py).

Two independent formulations are kept so that each checks the other:

* ``ReferenceSTFT`` follows ``padertorch/ops/_stft.py`` operation by operation (windowed
  DFT matrix applied as a strided 1-d convolution, transposed convolution for the
  inverse) with torch CPU ops.  This is the "port" that ``bench.py`` times as the CPU
  baseline because it has the reference's cost model (dense DFT GEMM).
* ``stft_rfft`` / ``istft_rfft`` follow the numpy STFT of paderbox (frame, window,
  ``rfft``), which the reference's tests use as *their* oracle
  (``tests/test_ops/test_stft.py:72-96``).  Used in float64 as the high-precision truth.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

_LEGAL_FADING = (None, True, False, 'full', 'half')
_LEGAL_REPRESENTATIONS = ('concat', 'stacked', 'complex')


# --------------------------------------------------------------------------- windows
def get_window(window='blackman', symmetric_window=False, window_length=1024):
    """paderbox ``_get_window`` (called at ``padertorch/ops/_stft.py:91-95``).

    A name resolves to ``scipy.signal.windows.<name>``; the periodic ("DFT-even") form
    ``w(L + 1)[:-1]`` is used unless ``symmetric_window``.  Pinned by the literal hann
    vector at ``padertorch/contrib/cb/transform.py:219-232``.
    """
    if isinstance(window, str):
        from scipy.signal import windows
        window = getattr(windows, window)
    if callable(window):
        if symmetric_window:
            return np.asarray(window(window_length), dtype=np.float64)
        return np.asarray(window(window_length + 1), dtype=np.float64)[:-1]
    window = np.asarray(window, dtype=np.float64)
    assert window.shape == (window_length,), (window.shape, window_length)
    return window


def biorthogonal_window(analysis_window, shift):
    """paderbox ``_biorthogonal_window_fastest`` (called at ``ops/_stft.py:27-28``).

    ``ws[n] = w[n] / sum_j w[(n mod shift) + j*shift]**2`` -- the synthesis window that
    makes windowed overlap-add the exact inverse.  Pinned by the round-trip property
    ``tests/test_ops/test_stft.py:36-42``.
    """
    w = np.asarray(analysis_window, dtype=np.float64)
    energy = np.zeros_like(w)
    for r in range(min(shift, len(w))):
        energy[r::shift] = np.sum(w[r::shift] ** 2)
    return w / energy


# --------------------------------------------------------------------------- frame arithmetic (integer, bit exact)
def _fading_extra(window_length, shift, fading):
    if fading in (None, False):
        return 0
    return (2 if fading != 'half' else 1) * (window_length - shift)


def samples_to_frames(samples, window_length, shift, pad=True, fading='full'):
    """``STFT.samples_to_frames`` (``ops/_stft.py:265-279`` -> paderbox
    ``_samples_to_stft_frames``).  Works on ints and integer arrays."""
    total = samples + _fading_extra(window_length, shift, fading)
    numerator = total - window_length + shift
    if pad:
        frames = -((-numerator) // shift)          # ceil for ints and int arrays alike
    else:
        frames = numerator // shift
    return frames


def frames_to_samples(frames, window_length, shift, fading='full'):
    """``STFT.frames_to_samples`` (``ops/_stft.py:295-307``)."""
    return frames * shift + window_length - shift - _fading_extra(window_length, shift, fading)


def sample_index_to_frame_index(sample_index, window_length, shift, fading='full'):
    """``STFT.sample_index_to_frame_index`` (``ops/_stft.py:281-293``).  PARITY UNPINNED:
    no reference test / doctest / call site exercises it; window-centre convention."""
    if fading in (None, False):
        offset = 0
    elif fading == 'half':
        offset = (window_length - shift) // 2
    else:
        offset = window_length - shift
    return np.maximum((sample_index + offset - window_length // 2) // shift, 0)


def fading_pad_widths(window_length, shift, fading):
    """Left / right zero padding of ``STFT.__call__`` (``ops/_stft.py:137-146``)."""
    assert fading in _LEGAL_FADING, fading
    if fading in (None, False):
        return 0, 0
    if fading == 'half':
        return (window_length - shift) // 2, math.ceil((window_length - shift) / 2)
    return window_length - shift, window_length - shift


def tail_pad(padded_length, window_length, shift, pad):
    """Tail zero padding of ``STFT.__call__`` (``ops/_stft.py:148-154``)."""
    if not pad:
        return 0
    if padded_length < window_length:
        return window_length - padded_length
    rest = (padded_length + shift - window_length) % shift
    if shift != 1 and rest != 0:
        return shift - rest
    return 0


# --------------------------------------------------------------------------- DFT matrices (float64)
def analysis_matrix(size, window):
    """``get_stft_kernel`` (``ops/_stft.py:11-23``): rows 0..F-1 ``cos(-2 pi n k / size) w[k]``,
    rows F..2F-1 ``sin(-2 pi n k / size) w[k]``; shape ``[2F, L]`` float64."""
    window = np.asarray(window, dtype=np.float64)
    n = np.arange(size // 2 + 1)[:, None]
    k = np.arange(len(window))[None, :]
    phase = (-1 * n * 2 * np.pi / size) * k
    return np.concatenate([np.cos(phase) * window, np.sin(phase) * window], axis=0)


def synthesis_matrices(size, shift, window):
    """``get_istft_kernel`` (``ops/_stft.py:26-43``): ``cos(2 pi f n / size) ws[n]`` and
    ``sin(-2 pi f n / size) ws[n]`` for f in 0..size-1, ``ws = biorthogonal(window) / size``."""
    ws = biorthogonal_window(window, shift) / size
    f = np.arange(size)[:, None]
    n = np.arange(len(ws))[None, :]
    real = np.cos((1 * f * 2 * np.pi / size) * n) * ws
    imag = np.sin((-1 * f * 2 * np.pi / size) * n) * ws
    return real, imag


# --------------------------------------------------------------------------- the port of ops/_stft.py
class ReferenceSTFT:
    """Operation-by-operation CPU restatement of ``padertorch.ops.STFT``
    (``ops/_stft.py:46-307``).  Attributes may be mutated after construction, as the
    reference's tests do with ``stft.fading`` (``tests/test_ops/test_stft.py:46,59``)."""

    def __init__(self, size=1024, shift=256, *, window='blackman', window_length=None,
                 fading='full', pad=True, symmetric_window=False,
                 complex_representation='complex'):
        assert complex_representation in _LEGAL_REPRESENTATIONS, complex_representation
        assert size % 2 == 0, 'only even FFT sizes'            # ops/_stft.py:85
        assert fading in _LEGAL_FADING, fading                  # ops/_stft.py:96
        self.size = size
        self.shift = shift
        self.window_length = size if window_length is None else window_length
        self.fading = fading
        self.pad = pad
        self.complex_representation = complex_representation
        self.window = get_window(window, symmetric_window, self.window_length)
        self._analysis = torch.from_numpy(analysis_matrix(size, self.window))[:, None, :]
        real, imag = synthesis_matrices(size, shift, self.window)
        self._synthesis_real = torch.from_numpy(real)[:, None, :]
        self._synthesis_imag = torch.from_numpy(imag)[:, None, :]

    # ops/_stft.py:103-174
    def __call__(self, inputs):
        inputs = torch.as_tensor(inputs)
        lead = inputs.shape[:-1]
        x = inputs.reshape(-1, inputs.shape[-1])
        left, right = fading_pad_widths(self.window_length, self.shift, self.fading)
        if left or right:
            x = F.pad(x, (left, right))
        extra = tail_pad(x.shape[-1], self.window_length, self.shift, self.pad)
        if extra:
            x = F.pad(x, (0, extra))
        spec = F.conv1d(x[:, None, :], self._analysis.to(x), stride=self.shift)   # [., 2F, M]
        spec = spec.reshape(*lead, *spec.shape[-2:]).transpose(-1, -2)               # [..., M, 2F]
        n_bins = self.size // 2 + 1
        real, imag = spec[..., :n_bins], spec[..., n_bins:]
        if self.complex_representation == 'complex':
            return torch.complex(real.contiguous(), imag.contiguous())
        if self.complex_representation == 'stacked':
            return torch.stack([real, imag], dim=-1)
        return torch.cat([real, imag], dim=-1)

    # ops/_stft.py:176-263
    def inverse(self, stft_signal):
        stft_signal = torch.as_tensor(stft_signal)
        if self.complex_representation == 'complex':
            real, imag = stft_signal.real, stft_signal.imag
        elif self.complex_representation == 'stacked':
            real, imag = stft_signal[..., 0], stft_signal[..., 1]
        else:
            n_bins = stft_signal.shape[-1] // 2
            real, imag = stft_signal[..., :n_bins], stft_signal[..., n_bins:]
        lead = real.shape[:-2]

        def overlap_add(part, kernel, sign):
            part = part.reshape(-1, *part.shape[-2:]).transpose(-1, -2)           # [., F, M]
            mirrored = part[:, 1:-1].flip(1)
            full = torch.cat([part, sign * mirrored], dim=1)                       # Hermitian half
            return F.conv_transpose1d(full, kernel.to(part), stride=self.shift)

        signal = (overlap_add(real, self._synthesis_real, 1)
                  + overlap_add(imag, self._synthesis_imag, -1))
        signal = signal.reshape(*lead, signal.shape[-1])
        if self.fading not in (None, False):
            cut = self.window_length - self.shift
            if self.fading == 'half':
                cut = cut / 2
            signal = signal[..., int(cut):signal.shape[-1] - math.ceil(cut)]
        return signal

    def samples_to_frames(self, samples):
        return samples_to_frames(samples, self.window_length, self.shift, self.pad, self.fading)

    def frames_to_samples(self, frames):
        return frames_to_samples(frames, self.window_length, self.shift, self.fading)

    def sample_index_to_frame_index(self, sample_index):
        return sample_index_to_frame_index(sample_index, self.window_length, self.shift,
                                           self.fading)


# --------------------------------------------------------------------------- rfft formulation (float64 truth)
def stft_rfft(x, size=1024, shift=256, *, window='blackman', window_length=None,
              fading='full', pad=True, symmetric_window=False):
    """paderbox numpy ``stft`` (the oracle of the reference's own tests,
    ``tests/test_ops/test_stft.py:72-78``): pad, frame, window, ``rfft``.  float64 in/out."""
    x = np.asarray(x, dtype=np.float64)
    window_length = size if window_length is None else window_length
    w = get_window(window, symmetric_window, window_length)
    left, right = fading_pad_widths(window_length, shift, fading)
    widths = [(0, 0)] * (x.ndim - 1)
    x = np.pad(x, widths + [(left, right)])
    if pad:
        x = np.pad(x, widths + [(0, tail_pad(x.shape[-1], window_length, shift, True))])
    n_frames = (x.shape[-1] - window_length) // shift + 1
    index = shift * np.arange(n_frames)[:, None] + np.arange(window_length)[None, :]
    return np.fft.rfft(x[..., index] * w, n=size, axis=-1)


def istft_rfft(spec, size=1024, shift=256, *, window='blackman', window_length=None,
               fading='full', symmetric_window=False):
    """paderbox numpy ``istft`` (``tests/test_ops/test_stft.py:89-96``): ``irfft``,
    synthesis window, overlap-add, crop the fading.  DC / Nyquist imaginary parts are
    ignored by ``irfft`` exactly as the reference's sine rows are zero there."""
    spec = np.asarray(spec, dtype=np.complex128)
    window_length = size if window_length is None else window_length
    ws = biorthogonal_window(get_window(window, symmetric_window, window_length), shift)
    pieces = np.fft.irfft(spec, n=size, axis=-1)[..., :window_length] * ws
    n_frames = spec.shape[-2]
    out = np.zeros(spec.shape[:-2] + ((n_frames - 1) * shift + window_length,))
    for m in range(n_frames):
        out[..., m * shift:m * shift + window_length] += pieces[..., m, :]
    if fading not in (None, False):
        cut = window_length - shift
        if fading == 'half':
            cut = cut / 2
        out = out[..., int(cut):out.shape[-1] - math.ceil(cut)]
    return out
