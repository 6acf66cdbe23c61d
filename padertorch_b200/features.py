"""Feature epilogues fused behind the STFT kernel: the call surface of
``padertorch/contrib/mk/modules/features/timefreq.py`` -- ``to_spectrogram`` (:171-183), ``Logarithm``
(:37-77) and ``MelTransform`` (:256-477) -- computed inside ``b2s_stft_features``: waveform in, (log-)
power / mel spectrogram out, the complex spectrum (and, with a filterbank, the linear spectrogram) never
reaching HBM.

``get_fbanks`` is paderbox's (``paderbox.transform.module_fbank.get_fbanks``), which is NOT part of the
reference tree: it is restated here from its published behaviour (mel-spaced triangular filters evaluated
at the FFT bin frequencies) and is PARITY UNPINNED -- what is pinned is the kernel against ``torch.matmul``
with the same basis, the filterbank's shape and its partition of unity between the first and the last
centre frequency (tests/test_features_cpu.py).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops._stft import STFT

LOG_KINDS = {False: 0, None: 1, 'e': 1, 10: 2, 10.0: 2, 2: 3, 2.0: 3}


def _log_kind(log_base):
    try:
        return LOG_KINDS[log_base]
    except (KeyError, TypeError):
        raise ValueError(f'log_base {log_base} is not supported')   # timefreq.py:66


# ------------------------------------------------------------------------------------------ filterbank
def hz2mel(frequency, htk_mel=False):
    frequency = np.asarray(frequency, dtype=np.float64)
    if htk_mel:
        return 2595.0 * np.log10(1.0 + frequency / 700.0)
    # Slaney (Auditory Toolbox): linear below 1 kHz, logarithmic above
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    linear = frequency / f_sp
    with np.errstate(divide='ignore', invalid='ignore'):
        logarithmic = min_log_mel + np.log(np.maximum(frequency, 1e-300) / min_log_hz) / logstep
    return np.where(frequency >= min_log_hz, logarithmic, linear)


def mel2hz(mel, htk_mel=False):
    mel = np.asarray(mel, dtype=np.float64)
    if htk_mel:
        return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(mel >= min_log_mel, min_log_hz * np.exp(logstep * (mel - min_log_mel)), f_sp * mel)


def mel_centre_frequencies(number_of_filters, lowest_frequency, highest_frequency, htk_mel=False):
    """number_of_filters + 2 frequencies, equally spaced on the mel scale."""
    points = np.linspace(hz2mel(lowest_frequency, htk_mel), hz2mel(highest_frequency, htk_mel),
                         number_of_filters + 2)
    return mel2hz(points, htk_mel)


def get_fbanks(sample_rate, stft_size, number_of_filters, lowest_frequency=0., highest_frequency=None,
               htk_mel=True):
    """[number_of_filters, stft_size // 2 + 1] triangular filters (PARITY UNPINNED restatement, see the
    module docstring): filter i rises from centre i to centre i + 1 and falls to centre i + 2."""
    if highest_frequency is None:
        highest_frequency = sample_rate / 2
    centres = mel_centre_frequencies(number_of_filters, lowest_frequency, highest_frequency, htk_mel)
    bins = np.arange(stft_size // 2 + 1, dtype=np.float64) * sample_rate / stft_size
    rising = (bins[None, :] - centres[:-2, None]) / (centres[1:-1] - centres[:-2])[:, None]
    falling = (centres[2:, None] - bins[None, :]) / (centres[2:] - centres[1:-1])[:, None]
    return np.maximum(0.0, np.minimum(rising, falling)).astype(np.float32)


def slaney_normalize(fbanks, sample_rate, stft_size, lowest_frequency, highest_frequency, htk_mel=False):
    """Area normalisation of ``MelTransform._normalize`` (timefreq.py:372-396, librosa's 'slaney' norm)."""
    centres = mel_centre_frequencies(fbanks.shape[0], lowest_frequency, highest_frequency, htk_mel)
    return (fbanks * (2.0 / (centres[2:] - centres[:-2]))[:, None]).astype(np.float32)


class MelFilterbank:
    """Device handle of a [bins, filters] basis (b2s_mel_create); per-device, created lazily."""

    def __init__(self, basis):
        basis = np.ascontiguousarray(np.asarray(basis, dtype=np.float32))
        assert basis.ndim == 2, basis.shape
        self.basis = basis
        self.bins, self.filters = basis.shape
        self._handles = {}

    def handle(self, device):
        index = torch.device(device).index or 0
        if index not in self._handles:
            lib = _lib.load()
            handle = ctypes.c_void_p()
            rc = lib.b2s_mel_create(ctypes.byref(handle), index, self.bins, self.filters,
                                    self.basis.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
            _lib.check(rc, 'b2s_mel_create')
            self._handles[index] = handle
        return self._handles[index]

    def __del__(self):
        try:
            lib = _lib.load()
            for handle in self._handles.values():
                lib.b2s_mel_destroy(handle)
            self._handles = {}
        except Exception:
            pass


# ------------------------------------------------------------------------------------------ kernels
def stft_features(stft, inputs, power=1.0, scale_spec=False, log_base=False, eps=1e-5, mel=None):
    """``log_b(max(eps, (|STFT(x)|^power [/ size]) [@ mel_basis]))`` in one kernel.

    inputs [..., T] float32 CUDA; returns [..., frames, F] or [..., frames, filters].  Forward only (the
    features of a separation front-end are data)."""
    lib = _lib.load()
    _lib.require_cuda_float(inputs, 'inputs')
    shape = inputs.shape
    x = inputs.detach().reshape(-1, shape[-1])
    if x.stride(-1) != 1:
        x = x.contiguous()
    frames, pad_left = stft._frames_of_call(shape[-1])
    plan = stft._plan(x.device)
    if not (plan.fast and stft.window_length == 1024 and stft.shift % 4 == 0 and stft.shift <= 1024):
        raise NotImplementedError(
            'the fused feature epilogue exists for STFT(size=1024, window_length=1024, shift % 4 == 0); use '
            'STFT.__call__ and torch ops for other geometries')
    width = stft.size // 2 + 1 if mel is None else mel.filters
    out = torch.empty((x.shape[0], frames, width), dtype=torch.float32, device=x.device)
    if mel is not None:
        assert mel.bins == stft.size // 2 + 1, (mel.bins, stft.size)
    with torch.cuda.device(x.device):
        rc = lib.b2s_stft_features(plan.handle, _lib.ptr(x), x.shape[0], x.shape[1], x.stride(0), pad_left,
                                   frames, float(power), 1.0 / stft.size if scale_spec else 1.0,
                                   _log_kind(log_base), float(eps),
                                   mel.handle(x.device) if mel is not None else None, _lib.ptr(out),
                                   _lib.stream_of(x.device))
    _lib.check(rc, 'b2s_stft_features')
    return out.view(*shape[:-1], frames, width)


class Logarithm:
    """``Logarithm`` (timefreq.py:37-77) as a description of the kernel's log stage: ``log_base`` False
    disables it, None / 'e' natural, 10, 2; ``eps`` clamps the argument."""

    def __init__(self, log_base=10, eps=1e-5):
        self.log_base, self.eps = log_base, eps
        self.kind = _log_kind(log_base)

    def inverse(self, x):
        if self.kind == 0:
            return x
        return torch.exp(x) if self.kind == 1 else torch.pow(10.0 if self.kind == 2 else 2.0, x)


class SpectrogramSTFT(STFT):
    """The feature STFT of timefreq.py:80-254 for ``spectrogram=True``: ``__call__(inputs, sequence_lengths)``
    returns ``(features, frames)`` with ``features = log(max(eps, |Y|^power [/ size]))`` as
    (batch, ..., bins, time) if ``sequence_last`` else (batch, ..., time, bins)."""

    def __init__(self, size=1024, shift=256, *, power=1., scale_spec=False, log_base=10, eps=1e-5,
                 sequence_last=True, **kwargs):
        super().__init__(size, shift, **kwargs)
        self.spectrogram = True
        self.power, self.scale_spec, self.sequence_last = power, scale_spec, sequence_last
        self.log = Logarithm(log_base, eps)

    def to_spectrogram(self, inputs):
        """``|STFT(x)|^power [/ size]`` straight from the waveform (timefreq.py:171-183)."""
        return stft_features(self, inputs, self.power, self.scale_spec, False)

    def __call__(self, inputs, sequence_lengths=None):
        encoded = stft_features(self, inputs, self.power, self.scale_spec, self.log.log_base, self.log.eps)
        if sequence_lengths is not None:
            sequence_lengths = self.samples_to_frames(np.asarray(sequence_lengths))
        if self.sequence_last:
            encoded = encoded.transpose(-2, -1)
        return encoded, sequence_lengths


class MelTransform:
    """``MelTransform`` (timefreq.py:256-477) with ``stft`` given: waveform -> log-mel spectrogram in one
    kernel.  ``forward(x, sequence_lengths)`` returns ``(features, frames)``, features of shape
    (batch, ..., number_of_filters, time) if ``sequence_last`` else (batch, ..., time, number_of_filters)."""

    def __init__(self, sampling_rate, stft_size, stft=None, number_of_filters=80, lowest_frequency=80,
                 highest_frequency=7600, htk_mel=False, norm='slaney', log_base=10, eps=1e-5,
                 sequence_last=True, power=1.0, scale_spec=False, mel_basis=None):
        self.sampling_rate, self.stft_size = sampling_rate, stft_size
        self.stft = stft if stft is not None else STFT(stft_size, stft_size // 4)
        assert self.stft.size == stft_size, (self.stft.size, stft_size)
        self.number_of_filters = number_of_filters
        self.lowest_frequency = lowest_frequency
        self.highest_frequency = sampling_rate // 2 if highest_frequency is None else highest_frequency
        self.htk_mel, self.norm, self.sequence_last = htk_mel, norm, sequence_last
        self.power, self.scale_spec = power, scale_spec
        if mel_basis is None:   # [bins, filters] like the reference's parameter (timefreq.py:338)
            fbanks = get_fbanks(sampling_rate, stft_size, number_of_filters, lowest_frequency,
                                self.highest_frequency, htk_mel)
            if norm == 'slaney':
                fbanks = slaney_normalize(fbanks, sampling_rate, stft_size, lowest_frequency,
                                          self.highest_frequency, htk_mel)
            elif norm is not None:
                raise ValueError(f'Unknown norm: {norm}')
            mel_basis = fbanks.T
        self.mel_basis = np.ascontiguousarray(np.asarray(mel_basis, dtype=np.float32))
        self.filterbank = MelFilterbank(self.mel_basis)
        self.log = Logarithm(log_base, eps)

    def forward(self, x, sequence_lengths=None):
        features = stft_features(self.stft, x, self.power, self.scale_spec, self.log.log_base, self.log.eps,
                                 mel=self.filterbank)
        if sequence_lengths is not None:
            sequence_lengths = self.stft.samples_to_frames(np.asarray(sequence_lengths))
        if self.sequence_last:
            features = features.transpose(-2, -1)
        return features, sequence_lengths

    __call__ = forward
