"""Packing glue of the mask estimators: ``padertorch.ops.sequence``'s
list <-> PackedSequence <-> padded conversions (ops/sequence/pack_module.py:29-37) and the pointwise
functions that accept a Tensor or a PackedSequence (ops/sequence/pointwise.py:20-40).  Host-side
bookkeeping over torch's own rnn utilities; no arithmetic of the hot path lives here (the fused
``log1p|Y|`` feature comes straight out of the STFT kernel: ``STFT.log1p_magnitude``)."""
import torch
from torch.nn.utils.rnn import (PackedSequence, pack_padded_sequence, pack_sequence,  # noqa: F401
                                pad_packed_sequence, pad_sequence)

__all__ = ['pack_sequence', 'unpack_sequence', 'pad_sequence', 'unpad_sequence', 'pad_packed_sequence',
           'pack_padded_sequence', 'sequence_elementwise', 'abs', 'ceil', 'clamp', 'exp', 'log', 'log1p',
           'log10', 'sigmoid', 'sqrt']


def unpad_sequence(padded_sequence: torch.Tensor, lengths):
    """[T_max, B, ...] + lengths -> list of [T_b, ...] views."""
    return [padded_sequence[:int(n), b] for b, n in enumerate(lengths)]


def unpack_sequence(packed_sequence: PackedSequence) -> list:
    """PackedSequence -> list of per-utterance tensors (the model outputs of pit/model.py:111, dc.py:73)."""
    padded, lengths = pad_packed_sequence(packed_sequence)
    return unpad_sequence(padded, lengths)


def sequence_elementwise(function, x, *args, **kwargs):
    """Apply `function` to a Tensor, or to the data of a PackedSequence keeping its batch sizes."""
    if isinstance(x, PackedSequence):
        return PackedSequence(function(x.data, *args, **kwargs), x.batch_sizes, x.sorted_indices,
                              x.unsorted_indices)
    return function(x, *args, **kwargs)


def _lift(function):
    def lifted(x, *args, **kwargs):
        return sequence_elementwise(function, x, *args, **kwargs)
    lifted.__name__ = function.__name__
    lifted.__doc__ = f'``torch.{function.__name__}`` for a Tensor or a PackedSequence.'
    return lifted


abs = _lift(torch.abs)
ceil = _lift(torch.ceil)
clamp = _lift(torch.clamp)
exp = _lift(torch.exp)
log = _lift(torch.log)
log10 = _lift(torch.log10)
log1p = _lift(torch.log1p)
sigmoid = _lift(torch.sigmoid)
sqrt = _lift(torch.sqrt)
