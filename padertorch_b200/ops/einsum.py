"""``pt.ops.einsum`` (padertorch/ops/einsum.py:10-19): einsum that accepts capital index letters and
operands as separate arguments (numpy style).  Pure index bookkeeping in front of ``torch.einsum``."""
import string

import torch

__all__ = ['einsum']


def einsum(operation: str, *operands):
    """Capital letters are renamed to lowercase letters the expression does not use yet, then the
    operands are handed to ``torch.einsum`` as a list."""
    free = [c for c in string.ascii_lowercase if c not in operation]
    capitals = sorted({c for c in operation if c in string.ascii_uppercase})
    if len(capitals) > len(free):
        raise ValueError(f'too many distinct index letters in {operation!r}')
    table = str.maketrans(dict(zip(capitals, free)))
    return torch.einsum(operation.translate(table), list(operands))
