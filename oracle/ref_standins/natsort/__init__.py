import re


def _natural_key(value):
    if isinstance(value, (tuple, list)):
        return tuple(_natural_key(v) for v in value)
    if isinstance(value, str):
        return tuple(int(tok) if tok.isdigit() else tok for tok in re.split(r'(\d+)', value))
    return value


def natsorted(seq, key=None, reverse=False):
    if key is None:
        return sorted(seq, key=_natural_key, reverse=reverse)
    return sorted(seq, key=lambda item: _natural_key(key(item)), reverse=reverse)
