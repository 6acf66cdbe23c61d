"""Device-side scratch shared by the kernel wrappers: zero-initialised reduction workspaces (the
kernels leave their ticket counters at zero, include/b200sep.h) and small int64 `meta` tables."""
import threading

import torch

_lock = threading.Lock()
_workspaces = {}
_meta_cache = {}
_META_CACHE_LIMIT = 256


def workspace(device, nbytes, tag):
    """A zero-filled byte buffer of >= nbytes, private to (device, current stream, tag)."""
    device = torch.device(device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 4096), dtype=torch.uint8, device=device)
        with _lock:
            _workspaces[key] = buf
    return buf


def meta_tensor(rows, device, cache_key=None):
    """int64 [len(rows), width] table on `device`, cached: by `cache_key` for tables that depend only on shapes,
    by content otherwise (tables of pointer differences repeat as long as the caching allocator hands the same
    blocks back, which is the steady state of a training loop -- and what makes the calls CUDA-graph capturable:
    an upload from pageable memory is not)."""
    device = torch.device(device)
    key = (device.index, cache_key if cache_key is not None else ('rows', tuple(tuple(int(v) for v in r) for r in rows)))
    hit = _meta_cache.get(key)
    if hit is not None:
        return hit
    table = torch.tensor(rows, dtype=torch.int64).to(device)
    with _lock:
        if len(_meta_cache) >= _META_CACHE_LIMIT:
            _meta_cache.clear()
        _meta_cache[key] = table
    return table
