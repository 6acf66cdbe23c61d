"""ctypes binding of libb200sep.so (C ABI: include/b200sep.h).

There is deliberately no fallback: if the shared library is missing or a kernel call fails the
caller gets an exception, never a CPU or eager-PyTorch substitute.
"""
import ctypes
import os
import threading

import torch  # noqa: F401  (loads libcudart / initialises the allocator the pointers come from)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.environ.get('B200SEP_LIBRARY', os.path.join(_HERE, 'libb200sep.so'))

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_void = ctypes.c_void_p
c_dbl = ctypes.c_double

# name -> (restype, argtypes); mirrors include/b200sep.h one to one
SIGNATURES = {
    'b2s_version': (c_int, []),
    'b2s_last_error': (ctypes.c_char_p, []),
    'b2s_stft_frames': (c_i64, [c_i64, c_int, c_int, c_int, c_int]),
    'b2s_stft_samples': (c_i64, [c_i64, c_int, c_int, c_int]),
    'b2s_stft_frame_index': (c_i64, [c_i64, c_int, c_int, c_int]),
    'b2s_stft_plan_create': (c_int, [ctypes.POINTER(c_void), c_int, c_int, c_int, c_int,
                                     ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]),
    'b2s_stft_plan_destroy': (c_int, [c_void]),
    'b2s_stft_plan_is_fast': (c_int, [c_void]),
    'b2s_stft_scratch_bytes': (c_i64, [c_void, c_i64, c_i64]),
    'b2s_stft_forward': (c_int, [c_void, c_void, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_void,
                                 c_void]),
    'b2s_stft_backward': (c_int, [c_void, c_void, c_i64, c_i64, c_int, c_i64, c_i64, c_void, c_void,
                                  c_void]),
    'b2s_istft_forward': (c_int, [c_void, c_void, c_i64, c_i64, c_int, c_i64, c_i64, c_void, c_void,
                                  c_void]),
    'b2s_istft_backward': (c_int, [c_void, c_void, c_i64, c_i64, c_i64, c_i64, c_int, c_void,
                                   c_void]),
    'b2s_mel_create': (c_int, [ctypes.POINTER(c_void), c_int, c_int, c_int, ctypes.POINTER(ctypes.c_float)]),
    'b2s_mel_destroy': (c_int, [c_void]),
    'b2s_stft_features': (c_int, [c_void, c_void, c_i64, c_i64, c_i64, c_i64, c_i64, ctypes.c_float, ctypes.c_float,
                                  c_int, ctypes.c_float, c_void, c_void, c_void]),
    'b2s_pit_workspace_bytes': (c_i64, [c_i64, c_i64, c_i64, c_int, c_int]),
    'b2s_pit_sse_forward': (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_int,
                                    c_i64, c_int, c_void, c_void, c_void, c_void, c_void]),
    'b2s_pit_sse_backward': (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_int,
                                     c_i64, c_int, c_void, c_void, c_void, c_void, c_void]),
    'b2s_pit_sse_backward_scaled': (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_int,
                                            c_i64, c_int, c_void, c_void, c_i64, c_dbl, c_void, c_void, c_void]),
    'b2s_pit_sse_forward_mean': (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_int,
                                         c_i64, c_int, c_void, c_void, c_void, c_void, c_void, c_void]),
    'b2s_pair_workspace_bytes': (c_i64, [c_i64, c_i64, c_int]),
    'b2s_pair_stats_forward': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_int, c_i64, c_i64,
                                       c_void, c_void, c_void]),
    'b2s_pair_loss': (c_int, [c_void, c_void, c_i64, c_i64, c_int, c_int, c_int, c_dbl, c_int, c_int,
                              c_void, c_void, c_void]),
    'b2s_pair_loss_set': (c_int, [c_void, c_void, c_i64, c_i64, c_int, c_int, ctypes.POINTER(c_int),
                                  ctypes.POINTER(c_int), c_int, c_dbl, c_void, c_void, c_void, c_void]),
    'b2s_pair_stats_loss_set': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_int, c_i64, c_i64, c_int,
                                        ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_dbl, c_void, c_void,
                                        c_void, c_void, c_void, c_void]),
    'b2s_pair_backward': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_i64, c_int, c_i64, c_i64,
                                  c_void, c_int, c_int, c_dbl, c_int, c_int, c_void, c_void, c_i64,
                                  c_dbl, c_void, c_void]),
    'b2s_pair_loss_matrix': (c_int, [c_void, c_void, c_i64, c_i64, c_int, c_int, c_int, c_dbl, c_int, c_void,
                                     c_void]),
    'b2s_pair_matrix_backward': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_i64, c_int, c_i64, c_i64,
                                         c_void, c_int, c_int, c_dbl, c_int, c_void, c_void, c_void]),
    'b2s_assign': (c_int, [c_void, c_i64, c_int, c_int, c_int, c_void, c_void, c_void]),
    'b2s_tf32_split': (c_int, [c_void, c_i64, c_void, c_void]),
    'b2s_linear_forward': (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_i64, c_i64, c_i64, c_int,
                                   c_void, c_void, c_void]),
    'b2s_dc_workspace_bytes': (c_i64, [c_i64, c_i64, c_i64, c_int]),
    'b2s_dc_forward': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_i64, c_int, c_int,
                               ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_void, c_void, c_void,
                               c_void]),
    'b2s_dc_backward': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_i64, c_int, c_int,
                                ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_void, c_void, c_void,
                                c_void]),
    'b2s_dc_forward_mean': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_i64, c_int, c_int,
                                    ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_void, c_void, c_void, c_void,
                                    c_void]),
    'b2s_dc_backward_scaled': (c_int, [c_void, c_void, c_void, c_i64, c_i64, c_i64, c_int, c_int,
                                       ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_void, c_void, c_i64, c_dbl,
                                       c_void, c_void]),
    'b2s_mask_spectrum': (c_int, [c_void, c_void, c_i64, c_int, c_i64, c_i64, c_void, c_void]),
    'b2s_stft_pit_targets': (c_int, [c_void, c_void, c_void, c_void, c_i64, c_i64, c_int, c_i64, c_i64, c_void,
                                     c_void, c_void, c_void]),
    'b2s_pit_targets': (c_int, [c_void, c_void, c_i64, c_int, c_i64, c_i64, c_void, c_void, c_void, c_void]),
    'b2s_stft_pit_backward': (c_int, [c_void, c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_int, c_i64,
                                      c_i64, c_void, c_void, c_void, c_void]),
    'b2s_stft_pit_workspace_bytes': (c_i64, [c_i64, c_i64, c_int]),
    'b2s_stft_pit_forward': (c_int, [c_void, c_void, c_void, c_void, c_void, c_void, c_i64, c_i64,
                                     c_int, c_i64, c_i64, c_void, c_void, c_void, c_void, c_void]),
}

# constants of include/b200sep.h
SPEC_INTERLEAVED, SPEC_CONCAT, SPEC_ABS, SPEC_LOG1P_ABS = 0, 1, 2, 3
PIT_META, PAIR_META, DC_META = 6, 3, 4
MAX_SOURCES, DC_MAX_CHANNELS = 8, 64
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
LOSS_MSE, LOSS_LOG_MSE, LOSS_LOG1P_MSE, LOSS_SDR, LOSS_SI_SDR, LOSS_SA_SDR = range(6)
FLAG_OFFSET_INVARIANT, FLAG_GRAD_STOP = 1, 2
REDUCE_NONE, REDUCE_SUM, REDUCE_MEAN = 0, 1, 2

_lock = threading.Lock()
_lib = None


class B200SepError(RuntimeError):
    pass


def load():
    """Load libb200sep.so once.  Raises ImportError with build instructions when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIBRARY_PATH):
            raise ImportError(
                f'{LIBRARY_PATH} is missing: build the CUDA library with '
                f'`python -m padertorch_b200.build` (needs nvcc; there is no CPU fallback).')
        lib = ctypes.CDLL(LIBRARY_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error():
    return load().b2s_last_error().decode('utf-8', 'replace')


def check(rc, what):
    if rc != 0:
        message = last_error()
        if rc == -1:
            raise ValueError(f'{what}: {message}')
        raise B200SepError(f'{what}: {message} (code {rc})')


def ptr(tensor):
    """Device pointer of a tensor (None -> NULL)."""
    return None if tensor is None else tensor.data_ptr()


def stream_of(device):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_float(tensor, name):
    """The product path has no CPU implementation: fail loudly, as the task demands."""
    if not isinstance(tensor, torch.Tensor):
        raise TypeError(f'{name} must be a torch.Tensor, not {type(tensor)}')
    if not tensor.is_cuda:
        raise RuntimeError(
            f'{name} lives on {tensor.device}: padertorch_b200 runs on CUDA devices only and has '
            f'no CPU fallback (use the reference padertorch ops for CPU tensors).')
    if tensor.dtype != torch.float32:
        raise TypeError(f'{name} must be float32 (got {tensor.dtype}); the sm_100a kernels '
                        f'compute in fp32.')
    return tensor
