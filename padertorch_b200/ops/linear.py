"""Dense projections of the mask networks on the tcgen05 tensor cores (b2s_linear_forward, csrc/gemm_umma.cu):
``linear(x, weight, bias, activation)`` == ``activation(torch.nn.functional.linear(x, weight, bias))`` for
the packed-sequence GEMMs of pit/model.py:96-102 (Linear(1200, 1200) + ReLU, Linear(1200, F K) + sigmoid),
the LSTM input projections and the 1 x 1 convolutions of modules/convnet.py:120-167.

precision='fp32' (default): 3-term TF32 split, fp32-faithful (the reference runs these GEMMs in fp32);
precision='tf32': one TF32 product per term (~1e-3 relative), what ``torch.backends.cuda.matmul.allow_tf32``
would give the reference.  The backward pass uses torch.matmul (cuBLAS): plain library GEMMs.
"""
import torch

from .. import _lib

_ACTIVATIONS = {None: _lib.ACT_NONE, 'identity': _lib.ACT_NONE, 'relu': _lib.ACT_RELU, 'sigmoid': _lib.ACT_SIGMOID}
_lo_cache = {}


def tf32_split(x):
    """x - tf32(x): the second operand of the 3-term split (b2s_tf32_split)."""
    lib = _lib.load()
    x = x.contiguous()
    lo = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.b2s_tf32_split(_lib.ptr(x), x.numel(), _lib.ptr(lo), _lib.stream_of(x.device))
    _lib.check(rc, 'b2s_tf32_split')
    return lo


def _weight_lo(weight):
    """The lo part of a weight is recomputed only when the weight changes (optimizer step)."""
    key = (weight.data_ptr(), tuple(weight.shape), weight.device.index)
    hit = _lo_cache.get(key)
    if hit is not None and hit[0] == weight._version:
        return hit[1]
    if len(_lo_cache) > 64:
        _lo_cache.clear()
    lo = tf32_split(weight.detach())
    _lo_cache[key] = (weight._version, lo)
    return lo


def _pad_k(t, k_pad):
    if t.shape[-1] == k_pad:
        return t
    out = t.new_zeros(*t.shape[:-1], k_pad)
    out[..., :t.shape[-1]] = t
    return out


def linear_forward(x2d, weight, bias, activation, precision='fp32', x_lo=None, want_lo=False):
    """[M, K] x [N, K]^T on the tensor cores.  Returns (y [M, N], y_lo or None)."""
    lib = _lib.load()
    m, k = x2d.shape
    n = weight.shape[0]
    assert weight.shape[1] == k, (x2d.shape, weight.shape)
    k_pad = (k + 3) // 4 * 4          # TMA wants 16-byte row pitches
    a = _pad_k(x2d.contiguous(), k_pad)
    w = _pad_k(weight.detach().contiguous(), k_pad)
    a_lo = w_lo = None
    if precision == 'fp32':
        a_lo = _pad_k(x_lo.contiguous(), k_pad) if x_lo is not None else tf32_split(a)
        w_lo = _weight_lo(weight) if k_pad == k else tf32_split(w)
    elif precision != 'tf32':
        raise ValueError(f"precision must be 'fp32' or 'tf32', not {precision!r}")
    y = torch.empty((m, n), dtype=torch.float32, device=x2d.device)
    y_lo = torch.empty_like(y) if want_lo else None
    b = bias.detach().contiguous() if bias is not None else None
    with torch.cuda.device(x2d.device):
        rc = lib.b2s_linear_forward(_lib.ptr(a), _lib.ptr(a_lo), _lib.ptr(w), _lib.ptr(w_lo), _lib.ptr(b), m, n, k,
                                    k_pad, k_pad, _ACTIVATIONS[activation], _lib.ptr(y), _lib.ptr(y_lo),
                                    _lib.stream_of(x2d.device))
    _lib.check(rc, 'b2s_linear_forward')
    return y, y_lo


class _LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, weight, bias, activation, precision):
        y, _ = linear_forward(x2d, weight, bias, activation, precision)
        ctx.save_for_backward(x2d, weight, y)
        ctx.activation, ctx.has_bias = activation, bias is not None
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x2d, weight, y = ctx.saved_tensors
        if ctx.activation == 'relu':
            grad_y = grad_y * (y > 0)
        elif ctx.activation == 'sigmoid':
            grad_y = grad_y * y * (1 - y)
        grad_x = grad_y @ weight if ctx.needs_input_grad[0] else None
        grad_w = grad_y.t() @ x2d if ctx.needs_input_grad[1] else None
        grad_b = grad_y.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return grad_x, grad_w, grad_b, None, None


def linear(input, weight, bias=None, activation=None, precision='fp32'):
    """``activation(F.linear(input, weight, bias))`` with input [..., K], weight [N, K]."""
    _lib.require_cuda_float(input, 'input')
    _lib.require_cuda_float(weight, 'weight')
    if activation not in _ACTIVATIONS:
        raise ValueError(f'activation must be one of {sorted(map(str, _ACTIVATIONS))}')
    lead = input.shape[:-1]
    y = _LinearFunction.apply(input.reshape(-1, input.shape[-1]), weight, bias, activation, precision)
    return y.view(*lead, weight.shape[0])


class FusedLinear(torch.nn.Linear):
    """torch.nn.Linear whose forward (+ a fused ReLU / sigmoid) runs on the tcgen05 kernel for CUDA float32
    inputs; same parameters and state dict, so it can replace ``linear1`` / ``linear2`` of the PIT model."""

    def __init__(self, in_features, out_features, bias=True, activation=None, precision='fp32'):
        super().__init__(in_features, out_features, bias=bias)
        self.activation, self.precision = activation, precision

    def forward(self, input):
        if input.is_cuda and input.dtype == torch.float32:
            return linear(input, self.weight, self.bias, self.activation, self.precision)
        y = torch.nn.functional.linear(input, self.weight, self.bias)
        return torch.relu(y) if self.activation == 'relu' else (torch.sigmoid(y) if self.activation == 'sigmoid' else y)
