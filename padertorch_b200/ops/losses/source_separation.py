"""Host-side mirror of ``padertorch/ops/losses/source_separation.py`` (same names, arguments and
assertions) over the PIT / deep-clustering kernels of libb200sep.so."""
import itertools

import torch
import torch.nn.functional

from ... import _lib
from ..._workspace import meta_tensor, workspace
from . import _pairs, _sse, regression

__all__ = [
    'deep_clustering_loss',
    'pit_loss',
    'compute_pairwise_losses',
    'pit_loss_from_loss_matrix',
]


# ------------------------------------------------------------------------------------------ deep clustering
class DcProblem:
    def __init__(self, emb_base, tgt_base, meta, batch, max_frames, bins, e_dim, k, emb_strides,
                 tgt_strides, grad_numel, grad_splits, keep_alive=(), covers_all=True):
        self.emb_base, self.tgt_base, self.meta = emb_base, tgt_base, meta
        self.batch, self.max_frames, self.bins, self.e_dim, self.k = batch, max_frames, bins, e_dim, k
        self.emb_strides, self.tgt_strides = emb_strides, tgt_strides
        self.grad_numel, self.grad_splits, self.keep_alive = grad_numel, grad_splits, keep_alive
        self.covers_all = covers_all

    def _strides(self):
        import ctypes
        es = (ctypes.c_int64 * 3)(*self.emb_strides)
        ts = (ctypes.c_int64 * 3)(*self.tgt_strides)
        return es, ts

    def forward(self):
        lib = _lib.load()
        device = self.emb_base.device
        c = self.e_dim + self.k
        loss = torch.empty(self.batch, dtype=torch.float32, device=device)
        gram = torch.empty((self.batch, c, c), dtype=torch.float64, device=device)
        if self.batch:
            nbytes = lib.b2s_dc_workspace_bytes(self.batch, self.max_frames, self.bins, c)
            ws = workspace(device, nbytes, 'dc')
            es, ts = self._strides()
            with torch.cuda.device(device):
                rc = lib.b2s_dc_forward(_lib.ptr(self.emb_base), _lib.ptr(self.tgt_base),
                                        _lib.ptr(self.meta), self.batch, self.max_frames, self.bins,
                                        self.e_dim, self.k, es, ts, _lib.ptr(loss), _lib.ptr(gram),
                                        _lib.ptr(ws), _lib.stream_of(device))
            _lib.check(rc, 'b2s_dc_forward')
        return loss, gram

    def backward(self, gram, grad_loss):
        lib = _lib.load()
        device = self.emb_base.device
        alloc = torch.empty if self.covers_all and self.batch else torch.zeros
        grad = alloc(self.grad_numel, dtype=torch.float32, device=device)
        if self.batch:
            es, ts = self._strides()
            grad_loss = grad_loss.to(torch.float32).contiguous()
            # the gradient buffer uses the embedding's strides; blocks start at meta's off_grad
            with torch.cuda.device(device):
                rc = lib.b2s_dc_backward(_lib.ptr(self.emb_base), _lib.ptr(self.tgt_base),
                                         _lib.ptr(self.meta), self.batch, self.max_frames, self.bins,
                                         self.e_dim, self.k, es, ts, _lib.ptr(gram),
                                         _lib.ptr(grad_loss), _lib.ptr(grad), _lib.stream_of(device))
            _lib.check(rc, 'b2s_dc_backward')
        return [grad[start:start + numel].view(shape) for start, numel, shape in self.grad_splits]


class DcFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, problem, *embeddings):
        loss, gram = problem.forward()
        ctx.problem, ctx.gram = problem, gram
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        return (None, *ctx.problem.backward(ctx.gram, grad_loss))


def _check_dc_target(t):
    if t.requires_grad:
        raise NotImplementedError('deep_clustering_loss does not propagate gradients into `t`')


def deep_clustering_loss(x, t):
    """Deep clustering loss as in Hershey 2016 paper (source_separation.py:13-31).

    Args:
        x: Shape (N, E), where it is assumed that each embedding vector
            is normalized to unit norm.
        t: Target mask with shape (N, K).
    """
    _lib.require_cuda_float(x, 'x')
    _lib.require_cuda_float(t, 't')
    _check_dc_target(t)
    assert x.dim() == 2 and t.dim() == 2 and x.shape[0] == t.shape[0], (x.shape, t.shape)
    x = x.contiguous()
    t = t.contiguous()
    n, e_dim = x.shape
    k = t.shape[1]
    if e_dim + k > _lib.DC_MAX_CHANNELS:
        raise ValueError(f'E + K = {e_dim + k} exceeds the kernel limit {_lib.DC_MAX_CHANNELS}')
    # points = N "bins" of one frame; channel stride 1, point stride E (resp. K)
    meta = meta_tensor([[1, 0, 0, 0]], x.device, cache_key=('dc-one',))
    problem = DcProblem(x, t, meta, 1, 1, n, e_dim, k, (0, 1, e_dim), (0, 1, k), x.numel(),
                        [(0, x.numel(), x.shape)])
    return DcFunction.apply(problem, x)[0]


# ------------------------------------------------------------------------------------------ PIT
# loss_fn identity -> how the kernels evaluate it inside pit_loss:
#   ('sse',)                              torch.nn.functional.mse_loss over [outer, K, inner]
#   ('pair', kind, flags, reduction)      time-domain regression loss with its default reduction
_FAST_PIT = {
    torch.nn.functional.mse_loss: ('sse',),
    regression.mse_loss: ('pair', _lib.LOSS_MSE, 0, _lib.REDUCE_SUM),
    regression.log_mse_loss: ('pair', _lib.LOSS_LOG_MSE, 0, _lib.REDUCE_SUM),
    regression.log1p_mse_loss: ('pair', _lib.LOSS_LOG1P_MSE, 0, _lib.REDUCE_SUM),
    regression.sdr_loss: ('pair', _lib.LOSS_SDR, 0, _lib.REDUCE_MEAN),
    regression.si_sdr_loss: ('pair', _lib.LOSS_SI_SDR, 0, _lib.REDUCE_MEAN),
}


def register_fast_loss(loss_fn, like):
    """Let `loss_fn` (e.g. the reference's own ``padertorch.ops.losses.si_sdr_loss`` object) take
    the kernel path of our function `like` inside pit_loss."""
    _FAST_PIT[loss_fn] = _FAST_PIT[like]


def _pair_pit_problem(estimate, target):
    """estimate / target [K, ..., T] -> groups over the middle axes."""
    k, length = estimate.shape[0], estimate.shape[-1]
    e = estimate.reshape(k, -1, length).contiguous()
    t = target.reshape(k, -1, length).contiguous()
    inner = e.shape[1]
    meta = _pairs.dense_meta(inner, length, length, e.device)
    return e, _pairs.PairProblem(e, t, meta, inner, inner, k, length, inner * length, inner * length,
                                 covers_all=True)


def pit_loss(
        estimate: torch.Tensor,
        target: torch.Tensor,
        axis: int,
        loss_fn=torch.nn.functional.mse_loss,
        return_permutation: bool = False
):
    """
    Permutation invariant loss function (source_separation.py:34-124).  Calls `loss_fn` on every
    possible permutation between `estimate`s and `target`s and returns the minimum loss among them.
    The tensors are permuted along `axis`; ``estimate[permutation[k]]`` is matched with
    ``target[k]``.  Does not support batch dimension.  Does not support PackedSequence.

    For ``torch.nn.functional.mse_loss`` and the regression losses of this package the K x K pair
    terms are accumulated in one kernel pass and the permutations are searched on the device; any
    other callable runs the reference's permutation loop on the device tensors.
    """
    sources = estimate.size()[axis]
    assert sources < 30, f'Are you sure? sources={sources}, estimate.shape={estimate.shape}, target.shape={target.shape}'

    if loss_fn in [torch.nn.functional.cross_entropy]:
        assert axis % estimate.ndimension() == 1, axis
        estimate_shape = list(estimate.shape)
        del estimate_shape[axis]
        assert estimate_shape == list(target.shape), (
            f'{estimate.shape} (N, K, ...) does not match {target.shape} (N, ...)'
        )
    else:
        assert estimate.size() == target.size(), (
            f'{estimate.size()} != {target.size()}'
        )

    fast = _FAST_PIT.get(loss_fn) if callable(loss_fn) else None
    if fast is not None and sources <= _lib.MAX_SOURCES:
        _lib.require_cuda_float(estimate, 'estimate')
        _lib.require_cuda_float(target, 'target')
        ndim = estimate.ndimension()
        if fast[0] == 'sse':
            ax = axis % ndim
            outer = 1
            for s in estimate.shape[:ax]:
                outer *= s
            e3 = estimate.reshape(outer, sources, -1).contiguous()
            t3 = target.reshape(outer, sources, -1).contiguous()
            if e3.shape[-1] > 0 and outer > 0:
                problem = _sse.dense_problem(e3, t3)
                if target.requires_grad:
                    loss, perm, _ = _sse.PitSseFunction.apply(problem, 1, e3, t3)
                else:
                    loss, perm, _ = _sse.PitSseFunction.apply(problem, 1, e3)
                min_loss = loss[0, 0]
                if return_permutation:
                    return min_loss, tuple(perm[0, 0].tolist())
                return min_loss
        elif axis % ndim == 0 and ndim >= 2:
            _pairs.check_target(target)
            dense, problem = _pair_pit_problem(estimate, target)
            _, kind, flags, reduction = fast
            loss, perm = _pairs.PairLossFunction.apply(dense, problem, kind, flags, -1.0, reduction, True)
            min_loss = loss[0]
            if return_permutation:
                return min_loss, tuple(perm[0].tolist())
            return min_loss

    # generic path: the reference algorithm on the caller's (device) tensors
    candidates = []
    indexer = [slice(None), ] * estimate.ndim
    permutations = list(itertools.permutations(range(sources)))
    for permutation in permutations:
        indexer[axis] = permutation
        candidates.append(loss_fn(
            estimate[tuple(indexer)],
            target
        ))
    min_loss, idx = torch.min(torch.stack(candidates), dim=0)

    if return_permutation:
        return min_loss, permutations[int(idx)]
    else:
        return min_loss


def compute_pairwise_losses(
        estimate: torch.Tensor,
        target: torch.Tensor,
        axis: int,
        loss_fn=torch.nn.functional.mse_loss,
):
    """K x K matrix ``loss_fn(estimate[i], target[j])`` along `axis` (source_separation.py:127-241).
    For ``mse_loss`` without autograd the matrix comes from one pass of the SSE kernel."""
    sources = estimate.size()[axis]
    assert sources < 30, f'Are you sure? sources={sources}'
    if loss_fn in [torch.nn.functional.cross_entropy]:
        import einops

        assert axis % estimate.ndimension() == 1, axis
        estimate_shape = list(estimate.shape)
        del estimate_shape[1]
        assert estimate_shape == list(target.shape), (
            f'{estimate.shape} (N, K, ...) does not match {target.shape} (N, ...)'
        )
        assert axis == 1, axis
        return einops.reduce(torch.einsum(
            'nc...,n...k->n...ck',
            -torch.nn.LogSoftmax(dim=1)(estimate),
            torch.nn.functional.one_hot(target, num_classes=sources).to(estimate.dtype)
        ), 'n ... c k -> c k', reduction='mean')

    assert estimate.size() == target.size(), (
        f'{estimate.size()} != {target.size()}'
    )
    needs_grad = torch.is_grad_enabled() and (estimate.requires_grad or target.requires_grad)
    if (_FAST_PIT.get(loss_fn) == ('sse',) and sources <= _lib.MAX_SOURCES and not needs_grad
            and estimate.is_cuda and estimate.dtype == torch.float32 and estimate.numel() > 0):
        ax = axis % estimate.ndimension()
        outer = 1
        for s in estimate.shape[:ax]:
            outer *= s
        e3 = estimate.reshape(outer, sources, -1).contiguous()
        t3 = target.reshape(outer, sources, -1).contiguous()
        _, _, sse = _sse.dense_problem(e3, t3).forward()
        return (sse[0, 0] / (outer * e3.shape[-1])).to(torch.float32)

    indexer_e = [slice(None), ] * estimate.ndim
    indexer_t = [slice(None), ] * target.ndim
    pair_wise_loss_matrix = []
    for i in range(sources):
        indexer_e[axis] = i
        for j in range(0, sources):
            indexer_t[axis] = j
            pair_wise_loss_matrix.append(loss_fn(
                estimate[tuple(indexer_e)],
                target[tuple(indexer_t)],
            ))
    return torch.stack(pair_wise_loss_matrix, 0).reshape(sources, sources)


def pit_loss_from_loss_matrix(
        pair_wise_loss_matrix,
        *,
        reduction='mean',
        algorithm='optimal',
        return_permutation=False,
):
    """PIT loss from a (K, K) pair-wise loss matrix (source_separation.py:244-312).  As in the
    reference the assignment runs on the host (scipy's Hungarian solver)."""
    import scipy.optimize

    assert len(pair_wise_loss_matrix.shape) == 2, pair_wise_loss_matrix.shape
    assert pair_wise_loss_matrix.shape[-2] == pair_wise_loss_matrix.shape[-1], pair_wise_loss_matrix.shape
    sources = pair_wise_loss_matrix.shape[-1]
    pair_wise_loss_np = pair_wise_loss_matrix.detach().cpu().numpy()

    if algorithm in ('optimal', 'hungarian'):
        row_ind, col_ind = scipy.optimize.linear_sum_assignment(pair_wise_loss_np)
    elif algorithm in ('greedy', 'brute_force'):
        from pb_bss.permutation_alignment import _mapping_from_score_matrix
        if algorithm == 'brute_force':
            algorithm = 'optimal'
        col_ind = _mapping_from_score_matrix(-pair_wise_loss_np, algorithm=algorithm)
        row_ind = range(sources)
    else:
        raise ValueError(algorithm)

    if reduction is None:
        min_loss = pair_wise_loss_matrix[row_ind, col_ind]
    elif reduction == 'mean':
        min_loss = pair_wise_loss_matrix[row_ind, col_ind].mean()
    elif reduction == 'sum':
        min_loss = pair_wise_loss_matrix[row_ind, col_ind].sum()
    else:
        raise ValueError(reduction)

    if return_permutation:
        return min_loss, col_ind
    else:
        return min_loss
