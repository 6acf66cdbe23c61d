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

    def forward(self, with_mean=False):
        """(loss [batch], gram) or, with_mean, (loss, gram, mean [1]): the batch mean folded by the same launch."""
        lib = _lib.load()
        device = self.emb_base.device
        c = self.e_dim + self.k
        loss = torch.empty(self.batch, dtype=torch.float32, device=device)
        gram = torch.empty((self.batch, c, c), dtype=torch.float64, device=device)
        mean = torch.empty(1, dtype=torch.float32, device=device) if with_mean else None
        if self.batch:
            nbytes = lib.b2s_dc_workspace_bytes(self.batch, self.max_frames, self.bins, c)
            ws = workspace(device, nbytes, 'dc')
            es, ts = self._strides()
            with torch.cuda.device(device):
                if with_mean:
                    rc = lib.b2s_dc_forward_mean(_lib.ptr(self.emb_base), _lib.ptr(self.tgt_base),
                                                 _lib.ptr(self.meta), self.batch, self.max_frames, self.bins,
                                                 self.e_dim, self.k, es, ts, _lib.ptr(loss), _lib.ptr(mean),
                                                 _lib.ptr(gram), _lib.ptr(ws), _lib.stream_of(device))
                else:
                    rc = lib.b2s_dc_forward(_lib.ptr(self.emb_base), _lib.ptr(self.tgt_base),
                                            _lib.ptr(self.meta), self.batch, self.max_frames, self.bins,
                                            self.e_dim, self.k, es, ts, _lib.ptr(loss), _lib.ptr(gram),
                                            _lib.ptr(ws), _lib.stream_of(device))
            _lib.check(rc, 'b2s_dc_forward')
        elif with_mean:
            mean.fill_(float('nan'))   # torch.mean of an empty batch
        return (loss, gram, mean) if with_mean else (loss, gram)

    def backward(self, gram, grad_loss, broadcast_scale=None):
        """broadcast_scale: `grad_loss` is ONE upstream value (the gradient of the batch mean) applied to every
        example times this factor."""
        lib = _lib.load()
        device = self.emb_base.device
        alloc = torch.empty if self.covers_all and self.batch else torch.zeros
        grad = alloc(self.grad_numel, dtype=torch.float32, device=device)
        if self.batch:
            es, ts = self._strides()
            grad_loss = grad_loss.to(torch.float32).contiguous()
            stride, scale = (1, 1.0) if broadcast_scale is None else (0, float(broadcast_scale))
            # the gradient buffer uses the embedding's strides; blocks start at meta's off_grad
            with torch.cuda.device(device):
                rc = lib.b2s_dc_backward_scaled(_lib.ptr(self.emb_base), _lib.ptr(self.tgt_base),
                                                _lib.ptr(self.meta), self.batch, self.max_frames, self.bins,
                                                self.e_dim, self.k, es, ts, _lib.ptr(gram),
                                                _lib.ptr(grad_loss), stride, scale, _lib.ptr(grad),
                                                _lib.stream_of(device))
            _lib.check(rc, 'b2s_dc_backward_scaled')
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


class DcMeanFunction(torch.autograd.Function):
    """Batch mean of the per-example losses (dc_loss of DeepClusteringModel.review, tcl/dc.py:83-84): the mean is folded by
    the Gram launch, the backward takes the one upstream value with the factor 1 / batch -- no reduction, expand or divide
    kernels around the two launches."""

    @staticmethod
    def forward(ctx, problem, *embeddings):
        _, gram, mean = problem.forward(with_mean=True)
        ctx.problem, ctx.gram = problem, gram
        return mean.reshape(())

    @staticmethod
    def backward(ctx, grad_mean):
        scale = 1.0 / max(ctx.problem.batch, 1)
        return (None, *ctx.problem.backward(ctx.gram, grad_mean.reshape(1), broadcast_scale=scale))


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
# loss_fn identity -> how the kernels evaluate it:
#   ('sse',)                              torch.nn.functional.mse_loss (mean over every element)
#   ('pair', kind, flags, reduction)      time-domain regression loss with its default reduction
_FAST_PIT = {
    torch.nn.functional.mse_loss: ('sse',),
    regression.mse_loss: ('pair', _lib.LOSS_MSE, 0, _lib.REDUCE_SUM),
    regression.log_mse_loss: ('pair', _lib.LOSS_LOG_MSE, 0, _lib.REDUCE_SUM),
    regression.log1p_mse_loss: ('pair', _lib.LOSS_LOG1P_MSE, 0, _lib.REDUCE_SUM),
    regression.sdr_loss: ('pair', _lib.LOSS_SDR, 0, _lib.REDUCE_MEAN),
    regression.si_sdr_loss: ('pair', _lib.LOSS_SI_SDR, 0, _lib.REDUCE_MEAN),
}
_CROSS_ENTROPY = (torch.nn.functional.cross_entropy,)


def register_fast_loss(loss_fn, like):
    """Let `loss_fn` (e.g. the reference's own ``padertorch.ops.losses.si_sdr_loss`` object) take
    the kernel path of our function `like` inside pit_loss / compute_pairwise_losses."""
    _FAST_PIT[loss_fn] = _FAST_PIT[like]


def _fast_spec(loss_fn):
    try:
        return _FAST_PIT.get(loss_fn)
    except TypeError:       # unhashable callable
        return None


def _check_pit_arguments(estimate, target, axis, loss_fn):
    """The argument contract of source_separation.py:95-110 (same messages)."""
    sources = estimate.size()[axis]
    assert sources < 30, f'Are you sure? sources={sources}, estimate.shape={estimate.shape}, target.shape={target.shape}'
    if loss_fn in _CROSS_ENTROPY:
        assert axis % estimate.ndimension() == 1, axis
        without_k = [n for d, n in enumerate(estimate.shape) if d != axis % estimate.ndimension()]
        assert without_k == list(target.shape), (
            f'{estimate.shape} (N, K, ...) does not match {target.shape} (N, ...)'
        )
    else:
        assert estimate.size() == target.size(), f'{estimate.size()} != {target.size()}'
    return sources


def _pair_rows(estimate, target, axis, spec):
    """Rows for the statistics kernel: [K, inner, length] with the permuted axis in front.
    'sse' (torch mse_loss = mean over every element): one row per source; regression losses: the last
    axis is time, the axes in between are reduced with the loss's own default reduction."""
    k = estimate.shape[axis]
    e = estimate.movedim(axis, 0)
    t = target.movedim(axis, 0)
    if spec[0] == 'sse':
        e, t = e.reshape(k, 1, -1), t.reshape(k, 1, -1)
    else:
        e, t = e.reshape(k, -1, e.shape[-1]), t.reshape(k, -1, t.shape[-1])
    e, t = e.contiguous(), t.contiguous()
    inner, length = e.shape[1], e.shape[2]
    meta = _pairs.dense_meta(inner, length, length, e.device)
    problem = _pairs.PairProblem(e, t, meta, inner, inner, k, length, inner * length, inner * length,
                                 covers_all=True)
    return e, problem


def _pair_args(spec):
    if spec[0] == 'sse':
        return _lib.LOSS_MSE, 0, _lib.REDUCE_SUM
    return spec[1], spec[2], spec[3]


def _cross_entropy_matrix(estimate, target, sources):
    """[c, k] = mean over positions of -log_softmax(estimate)[:, c] where the label is k
    (the matrix of source_separation.py:203-221; cross entropy is not on the kernel path)."""
    log_p = torch.log_softmax(estimate, dim=1).movedim(1, -1)              # [..., c]
    hit = torch.nn.functional.one_hot(target, num_classes=sources).to(estimate.dtype)   # [..., k]
    positions = target.numel()
    return -(log_p.reshape(positions, sources).t() @ hit.reshape(positions, sources)) / positions


def _first_minimum(matrix_batch, orientation=0, greedy=False):
    """Device assignment (b2s_assign) of float32 [B, K, K] cost matrices -> (int32 [B, K], float32 [B])."""
    lib = _lib.load()
    m = matrix_batch.detach().to(torch.float32).contiguous()
    batch, k = m.shape[0], m.shape[1]
    out = torch.empty((batch, k), dtype=torch.int32, device=m.device)
    value = torch.empty(batch, dtype=torch.float32, device=m.device)
    with torch.cuda.device(m.device):
        rc = lib.b2s_assign(_lib.ptr(m), batch, k, orientation, int(greedy), _lib.ptr(out), _lib.ptr(value),
                            _lib.stream_of(m.device))
    _lib.check(rc, 'b2s_assign')
    return out, value


def pit_loss(
        estimate: torch.Tensor,
        target: torch.Tensor,
        axis: int,
        loss_fn=torch.nn.functional.mse_loss,
        return_permutation: bool = False
):
    """
    Permutation invariant loss function (source_separation.py:34-124): the minimum of `loss_fn` over
    all permutations of `estimate` along `axis`; ``estimate[permutation[k]]`` is matched with
    ``target[k]``, the first minimum in ``itertools.permutations`` order wins.  Does not support a
    batch dimension or PackedSequence.

    ``torch.nn.functional.mse_loss`` and the regression losses of this package: the K x K pair terms
    are accumulated in one kernel pass and the permutations are searched on the device.
    ``cross_entropy``: pair matrix + the same device search.  Any other callable has no pairwise
    structure to exploit: it is evaluated once per permutation on the device tensors.
    """
    sources = _check_pit_arguments(estimate, target, axis, loss_fn)
    spec = _fast_spec(loss_fn)
    ndim = estimate.ndimension()

    if spec is not None and sources <= _lib.MAX_SOURCES:
        _lib.require_cuda_float(estimate, 'estimate')
        _lib.require_cuda_float(target, 'target')
        if spec[0] == 'sse':
            ax = axis % ndim
            outer = 1
            for n in estimate.shape[:ax]:
                outer *= n
            e3 = estimate.reshape(outer, sources, -1).contiguous()
            t3 = target.reshape(outer, sources, -1).contiguous()
            if e3.shape[-1] > 0 and outer > 0:
                problem = _sse.dense_problem(e3, t3)
                inputs = (e3, t3) if target.requires_grad else (e3,)
                loss, perm, _ = _sse.PitSseFunction.apply(problem, 1, *inputs)
                if return_permutation:
                    return loss[0, 0], tuple(perm[0, 0].tolist())
                return loss[0, 0]
        elif ndim >= 2:
            _pairs.check_target(target)
            dense, problem = _pair_rows(estimate, target, axis, spec)
            kind, flags, reduction = _pair_args(spec)
            loss, perm = _pairs.PairLossFunction.apply(dense, problem, kind, flags, -1.0, reduction, True)
            if return_permutation:
                return loss[0], tuple(perm[0].tolist())
            return loss[0]

    if loss_fn in _CROSS_ENTROPY and estimate.is_cuda and sources <= _lib.MAX_SOURCES:
        matrix = _cross_entropy_matrix(estimate, target, sources)            # [c, k]
        perm, _ = _first_minimum(matrix[None])
        picked = matrix[perm[0].long(), torch.arange(sources, device=matrix.device)].sum()
        if return_permutation:
            return picked, tuple(perm[0].tolist())
        return picked

    # opaque callable: one evaluation per permutation, all on the device; the scalar candidates are compared on
    # the host, which is where the first-minimum rule of a CPU torch.min (:119) is defined
    orders = list(itertools.permutations(range(sources)))
    rows = torch.tensor(orders, dtype=torch.long, device=estimate.device).reshape(len(orders), sources)
    values = [loss_fn(estimate.index_select(axis, row), target) for row in rows]
    stacked = torch.stack(values)
    winner = int(torch.min(stacked.detach().cpu(), dim=0).indices)
    if return_permutation:
        return stacked[winner], orders[winner]
    return stacked[winner]


class _PairMatrixFunction(torch.autograd.Function):
    """estimate rows -> K x K loss matrix of one example, from the pair statistics (b2s_pair_loss_matrix),
    differentiable w.r.t. the estimate (b2s_pair_matrix_backward)."""

    @staticmethod
    def forward(ctx, rows, problem, kind, flags, tau, reduction):
        lib = _lib.load()
        stats = problem.stats()
        examples = problem.groups // problem.inner
        matrix = torch.empty((examples, problem.k, problem.k), dtype=torch.float32, device=rows.device)
        with torch.cuda.device(rows.device):
            rc = lib.b2s_pair_loss_matrix(_lib.ptr(stats), _lib.ptr(problem.meta), problem.groups, problem.inner,
                                          problem.k, kind, flags, tau, reduction, _lib.ptr(matrix),
                                          _lib.stream_of(rows.device))
        _lib.check(rc, 'b2s_pair_loss_matrix')
        ctx.problem, ctx.stats, ctx.args = problem, stats, (kind, flags, tau, reduction)
        return matrix

    @staticmethod
    def backward(ctx, grad_matrix):
        lib = _lib.load()
        p = ctx.problem
        kind, flags, tau, reduction = ctx.args
        grad = torch.empty_like(p.estimate)
        g = grad_matrix.to(torch.float32).contiguous()
        with torch.cuda.device(grad.device):
            rc = lib.b2s_pair_matrix_backward(
                _lib.ptr(p.estimate), _lib.ptr(p.target), _lib.ptr(p.meta), p.groups, p.inner, p.max_length, p.k,
                p.est_stride, p.tgt_stride, _lib.ptr(ctx.stats), kind, flags, tau, reduction, _lib.ptr(g),
                _lib.ptr(grad), _lib.stream_of(grad.device))
        _lib.check(rc, 'b2s_pair_matrix_backward')
        return grad, None, None, None, None, None


def compute_pairwise_losses(
        estimate: torch.Tensor,
        target: torch.Tensor,
        axis: int,
        loss_fn=torch.nn.functional.mse_loss,
):
    """K x K matrix ``loss_fn(estimate[i], target[j])`` along `axis` (source_separation.py:127-241).

    For ``mse_loss`` and the regression losses every entry is a function of the pair statistics: one
    read of the signals yields the whole matrix, with autograd w.r.t. the estimate.  Cross entropy uses
    its closed form; any other callable is evaluated pair by pair.
    """
    sources = _check_pit_arguments(estimate, target, axis, loss_fn)
    if loss_fn in _CROSS_ENTROPY:
        assert axis == 1, axis
        return _cross_entropy_matrix(estimate, target, sources)

    spec = _fast_spec(loss_fn)
    if (spec is not None and sources <= _lib.MAX_SOURCES and estimate.is_cuda and target.is_cuda
            and estimate.dtype == torch.float32 and estimate.numel() > 0
            and not (torch.is_grad_enabled() and target.requires_grad)):
        rows, problem = _pair_rows(estimate, target, axis, spec)
        kind, flags, reduction = _pair_args(spec)
        return _PairMatrixFunction.apply(rows, problem, kind, flags, -1.0, reduction)[0]

    e_slices = estimate.unbind(axis)
    t_slices = target.unbind(axis)
    return torch.stack([torch.stack([loss_fn(e, t) for t in t_slices]) for e in e_slices])


def pit_loss_from_loss_matrix(
        pair_wise_loss_matrix,
        *,
        reduction='mean',
        algorithm='optimal',
        return_permutation=False,
):
    """PIT loss from a (K, K) pair-wise loss matrix (source_separation.py:244-312): the assignment
    ``col_ind`` minimising ``sum_i matrix[i, col_ind[i]]``.

    CUDA matrices with K <= 8 are solved on the device (b2s_assign: exhaustive, exact; the reference moves
    the matrix to the host for scipy's Hungarian solver, :285-288) and the selected entries are gathered
    with a device index, so the loss itself never leaves the GPU; larger or CPU matrices take the host
    solver like the reference.  'greedy' takes the smallest remaining entry first.
    """
    shape = tuple(pair_wise_loss_matrix.shape)
    assert len(shape) == 2, shape
    assert shape[-2] == shape[-1], shape
    sources = shape[-1]
    if algorithm not in ('optimal', 'hungarian', 'greedy', 'brute_force'):
        raise ValueError(algorithm)
    if reduction not in (None, 'mean', 'sum'):
        raise ValueError(reduction)
    greedy = algorithm == 'greedy'

    if pair_wise_loss_matrix.is_cuda and sources <= _lib.MAX_SOURCES:
        cols, _ = _first_minimum(pair_wise_loss_matrix[None], orientation=1, greedy=greedy)
        col_index = cols[0].long()
        col_ind = None
    else:
        import numpy as np
        import scipy.optimize
        host = pair_wise_loss_matrix.detach().cpu().numpy()
        if greedy:
            work = host.astype(np.float64).copy()
            col_ind = np.zeros(sources, dtype=np.int64)
            for _ in range(sources):
                i, j = np.unravel_index(np.argmin(work), work.shape)
                col_ind[i] = j
                work[i, :] = np.inf
                work[:, j] = np.inf
        else:
            col_ind = scipy.optimize.linear_sum_assignment(host)[1]
        col_index = torch.as_tensor(col_ind, dtype=torch.long, device=pair_wise_loss_matrix.device)

    rows = torch.arange(sources, device=pair_wise_loss_matrix.device)
    picked = pair_wise_loss_matrix[rows, col_index]
    min_loss = picked if reduction is None else (picked.mean() if reduction == 'mean' else picked.sum())
    if return_permutation:
        if col_ind is None:
            col_ind = col_index.cpu().numpy()
        return min_loss, col_ind
    return min_loss
