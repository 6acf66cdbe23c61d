"""Host-side mirror of ``padertorch/ops/losses/regression.py`` (same names, arguments, reductions and
assertions) over the pair-statistics kernels of libb200sep.so.  float32 CUDA tensors only."""
import torch

from ... import _lib
from . import _pairs

__all__ = ['mse_loss', 'log_mse_loss', 'sdr_loss', 'si_sdr_loss', 'log1p_mse_loss',
           'source_aggregated_sdr_loss']


def _reduce(array, reduction):
    """padertorch/ops/losses/regression.py:27-36."""
    if reduction is None or reduction == 'none':
        return array
    if reduction == 'sum':
        return torch.sum(array)
    elif reduction == 'mean':
        return torch.mean(array)
    else:
        raise ValueError(
            f'Unknown reduction: {reduction}. Choose from "sum", "mean".')


def _get_threshold(soft_sdr_max):
    """padertorch/ops/losses/regression.py:39-44; returns the kernel's tau (< 0: disabled)."""
    if soft_sdr_max is None:
        return -1.0
    assert 1 < soft_sdr_max < 50, f'Uncommon value for soft_sdr_max: {soft_sdr_max}'
    return 10 ** (-soft_sdr_max / 10)


def _real_view(estimate, target):
    """abs(e - t)^2 of complex signals (regression.py:4-18 take ``torch.abs`` first) is the squared norm of
    the difference's (re, im) view: complex inputs become real rows of twice the length."""
    if torch.is_complex(estimate) or torch.is_complex(target):
        estimate = torch.view_as_real(estimate.to(torch.complex64)).flatten(-2)
        target = torch.view_as_real(target.to(torch.complex64)).flatten(-2)
    return estimate, target


def _rowwise(estimate, target, kind, flags=0, tau=-1.0):
    _lib.require_cuda_float(estimate, 'estimate')
    _lib.require_cuda_float(target, 'target')
    _pairs.check_target(target)
    if estimate.shape != target.shape:
        estimate, target = torch.broadcast_tensors(estimate, target)
    lead = estimate.shape[:-1]
    if lead.numel() == 0:   # no rows: nothing to launch
        return estimate.new_zeros(lead) + 0 * estimate.sum()
    dense, problem = _pairs.rowwise_problem(estimate, target)
    values, _ = _pairs.PairLossFunction.apply(dense, problem, kind, flags, tau, _lib.REDUCE_NONE, False)
    return values.view(lead)


def mse_loss(estimate: torch.Tensor, target: torch.Tensor, reduction: str = 'sum'):
    """``mse_loss`` (regression.py:47-68): mean over time, `reduction` over the other axes.

    >>> mse_loss(torch.tensor([[1., 2, 3], [4, 5, 6]]).cuda(), torch.tensor([[2., 3, 4], [4, 0, 6]]).cuda())  # doctest: +SKIP
    tensor(9.3333, device='cuda:0')
    """
    complex_input = torch.is_complex(estimate) or torch.is_complex(target)
    estimate, target = _real_view(estimate, target)
    value = _rowwise(estimate, target, _lib.LOSS_MSE)
    if complex_input:   # the mean runs over T complex samples, the kernel averaged 2T real ones
        value = value * 2
    return _reduce(value, reduction=reduction)


def log_mse_loss(estimate: torch.Tensor, target: torch.Tensor, reduction: str = 'sum',
                 soft_sdr_max: float = None):
    """``log_mse_loss`` (regression.py:71-128): log10 of the time-mean squared error, optionally
    soft-thresholded with ``tau * mean(target^2)``."""
    tau = _get_threshold(soft_sdr_max) if soft_sdr_max else -1.0
    return _reduce(_rowwise(estimate, target, _lib.LOSS_LOG_MSE, tau=tau), reduction=reduction)


def sdr_loss(estimate: torch.Tensor, target: torch.Tensor, reduction: str = 'mean',
             soft_sdr_max: float = None):
    """``sdr_loss`` (regression.py:131-175): -10 log10(|t|^2 / (|e - t|^2 [+ tau |t|^2]))."""
    tau = _get_threshold(soft_sdr_max)
    estimate, target = _real_view(estimate, target)   # doctest regression.py:153-156
    # the kernel returns -10 log10(.) per row, i.e. already the negated SDR
    return _reduce(_rowwise(estimate, target, _lib.LOSS_SDR, tau=tau), reduction=reduction)


def si_sdr_loss(estimate, target, reduction='mean', offset_invariant=False,
                grad_stop=False, soft_sdr_max: float = None):
    """``si_sdr_loss`` (regression.py:178-296): sdr_loss against the optimally scaled target."""
    assert estimate.shape == target.shape, (estimate.shape, target.shape)
    assert len(estimate.shape) >= 1, estimate.shape
    assert len(estimate.shape) == 1 or estimate.shape[-2] < 10, (
        f'Number of speakers should be small (<10, not {estimate.shape[-2]})!'
    )
    flags = ((_lib.FLAG_OFFSET_INVARIANT if offset_invariant else 0)
             | (_lib.FLAG_GRAD_STOP if grad_stop else 0))
    tau = _get_threshold(soft_sdr_max)
    return _reduce(_rowwise(estimate, target, _lib.LOSS_SI_SDR, flags=flags, tau=tau),
                   reduction=reduction)


def log1p_mse_loss(estimate: torch.Tensor, target: torch.Tensor, reduction: str = 'sum'):
    """``log1p_mse_loss`` (regression.py:299-341): log10(1 + mse)."""
    return _reduce(_rowwise(estimate, target, _lib.LOSS_LOG1P_MSE), reduction=reduction)


def source_aggregated_sdr_loss(estimate: torch.Tensor, target: torch.Tensor,
                               soft_sdr_max: float = None) -> torch.Tensor:
    """``source_aggregated_sdr_loss`` (regression.py:344-376): squares of all targets and all
    errors are summed before the ratio."""
    estimate, target = _real_view(estimate, target)
    _lib.require_cuda_float(estimate, 'estimate')
    _lib.require_cuda_float(target, 'target')
    _pairs.check_target(target)
    tau = _get_threshold(soft_sdr_max)
    if estimate.shape != target.shape:   # the reference broadcasts through ``estimate - target``
        estimate, target = torch.broadcast_tensors(estimate, target)
    if estimate.numel() == 0:            # 0 / 0, as the reference's sums over nothing give
        return estimate.new_full((), float('nan'))
    dense, problem = _pairs.rowwise_problem(estimate, target)
    problem.inner = problem.groups          # one example made of every row
    value, _ = _pairs.PairLossFunction.apply(dense, problem, _lib.LOSS_SA_SDR, 0, tau,
                                             _lib.REDUCE_NONE, False)
    return value.view(())
