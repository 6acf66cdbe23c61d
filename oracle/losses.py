"""CPU oracle: separation and regression losses of the reference (TEST INFRASTRUCTURE,
see oracle/__init__.py).  torch CPU ops, dtype follows the inputs (pass float64 tensors
for a high-precision truth)."""
import itertools

import torch


# --------------------------------------------------------------------------- regression.py helpers
def _energy(x, dim=None):
    """``_sqnorm`` (``ops/losses/regression.py:4-10``): sum |x|^2."""
    mag = torch.abs(x)
    return torch.sum(mag * mag) if dim is None else torch.sum(mag * mag, dim=dim)


def _mean_sq_err(estimate, target):
    """``_mse`` (``regression.py:13-18``): mean over the last axis of |e - t|^2."""
    err = torch.abs(estimate - target)
    return torch.mean(err * err, dim=-1)


def _apply_reduction(values, reduction):
    """``_reduce`` (``regression.py:27-36``)."""
    if reduction in (None, 'none'):
        return values
    if reduction == 'sum':
        return values.sum()
    if reduction == 'mean':
        return values.mean()
    raise ValueError(f'Unknown reduction: {reduction}. Choose from "sum", "mean".')


def _soft_threshold(soft_sdr_max):
    """``_get_threshold`` (``regression.py:39-44``): tau = 10^(-soft_sdr_max / 10)."""
    if soft_sdr_max is None:
        return None
    assert 1 < soft_sdr_max < 50, f'Uncommon value for soft_sdr_max: {soft_sdr_max}'
    return 10 ** (-soft_sdr_max / 10)


# --------------------------------------------------------------------------- regression losses
def mse_loss(estimate, target, reduction='sum'):
    """``mse_loss`` (``regression.py:47-68``); known answers 9.3333 / [1.0, 8.3333]."""
    return _apply_reduction(_mean_sq_err(estimate, target), reduction)


def log_mse_loss(estimate, target, reduction='sum', soft_sdr_max=None):
    """``log_mse_loss`` (``regression.py:71-128``); known answers 0.9208, -1.7758."""
    value = _mean_sq_err(estimate, target)
    if soft_sdr_max:
        value = value + _soft_threshold(soft_sdr_max) * torch.mean(target * target, dim=-1)
    return _apply_reduction(torch.log10(value), reduction)


def log1p_mse_loss(estimate, target, reduction='sum'):
    """``log1p_mse_loss`` (``regression.py:299-341``); known answer 1.2711."""
    return _apply_reduction(torch.log10(1 + _mean_sq_err(estimate, target)), reduction)


def sdr_loss(estimate, target, reduction='mean', soft_sdr_max=None):
    """``sdr_loss`` (``regression.py:131-175``); known answers -6.5167, -20., -11.9498."""
    signal = _energy(target, dim=-1)
    noise = _energy(estimate - target, dim=-1)
    if soft_sdr_max is not None:
        noise = noise + _soft_threshold(soft_sdr_max) * signal
    return -_apply_reduction(10 * torch.log10(signal / noise), reduction)


def si_sdr_loss(estimate, target, reduction='mean', offset_invariant=False,
                grad_stop=False, soft_sdr_max=None):
    """``si_sdr_loss`` (``regression.py:178-296``); known answers -10.7099,
    [-18.2391, -3.1806], 25.1277, -0.4811, -6.3705 and the NaN cases."""
    assert estimate.shape == target.shape, (estimate.shape, target.shape)
    assert estimate.dim() >= 1, estimate.shape
    assert estimate.dim() == 1 or estimate.shape[-2] < 10, (
        f'Number of speakers should be small (<10, not {estimate.shape[-2]})!')
    if offset_invariant:
        estimate = estimate - estimate.mean(dim=-1, keepdim=True)
        target = target - target.mean(dim=-1, keepdim=True)
    # regression.py:21-24  alpha = <e, t> / |t|^2
    alpha = (torch.sum(estimate * target, dim=-1, keepdim=True)
             / _energy(target, dim=-1).unsqueeze(-1))
    if grad_stop:
        alpha = alpha.detach()
    return sdr_loss(estimate, alpha * target, reduction=reduction, soft_sdr_max=soft_sdr_max)


def source_aggregated_sdr_loss(estimate, target, soft_sdr_max=None):
    """``source_aggregated_sdr_loss`` (``regression.py:344-376``); known answers -4.6133,
    -9.8528."""
    signal = _energy(target)
    noise = _energy(estimate - target)
    if soft_sdr_max is not None:
        noise = noise + _soft_threshold(soft_sdr_max) * signal
    return -(10 * torch.log10(signal / noise))


# --------------------------------------------------------------------------- source_separation.py
def deep_clustering_loss(x, t):
    """``deep_clustering_loss`` (``ops/losses/source_separation.py:13-31``):
    (|x^T x|_F^2 - 2 |x^T t|_F^2 + |t^T t|_F^2) / N^2 for x [N, E], t [N, K]."""
    n = x.shape[0]
    xx = x.transpose(0, 1) @ x
    xt = x.transpose(0, 1) @ t
    tt = t.transpose(0, 1) @ t
    return ((xx ** 2).sum() - 2 * (xt ** 2).sum() + (tt ** 2).sum()) / n ** 2


def pit_loss(estimate, target, axis, loss_fn=torch.nn.functional.mse_loss,
             return_permutation=False):
    """``pit_loss`` (``source_separation.py:34-124``): ``loss_fn`` for every permutation of
    the estimate along ``axis`` in ``itertools.permutations`` order, the first minimum wins
    (``torch.min`` on CPU, ``:119``).  ``estimate[perm[k]]`` is matched with ``target[k]``."""
    n_sources = estimate.shape[axis]
    assert n_sources < 30, f'Are you sure? sources={n_sources}'
    if loss_fn is torch.nn.functional.cross_entropy:
        assert axis % estimate.dim() == 1, axis
        remaining = list(estimate.shape)
        del remaining[axis]
        assert remaining == list(target.shape), (estimate.shape, target.shape)
    else:
        assert estimate.shape == target.shape, f'{estimate.shape} != {target.shape}'
    orders = list(itertools.permutations(range(n_sources)))
    candidates = torch.stack([
        loss_fn(estimate.index_select(axis, torch.tensor(order)), target) for order in orders
    ])
    best, where = torch.min(candidates, dim=0)
    if return_permutation:
        return best, orders[int(where)]
    return best


def compute_pairwise_losses(estimate, target, axis, loss_fn=torch.nn.functional.mse_loss):
    """``compute_pairwise_losses`` (``source_separation.py:127-241``), regression branch:
    entry [i, j] = loss_fn(estimate[i], target[j]) along ``axis``."""
    assert estimate.shape == target.shape, f'{estimate.shape} != {target.shape}'
    n_sources = estimate.shape[axis]
    assert n_sources < 30, f'Are you sure? sources={n_sources}'
    rows = []
    for i in range(n_sources):
        e = estimate.select(axis, i)
        rows.append(torch.stack([loss_fn(e, target.select(axis, j)) for j in range(n_sources)]))
    return torch.stack(rows)


def pit_loss_from_loss_matrix(pair_wise_loss_matrix, *, reduction='mean', algorithm='optimal',
                              return_permutation=False):
    """``pit_loss_from_loss_matrix`` (``source_separation.py:244-312``), 'optimal' branch
    (Hungarian via scipy; the 'greedy' branch needs pb_bss, an absent dependency)."""
    import scipy.optimize
    assert pair_wise_loss_matrix.dim() == 2, pair_wise_loss_matrix.shape
    assert pair_wise_loss_matrix.shape[0] == pair_wise_loss_matrix.shape[1]
    if algorithm not in ('optimal', 'hungarian'):
        raise ValueError(algorithm)
    rows, cols = scipy.optimize.linear_sum_assignment(
        pair_wise_loss_matrix.detach().cpu().numpy())
    picked = pair_wise_loss_matrix[rows, cols]
    if reduction == 'mean':
        picked = picked.mean()
    elif reduction == 'sum':
        picked = picked.sum()
    elif reduction is not None:
        raise ValueError(reduction)
    if return_permutation:
        return picked, cols
    return picked
