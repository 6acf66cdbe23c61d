"""CPU oracle: the callers either side of the ops -- the ``review`` loops of the three
hot-path models and the STFT->mask->PIT step the bench measures (TEST INFRASTRUCTURE,
see oracle/__init__.py)."""
import torch

from . import losses
from .stft import ReferenceSTFT


def pit_review_losses(masks, y_abs, x_abs, cos_phase_difference=None,
                      return_permutation=False):
    """``PermutationInvariantTrainingModel.review`` (``contrib/examples/source_separation/
    pit/model.py:112-140``): per example ``pit_loss(mask * Y_abs[:, None, :], target,
    axis=-2)`` for the MSE target and the ideal-phase-sensitive target; batch means.

    masks: list of [T_b, K, F]; y_abs: list of [T_b, F]; x_abs / cos_phase_difference:
    lists of [T_b, K, F].  Returns dict of 0-dim tensors (+ per-example values / perms)."""
    mse, ips, perms = [], [], []
    for b, (mask, observation, target) in enumerate(zip(masks, y_abs, x_abs)):
        estimate = mask * observation[:, None, :]
        value, perm = losses.pit_loss(estimate, target, axis=-2, return_permutation=True)
        mse.append(value)
        perms.append(perm)
        if cos_phase_difference is not None:
            ips.append(losses.pit_loss(estimate, target * cos_phase_difference[b], axis=-2))
    out = {'pit_mse_loss': torch.mean(torch.stack(mse))}
    if ips:
        out['pit_ips_loss'] = torch.mean(torch.stack(ips))
    if return_permutation:
        out['per_example_mse'] = torch.stack(mse)
        out['permutations'] = perms
    return out


def dc_review_loss(embeddings, target_masks):
    """``DeepClusteringModel.review`` (``contrib/tcl/dc.py:76-84``): per example
    ``deep_clustering_loss('t e f -> (t f) e', 't k f -> (t f) k')``; batch mean."""
    values = []
    for embedding, target_mask in zip(embeddings, target_masks):
        e_dim, k_dim = embedding.shape[1], target_mask.shape[1]
        x = embedding.permute(0, 2, 1).reshape(-1, e_dim)
        t = target_mask.permute(0, 2, 1).reshape(-1, k_dim)
        values.append(losses.deep_clustering_loss(x, t))
    return torch.mean(torch.stack(values)), torch.stack(values)


_TASNET_LOSSES = {
    'si-sdr': losses.si_sdr_loss,
    'log-mse': losses.log_mse_loss,
    'log1p-mse': losses.log1p_mse_loss,
}


def tasnet_losses(estimates, targets, num_samples):
    """``TasNet.loss`` (``contrib/examples/source_separation/tasnet/model.py:154-176``):
    per example crop to its length, ``pit_loss(axis=0)`` for si-sdr / log-mse / log1p-mse;
    batch mean of each.  estimates / targets: [B, K, T]."""
    collected = {name: [] for name in _TASNET_LOSSES}
    for length, estimated, target in zip(num_samples, estimates, targets):
        for name, fn in _TASNET_LOSSES.items():
            collected[name].append(losses.pit_loss(
                estimated[..., :length], target[..., :length], axis=0, loss_fn=fn))
    return {name: torch.mean(torch.stack(v)) for name, v in collected.items()}


def prepare_pit_example(y, s, stft):
    """Feature / target preparation of the PIT example (``contrib/examples/
    source_separation/pit/data.py:49-77``): Y = stft(y) [T, F], X = stft(s) as [T, K, F];
    returns Y_abs, X_abs, cos(angle(Y) - angle(X))."""
    Y = stft(y)
    X = stft(s).transpose(0, 1)
    cos_phase_difference = torch.cos(torch.angle(Y[:, None, :]) - torch.angle(X))
    return Y.abs(), X.abs(), cos_phase_difference


def stft_mask_pit_step(y, s, masks, size=1024, shift=256, stft=None):
    """The step ``bench.py`` measures (BASELINE.json metric): for a
    batch of mixtures y [B, T] with sources s [B, K, T] and mask-network outputs
    masks [B, M, K, F]:  |Y| = |STFT(y)| (front-end feature),  X_k = |STFT(s_k)| (targets),
    then per example ``pit_loss(mask * |Y|[:, None, :], X, axis=-2)``
    (``pit/model.py:117-128``).  Returns (per-example loss [B], permutations [B][K], |Y|)."""
    if stft is None:
        stft = ReferenceSTFT(size, shift)
    y_abs = stft(y).abs()                                  # [B, M, F]
    x_abs = stft(s).abs().transpose(1, 2)                  # [B, M, K, F]
    values, perms = [], []
    for b in range(y.shape[0]):
        value, perm = losses.pit_loss(masks[b] * y_abs[b][:, None, :], x_abs[b], axis=-2,
                                      return_permutation=True)
        values.append(value)
        perms.append(perm)
    return torch.stack(values), perms, y_abs
