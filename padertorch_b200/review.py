"""Batched replacements for the per-example loss loops of the three hot-path models -- the callers
one level above the ops:

* ``pit_review_losses``  <- PermutationInvariantTrainingModel.review
                            (padertorch/contrib/examples/source_separation/pit/model.py:112-140)
* ``dc_review_loss``     <- DeepClusteringModel.review (padertorch/contrib/tcl/dc.py:76-84)
* ``tasnet_losses``      <- TasNet.loss (padertorch/contrib/examples/source_separation/tasnet/model.py:154-176)
* ``stft_mask_pit_step`` <- the STFT -> |.| -> mask -> PIT step of BASELINE.json's metric
                            (pit/data.py:49-77 + pit/model.py:117-128), one fused kernel.

Each processes the whole (ragged) minibatch in one or two kernel launches and returns the same
values the reference loops produce: the batch mean of per-example losses.
"""
import torch

from . import _lib
from ._workspace import meta_tensor, workspace
from .ops.losses import _pairs, _sse
from .ops.losses.source_separation import DcFunction, DcMeanFunction, DcProblem, _check_dc_target
from .ops._stft import STFT


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else None


def _pit_problem(masks, observations, targets, cos_phase_difference, lengths):
    """(SseProblem, autograd inputs, dual) of a minibatch.  masks / targets / cos_phase_difference: lists of [T_b, K, F]
    tensors, or padded [B, T, K, F] tensors with `lengths`; observations: list of [T_b, F] or padded [B, T, F]."""
    dual = cos_phase_difference is not None
    mask_list = _as_list(masks)
    if mask_list is not None:
        for m in mask_list:
            _lib.require_cuda_float(m, 'mask')
        problem, inputs = _sse.list_problem(mask_list, _as_list(observations), _as_list(targets),
                                            _as_list(cos_phase_difference) if dual else None, dual)
        return problem, list(inputs), dual
    _lib.require_cuda_float(masks, 'mask')
    problem, dense = _sse.padded_problem(masks, observations, targets,
                                         cos_phase_difference if dual else None, lengths, dual)
    return problem, [dense], dual


def pit_losses_per_example(masks, observations, targets, cos_phase_difference=None, lengths=None):
    """Per-example PIT losses of a minibatch (inputs as `_pit_problem`).
    Returns (mse [B], mse_perm [B, K] int32, ips [B] or None, ips_perm or None).
    """
    problem, inputs, dual = _pit_problem(masks, observations, targets, cos_phase_difference, lengths)
    loss, perm, _ = _sse.PitSseFunction.apply(problem, len(inputs), *inputs)
    if dual:
        return loss[0], perm[0], loss[1], perm[1]
    return loss[0], perm[0], None, None


def pit_review_losses(masks, observations, targets, cos_phase_difference=None, lengths=None):
    """``dict(pit_mse_loss=..., pit_ips_loss=...)`` exactly as the ``losses`` entry of
    PermutationInvariantTrainingModel.review (pit/model.py:137-140): batch means -- both from ONE one-warp-per-loss
    launch behind the loss kernel, their gradients taken by the backward kernel directly (no ATen mean / expand /
    divide kernels)."""
    problem, inputs, dual = _pit_problem(masks, observations, targets, cos_phase_difference, lengths)
    if problem.batch == 0:
        loss, _, _ = _sse.PitSseFunction.apply(problem, len(inputs), *inputs)
        means = loss.mean(dim=1)
    else:
        means, _ = _sse.PitSseMeanFunction.apply(problem, len(inputs), *inputs)
    out = {'pit_mse_loss': means[0]}
    if dual:
        out['pit_ips_loss'] = means[1]
    return out


def _dc_problem(embeddings, target_masks, lengths):
    """(DcProblem, autograd inputs) of a minibatch.  embeddings: list of [T_b, E, F] (the model's native
    't e f' layout, tcl/dc.py:66-74) or padded [B, T, E, F] with `lengths`; target_masks likewise [T_b, K, F]."""
    emb_list = _as_list(embeddings)
    if emb_list is not None:
        tgt_list = _as_list(target_masks)
        emb_list = [e if e.is_contiguous() else e.contiguous() for e in emb_list]
        tgt_list = [t if t.is_contiguous() else t.contiguous() for t in tgt_list]
        for e, t in zip(emb_list, tgt_list):
            _lib.require_cuda_float(e, 'embedding')
            _lib.require_cuda_float(t, 'target_mask')
            _check_dc_target(t)
        e_dim, bins = emb_list[0].shape[1:]
        k = tgt_list[0].shape[1]
        rows, splits, cursor = [], [], 0
        for e, t in zip(emb_list, tgt_list):
            assert e.shape[1:] == (e_dim, bins) and t.shape == (e.shape[0], k, bins), (e.shape, t.shape)
            rows.append([e.shape[0], _sse._offset(e, emb_list[0]), _sse._offset(t, tgt_list[0]), cursor])
            splits.append((cursor, e.numel(), e.shape))
            cursor += e.numel()
        meta = meta_tensor(rows, emb_list[0].device)
        problem = DcProblem(emb_list[0], tgt_list[0], meta, len(emb_list),
                            max(e.shape[0] for e in emb_list), bins, e_dim, k,
                            (e_dim * bins, bins, 1), (k * bins, bins, 1), cursor, splits,
                            (emb_list, tgt_list))
        return problem, emb_list
    emb = _lib.require_cuda_float(embeddings, 'embedding').contiguous()
    tgt = _lib.require_cuda_float(target_masks, 'target_mask').contiguous()
    _check_dc_target(tgt)
    batch, frames, e_dim, bins = emb.shape
    k = tgt.shape[2]
    assert tgt.shape == (batch, frames, k, bins), (emb.shape, tgt.shape)
    lengths = [frames] * batch if lengths is None else [int(v) for v in lengths]
    assert len(lengths) == batch and max(lengths, default=0) <= frames and min(lengths, default=0) >= 0, \
        (lengths, emb.shape)
    rows = [[lengths[b], b * frames * e_dim * bins, b * frames * k * bins, b * frames * e_dim * bins]
            for b in range(batch)]
    meta = meta_tensor(rows, emb.device, cache_key=('dc-padded', frames, e_dim, k, bins, tuple(lengths)))
    problem = DcProblem(emb, tgt, meta, batch, frames, bins, e_dim, k, (e_dim * bins, bins, 1),
                        (k * bins, bins, 1), emb.numel(), [(0, emb.numel(), emb.shape)],
                        covers_all=all(n == frames for n in lengths))
    return problem, [emb]


def dc_losses_per_example(embeddings, target_masks, lengths=None):
    """Per-example deep-clustering losses [B] (inputs as `_dc_problem`)."""
    problem, inputs = _dc_problem(embeddings, target_masks, lengths)
    return DcFunction.apply(problem, *inputs)


def dc_review_loss(embeddings, target_masks, lengths=None):
    """``dc_loss`` of DeepClusteringModel.review (tcl/dc.py:83-84): batch mean, folded by the Gram launch; the backward
    launch takes the mean's upstream gradient directly (two launches per training step in all)."""
    problem, inputs = _dc_problem(embeddings, target_masks, lengths)
    if problem.batch == 0:
        return DcFunction.apply(problem, *inputs).mean()
    return DcMeanFunction.apply(problem, *inputs)


_TASNET_KINDS = {
    'si-sdr': (_lib.LOSS_SI_SDR, _lib.REDUCE_MEAN),
    'log-mse': (_lib.LOSS_LOG_MSE, _lib.REDUCE_SUM),
    'log1p-mse': (_lib.LOSS_LOG1P_MSE, _lib.REDUCE_SUM),
}


class _TasnetFunction(torch.autograd.Function):
    """One statistics pass, three loss kinds (TasNet.loss evaluates all of them every step)."""

    @staticmethod
    def forward(ctx, estimate, problem, names):
        kinds = [_TASNET_KINDS[name][0] for name in names]
        reductions = [_TASNET_KINDS[name][1] for name in names]
        stats, _, perms, means = problem.stats_loss_set(kinds, reductions)   # two launches (or one: _pairs.py)
        ctx.problem, ctx.stats, ctx.perms, ctx.names = problem, stats, perms, names
        ctx.set_materialize_grads(False)   # losses without a weight get no backward pass
        return tuple(means[i] for i in range(len(names)))

    @staticmethod
    def backward(ctx, *grads):
        total = None
        examples = ctx.problem.groups // ctx.problem.inner
        for i, (name, grad) in enumerate(zip(ctx.names, grads)):
            if grad is None:
                continue
            kind, reduction = _TASNET_KINDS[name]
            g = ctx.problem.backward(ctx.stats, kind, 0, -1.0, reduction, True, ctx.perms[i], grad.reshape(1),
                                     broadcast_scale=1.0 / examples)
            total = g if total is None else total + g
        return total, None, None


def tasnet_losses(estimates, targets, num_samples, names=('si-sdr', 'log-mse', 'log1p-mse')):
    """Batch means of ``pit_loss(estimate[..., :n], target[..., :n], axis=0, loss_fn)`` for the three
    loss functions of TasNet.loss (tasnet/model.py:154-176).  estimates / targets: [B, K, T]."""
    _lib.require_cuda_float(estimates, 'estimates')
    _lib.require_cuda_float(targets, 'targets')
    _pairs.check_target(targets)
    e = estimates.contiguous()
    t = targets.contiguous()
    batch, k, length = e.shape
    num_samples = [int(n) for n in num_samples]
    assert len(num_samples) == batch and max(num_samples) <= length, (num_samples, e.shape)
    rows = [[num_samples[b], b * k * length, b * k * length] for b in range(batch)]
    meta = meta_tensor(rows, e.device, cache_key=('tasnet', k, length, tuple(num_samples)))
    problem = _pairs.PairProblem(e, t, meta, batch, 1, k, max(num_samples), length, length,
                                 covers_all=all(n == length for n in num_samples))
    values = _TasnetFunction.apply(e, problem, tuple(names))
    return dict(zip(names, values))


class _FusedStepFunction(torch.autograd.Function):
    """masks -> (loss [B], permutation [B, K]) through b2s_stft_pit_forward; the gradient w.r.t. the masks
    through b2s_stft_pit_backward (the target spectra are recomputed in registers in both directions)."""

    @staticmethod
    def forward(ctx, masks, mixture, observation_abs, sources, meta, stft, frames_call, pad_left, ragged):
        lib = _lib.load()
        batch, k, samples = sources.shape
        device = sources.device
        plan = stft._plan(device)
        loss = torch.empty(batch, dtype=torch.float32, device=device)
        perm = torch.empty((batch, k), dtype=torch.int32, device=device)
        sse = torch.empty((batch, k, k), dtype=torch.float64, device=device)
        ws = workspace(device, lib.b2s_stft_pit_workspace_bytes(batch, frames_call, k), 'fused')
        with torch.cuda.device(device):
            rc = lib.b2s_stft_pit_forward(
                plan.handle, _lib.ptr(mixture), _lib.ptr(observation_abs), _lib.ptr(sources),
                _lib.ptr(masks), _lib.ptr(meta), batch, samples, k, frames_call, pad_left, _lib.ptr(loss),
                _lib.ptr(perm), _lib.ptr(sse), _lib.ptr(ws), _lib.stream_of(device))
        _lib.check(rc, 'b2s_stft_pit_forward')
        ctx.save_for_backward(masks, perm)
        ctx.data = (mixture, observation_abs, sources, meta, stft, frames_call, pad_left, ragged)
        ctx.mark_non_differentiable(perm)
        return loss, perm

    @staticmethod
    def backward(ctx, grad_loss, _grad_perm):
        lib = _lib.load()
        masks, perm = ctx.saved_tensors
        mixture, observation_abs, sources, meta, stft, frames_call, pad_left, ragged = ctx.data
        batch, k, samples = sources.shape
        device = sources.device
        plan = stft._plan(device)
        # frames beyond an example's length are not written by the kernel
        grad = (torch.zeros_like if ragged else torch.empty_like)(masks)
        g = grad_loss.to(torch.float32).contiguous()
        with torch.cuda.device(device):
            rc = lib.b2s_stft_pit_backward(
                plan.handle, _lib.ptr(mixture), _lib.ptr(observation_abs), _lib.ptr(sources), _lib.ptr(masks),
                _lib.ptr(meta), batch, samples, k, frames_call, pad_left, _lib.ptr(perm), _lib.ptr(g),
                _lib.ptr(grad), _lib.stream_of(device))
        _lib.check(rc, 'b2s_stft_pit_backward')
        return grad, None, None, None, None, None, None, None, None


def stft_mask_pit_step(mixture, sources, masks, stft=None, observation_abs=None, num_samples=None):
    """The fused north-star step: per example ``pit_loss(mask * |STFT(y)|[:, None, :],
    |STFT(s)|, axis=-2)`` with the target spectra recomputed in registers (b2s_stft_pit_forward /
    b2s_stft_pit_backward).

    mixture [B, T] (may be None when `observation_abs` [B, M, F] is given), sources [B, K, T],
    masks [B, M, K, F].  Returns (loss [B], permutation [B, K] int32); differentiable w.r.t. `masks`
    (the waveforms and `observation_abs` are data on this path, as in pit/model.py:117-128).
    """
    stft = STFT(1024, 256) if stft is None else stft
    sources = _lib.require_cuda_float(sources, 'sources').detach().contiguous()
    masks = _lib.require_cuda_float(masks, 'masks').contiguous()
    batch, k, samples = sources.shape
    frames_call, pad_left = stft._frames_of_call(samples)
    assert masks.shape == (batch, frames_call, k, stft.size // 2 + 1), (masks.shape, frames_call)
    if mixture is not None:
        mixture = _lib.require_cuda_float(mixture, 'mixture').detach().contiguous()
        assert mixture.shape == (batch, samples), (mixture.shape, sources.shape)
    if observation_abs is not None:
        observation_abs = _lib.require_cuda_float(observation_abs, 'observation_abs').detach().contiguous()
        assert observation_abs.shape == (batch, frames_call, stft.size // 2 + 1)
    meta = None
    if num_samples is not None:
        assert len(num_samples) == batch and max(int(n) for n in num_samples) <= samples, (num_samples, samples)
        rows = [[int(n), stft._frames_of_call(int(n))[0]] for n in num_samples]
        meta = meta_tensor(rows, sources.device, cache_key=('fused', tuple(map(tuple, rows))))
    return _FusedStepFunction.apply(masks, mixture, observation_abs, sources, meta, stft, frames_call, pad_left,
                                    num_samples is not None)


def prepare_pit_targets(mixture, sources, stft=None, num_samples=None):
    """Batched ``pre_batch_transform`` of the PIT example on the device
    (contrib/examples/source_separation/pit/data.py:49-77): from raw waveforms
    mixture [B, T] and sources [B, K, T] to the tensors the model's review consumes,

        Y_abs [B, M, F], X_abs [B, M, K, F], cos_phase_difference [B, M, K, F], num_frames

    (the reference computes them per example with numpy on the data-loader workers and ships
    4 M F (1 + 2 K) bytes per utterance over PCIe instead of 4 T (1 + K)).  Forward only.

    `num_samples` (one length per example of a zero-padded batch): samples beyond an example's length are not read
    and the rows of frames beyond its own frame count are zeros -- the tensors equal the zero-padded collation of
    the per-example results; `num_frames` is then the list of frame counts.
    """
    stft = STFT(1024, 256) if stft is None else stft
    lib = _lib.load()
    mixture = _lib.require_cuda_float(mixture, 'mixture').detach().contiguous()
    sources = _lib.require_cuda_float(sources, 'sources').detach().contiguous()
    assert mixture.dim() == 2 and sources.dim() == 3 and sources.shape[0] == mixture.shape[0] and \
        sources.shape[2] == mixture.shape[1], (mixture.shape, sources.shape)
    batch, k, samples = sources.shape
    device = mixture.device
    plan = stft._plan(device)
    frames_call, pad_left = stft._frames_of_call(samples)
    meta, num_frames = None, frames_call
    if num_samples is not None:
        assert len(num_samples) == batch and max(int(n) for n in num_samples) <= samples, (num_samples, samples)
        rows = [[int(n), stft._frames_of_call(int(n))[0]] for n in num_samples]
        num_frames = [r[1] for r in rows]
        meta = meta_tensor(rows, device, cache_key=('targets', tuple(map(tuple, rows))))
    if (lib.b2s_stft_plan_is_fast(plan.handle) and stft.window_length == 1024 and stft.shift % 4 == 0
            and stft.shift <= 1024 and k <= 4 and batch * frames_call > 0):
        # one kernel: transforms, magnitudes and phase term in registers (b2s_stft_pit_targets)
        bins = stft.size // 2 + 1
        y_abs = torch.empty((batch, frames_call, bins), dtype=torch.float32, device=device)
        x_abs = torch.empty((batch, frames_call, k, bins), dtype=torch.float32, device=device)
        cpd = torch.empty((batch, frames_call, k, bins), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            rc = lib.b2s_stft_pit_targets(plan.handle, _lib.ptr(mixture), _lib.ptr(sources), _lib.ptr(meta), batch,
                                          samples, k, frames_call, pad_left, _lib.ptr(y_abs), _lib.ptr(x_abs),
                                          _lib.ptr(cpd), _lib.stream_of(device))
        _lib.check(rc, 'b2s_stft_pit_targets')
        return dict(Y_abs=y_abs, X_abs=x_abs, cos_phase_difference=cpd, num_frames=num_frames)
    if num_samples is not None:   # other plans: silence beyond each example's length, rows beyond its frames zeroed
        keep = torch.arange(samples, device=device)[None, :] < torch.as_tensor(num_samples, device=device)[:, None]
        mixture = torch.where(keep, mixture, torch.zeros_like(mixture))
        sources = torch.where(keep[:, None, :], sources, torch.zeros_like(sources))
    spec_y = stft._spectrum(mixture, _lib.SPEC_INTERLEAVED)       # [B, M, F, 2]
    spec_x = stft._spectrum(sources, _lib.SPEC_INTERLEAVED)       # [B, K, M, F, 2]
    frames, bins = spec_y.shape[1], spec_y.shape[2]
    y_abs = torch.empty((batch, frames, bins), dtype=torch.float32, device=device)
    x_abs = torch.empty((batch, frames, k, bins), dtype=torch.float32, device=device)
    cpd = torch.empty((batch, frames, k, bins), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = lib.b2s_pit_targets(_lib.ptr(spec_y), _lib.ptr(spec_x), batch, k, frames, bins,
                                 _lib.ptr(y_abs), _lib.ptr(x_abs), _lib.ptr(cpd), _lib.stream_of(device))
    _lib.check(rc, 'b2s_pit_targets')
    if num_samples is not None:
        live = torch.arange(frames, device=device)[None, :] < torch.as_tensor(num_frames, device=device)[:, None]
        y_abs = y_abs * live[:, :, None]
        x_abs = x_abs * live[:, :, None, None]
        cpd = cpd * live[:, :, None, None]
        return dict(Y_abs=y_abs, X_abs=x_abs, cos_phase_difference=cpd, num_frames=num_frames)
    return dict(Y_abs=y_abs, X_abs=x_abs, cos_phase_difference=cpd, num_frames=frames)


# ---------------------------------------------------------------------------------------------- evaluation path
def separate(masks, mixture_spectrum, stft, num_samples=None):
    """``z = istft(mask * Y[:, None, :])`` of the evaluation loop (pit/evaluate.py:147-152) for a whole batch on
    the device: masks [B, M, K, F], mixture_spectrum [B, M, F] complex64 (or [B, M, F, 2]) -> estimates
    [B, K, samples].  Two launches (b2s_mask_spectrum, b2s_istft_forward)."""
    lib = _lib.load()
    masks = _lib.require_cuda_float(masks, 'masks').detach().contiguous()
    spec = torch.view_as_real(mixture_spectrum) if torch.is_complex(mixture_spectrum) else mixture_spectrum
    spec = _lib.require_cuda_float(spec, 'mixture_spectrum').detach().contiguous()
    batch, frames, k, bins = masks.shape
    assert spec.shape == (batch, frames, bins, 2), (spec.shape, masks.shape)
    masked = torch.empty((batch, k, frames, bins, 2), dtype=torch.float32, device=masks.device)
    with torch.cuda.device(masks.device):
        rc = lib.b2s_mask_spectrum(_lib.ptr(masks), _lib.ptr(spec), batch, k, frames, bins, _lib.ptr(masked),
                                   _lib.stream_of(masks.device))
    _lib.check(rc, 'b2s_mask_spectrum')
    estimates = stft._inverse_layout(masked, _lib.SPEC_INTERLEAVED)      # [B, K, samples]
    if num_samples is not None:
        estimates = estimates[..., :max(int(n) for n in num_samples)]
    return estimates


def evaluate_separation(masks, mixture, sources, stft=None, num_samples=None):
    """Batched evaluation (pit/evaluate.py:144-176 with the metrics of this package
    in place of pb_bss / mir_eval): mask * STFT(y) -> iSTFT -> K x K SI-SDR / SDR matrices from ONE statistics
    pass -> assignment on the device.

    masks [B, M, K, F], mixture [B, T], sources [B, K, T].  Returns a dict of device tensors:
    ``estimates`` [B, K, T], ``permutation`` [B, K] (estimate matched with source k, chosen by SI-SDR),
    ``si_sdr`` / ``sdr`` [B, K] of the matched pairs in dB, ``input_si_sdr`` / ``input_sdr`` [B, K] (the
    mixture as the estimate of every source) and their ``*_improvement``."""
    from .ops.losses import source_separation as ss
    stft = STFT(1024, 256) if stft is None else stft
    lib = _lib.load()
    mixture = _lib.require_cuda_float(mixture, 'mixture').detach().contiguous()
    sources = _lib.require_cuda_float(sources, 'sources').detach().contiguous()
    batch, k, samples = sources.shape
    spectrum = stft._spectrum(mixture, _lib.SPEC_INTERLEAVED)            # [B, M, F, 2]
    estimates = separate(masks, spectrum, stft)[..., :samples].contiguous()
    if estimates.shape[-1] < samples:                                    # pad=False crops the tail
        sources = sources[..., :estimates.shape[-1]].contiguous()
        mixture = mixture[..., :estimates.shape[-1]].contiguous()
        samples = estimates.shape[-1]
    lengths = [samples] * batch if num_samples is None else [min(int(n), samples) for n in num_samples]

    def matrices(est):
        rows = [[lengths[b], b * k * samples, b * k * samples] for b in range(batch)]
        meta = meta_tensor(rows, est.device, cache_key=('eval', k, samples, tuple(lengths)))
        problem = _pairs.PairProblem(est, sources, meta, batch, 1, k, max(lengths), samples, samples)
        stats = problem.stats()
        out = []
        for kind in (_lib.LOSS_SI_SDR, _lib.LOSS_SDR):
            m = torch.empty((batch, k, k), dtype=torch.float32, device=est.device)
            with torch.cuda.device(est.device):
                rc = lib.b2s_pair_loss_matrix(_lib.ptr(stats), _lib.ptr(meta), batch, 1, k, kind, 0, -1.0,
                                              _lib.REDUCE_SUM, _lib.ptr(m), _lib.stream_of(est.device))
            _lib.check(rc, 'b2s_pair_loss_matrix')
            out.append(m)
        return out

    loss_si, loss_sdr = matrices(estimates)                  # entries are losses = -dB
    perm, _ = ss._first_minimum(loss_si, orientation=0)      # perm[b, k] = estimate matched with source k
    index = perm.long()[:, None, :]                          # gather rows perm[k] of column k
    si_sdr = -torch.gather(loss_si, 1, index)[:, 0, :]
    sdr = -torch.gather(loss_sdr, 1, index)[:, 0, :]
    observation = mixture[:, None, :].expand(batch, k, samples).contiguous()
    in_si, in_sdr = matrices(observation)
    diag = torch.arange(k, device=sources.device)
    input_si_sdr, input_sdr = -in_si[:, diag, diag], -in_sdr[:, diag, diag]
    return dict(estimates=estimates, permutation=perm, si_sdr=si_sdr, sdr=sdr, input_si_sdr=input_si_sdr,
                input_sdr=input_sdr, si_sdr_improvement=si_sdr - input_si_sdr, sdr_improvement=sdr - input_sdr)
