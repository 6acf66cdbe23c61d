"""Driver of the spectral PIT kernels (b2s_pit_sse_forward / b2s_pit_sse_backward)."""
import torch

from ... import _lib
from ..._workspace import meta_tensor, workspace


class SseProblem:
    """One launch worth of examples.  `blocks` are the per-example mask / target / observation /
    scale tensors expressed as offsets (in floats) from base tensors; see include/b200sep.h."""

    def __init__(self, mask_base, obs_base, target_base, scale_base, meta, batch, max_frames,
                 sources, bins, dual, grad_numel, grad_splits, keep_alive=(), covers_all=True):
        self.mask_base, self.obs_base = mask_base, obs_base
        self.target_base, self.scale_base = target_base, scale_base
        self.meta, self.batch, self.max_frames = meta, batch, max_frames
        self.sources, self.bins, self.dual = sources, bins, dual
        self.grad_numel, self.grad_splits = grad_numel, grad_splits
        self.keep_alive = keep_alive
        self.covers_all = covers_all     # every element of the gradient buffer belongs to some example

    @property
    def device(self):
        return self.mask_base.device

    def forward(self, with_mean=False):
        """(loss [slots, B], perm, sse) or, with_mean, (loss, perm, sse, mean [slots])."""
        lib = _lib.load()
        slots = 2 if self.dual else 1
        k = self.sources
        loss = torch.empty((slots, self.batch), dtype=torch.float32, device=self.device)
        perm = torch.empty((slots, self.batch, k), dtype=torch.int32, device=self.device)
        sse = torch.empty((self.batch, slots, k, k), dtype=torch.float64, device=self.device)
        mean = torch.empty(slots, dtype=torch.float32, device=self.device) if with_mean else None
        if self.batch:
            nbytes = lib.b2s_pit_workspace_bytes(self.batch, self.max_frames, self.bins, k, int(self.dual))
            ws = workspace(self.device, nbytes, 'pit')
            with torch.cuda.device(self.device):
                if with_mean:
                    rc = lib.b2s_pit_sse_forward_mean(
                        _lib.ptr(self.mask_base), _lib.ptr(self.obs_base), _lib.ptr(self.target_base),
                        _lib.ptr(self.scale_base), _lib.ptr(self.meta), self.batch, self.max_frames, k,
                        self.bins, int(self.dual), _lib.ptr(loss), _lib.ptr(mean), _lib.ptr(perm), _lib.ptr(sse),
                        _lib.ptr(ws), _lib.stream_of(self.device))
                else:
                    rc = lib.b2s_pit_sse_forward(
                        _lib.ptr(self.mask_base), _lib.ptr(self.obs_base), _lib.ptr(self.target_base),
                        _lib.ptr(self.scale_base), _lib.ptr(self.meta), self.batch, self.max_frames, k,
                        self.bins, int(self.dual), _lib.ptr(loss), _lib.ptr(perm), _lib.ptr(sse),
                        _lib.ptr(ws), _lib.stream_of(self.device))
            _lib.check(rc, 'b2s_pit_sse_forward')
        elif with_mean:
            mean.fill_(float('nan'))   # torch.mean of an empty batch
        return (loss, perm, sse, mean) if with_mean else (loss, perm, sse)

    def backward(self, perm, grad_loss, want_target_grad=False, broadcast_scale=None):
        """broadcast_scale: `grad_loss` holds ONE upstream value per slot (the gradients of the batch means), applied to
        every example times this factor."""
        lib = _lib.load()
        alloc = torch.empty if self.covers_all and self.batch else torch.zeros
        grad_mask = alloc(self.grad_numel, dtype=torch.float32, device=self.device)
        grad_target = alloc(self.grad_numel, dtype=torch.float32, device=self.device) if want_target_grad else None
        if self.batch:
            grad_loss = grad_loss.to(torch.float32).contiguous()
            stride, scale = (1, 1.0) if broadcast_scale is None else (0, float(broadcast_scale))
            with torch.cuda.device(self.device):
                rc = lib.b2s_pit_sse_backward_scaled(
                    _lib.ptr(self.mask_base), _lib.ptr(self.obs_base), _lib.ptr(self.target_base),
                    _lib.ptr(self.scale_base), _lib.ptr(self.meta), self.batch, self.max_frames,
                    self.sources, self.bins, int(self.dual), _lib.ptr(perm), _lib.ptr(grad_loss), stride, scale,
                    _lib.ptr(grad_mask), _lib.ptr(grad_target), _lib.stream_of(self.device))
            _lib.check(rc, 'b2s_pit_sse_backward_scaled')
        return grad_mask, grad_target

    def split(self, flat):
        """Carve the flat gradient buffer into tensors shaped like the differentiable inputs."""
        return [flat[start:start + numel].view(shape) for start, numel, shape in self.grad_splits]


class PitSseFunction(torch.autograd.Function):
    """(problem, n_masks, *mask tensors[, target tensor]) -> (loss [slots, B], perm [slots, B, K])."""

    @staticmethod
    def forward(ctx, problem, n_masks, *tensors):
        loss, perm, sse = problem.forward()
        ctx.problem, ctx.perm, ctx.n_masks, ctx.n_tensors = problem, perm, n_masks, len(tensors)
        ctx.mark_non_differentiable(perm, sse)
        return loss, perm, sse

    @staticmethod
    def backward(ctx, grad_loss, _grad_perm, _grad_sse):
        problem = ctx.problem
        want_target = ctx.n_tensors > ctx.n_masks and ctx.needs_input_grad[2 + ctx.n_masks]
        grad_mask, grad_target = problem.backward(ctx.perm, grad_loss, want_target)
        grads = problem.split(grad_mask)
        if ctx.n_tensors > ctx.n_masks:
            grads = grads + ([grad_target.view(problem.grad_splits[0][2])] if want_target else [None])
        return (None, None, *grads)


class PitSseMeanFunction(torch.autograd.Function):
    """(problem, n_masks, *mask tensors) -> (mean [slots], perm): the batch means of the per-example PIT losses (the
    `losses` entry of PermutationInvariantTrainingModel.review, pit/model.py:137-140) from one extra one-warp launch;
    the backward takes the means' upstream gradients directly (factor 1 / batch inside the kernel)."""

    @staticmethod
    def forward(ctx, problem, n_masks, *tensors):
        _, perm, _, mean = problem.forward(with_mean=True)
        ctx.problem, ctx.perm = problem, perm
        ctx.mark_non_differentiable(perm)
        return mean, perm

    @staticmethod
    def backward(ctx, grad_mean, _grad_perm):
        problem = ctx.problem
        grad_mask, _ = problem.backward(ctx.perm, grad_mean, False, broadcast_scale=1.0 / max(problem.batch, 1))
        return (None, None, *problem.split(grad_mask))


def _offset(tensor, base):
    delta = tensor.data_ptr() - base.data_ptr()
    assert delta % 4 == 0, delta
    return delta // 4


def dense_problem(estimate, target):
    """Single example, estimate / target contiguous [outer, K, inner] (the op-level pit_loss)."""
    outer, k, inner = estimate.shape
    meta = meta_tensor([[outer, 0, 0, 0, 0, 0]], estimate.device, cache_key=('sse-dense', outer))
    numel = estimate.numel()
    return SseProblem(estimate, None, target, None, meta, 1, outer, k, inner, False, numel,
                      [(0, numel, estimate.shape)])


def list_problem(masks, observations, targets, scales=None, dual=False):
    """Ragged batch given as lists of per-example tensors: masks / targets / scales [T_b, K, F],
    observations [T_b, F] (None: masks already are the estimates)."""
    masks = [m if m.is_contiguous() else m.contiguous() for m in masks]
    targets = [t if t.is_contiguous() else t.contiguous() for t in targets]
    if observations is not None:
        observations = [o if o.is_contiguous() else o.contiguous() for o in observations]
    if scales is not None:
        scales = [s if s.is_contiguous() else s.contiguous() for s in scales]
    batch = len(masks)
    k, bins = masks[0].shape[-2], masks[0].shape[-1]
    rows, splits, cursor = [], [], 0
    for b in range(batch):
        frames = masks[b].shape[0]
        assert masks[b].shape == targets[b].shape, (masks[b].shape, targets[b].shape)
        assert masks[b].shape[1:] == (k, bins), (masks[b].shape, k, bins)
        if observations is not None:
            assert observations[b].shape == (frames, bins), (observations[b].shape, frames, bins)
        if scales is not None:
            assert scales[b].shape == masks[b].shape, (scales[b].shape, masks[b].shape)
        rows.append([
            frames, _offset(masks[b], masks[0]),
            _offset(observations[b], observations[0]) if observations is not None else 0,
            _offset(targets[b], targets[0]),
            _offset(scales[b], scales[0]) if scales is not None else 0,
            cursor])
        splits.append((cursor, masks[b].numel(), masks[b].shape))
        cursor += masks[b].numel()
    device = masks[0].device
    meta = meta_tensor(rows, device)
    keep = (masks, observations, targets, scales)
    return SseProblem(masks[0], observations[0] if observations is not None else None, targets[0],
                      scales[0] if scales is not None else None, meta, batch,
                      max(m.shape[0] for m in masks), k, bins, dual, cursor, splits, keep), masks


def padded_problem(mask, observation, target, scale, lengths, dual=False):
    """Padded batch: mask / target / scale [B, T, K, F], observation [B, T, F] (or None), `lengths`
    a sequence of valid frame counts (None: all T)."""
    mask = mask.contiguous()
    target = target.contiguous()
    observation = None if observation is None else observation.contiguous()
    scale = None if scale is None else scale.contiguous()
    batch, frames, k, bins = mask.shape
    assert target.shape == mask.shape, (target.shape, mask.shape)
    lengths = [frames] * batch if lengths is None else [int(v) for v in lengths]
    assert len(lengths) == batch and max(lengths, default=0) <= frames, (lengths, mask.shape)
    block = frames * k * bins
    rows = [[lengths[b], b * block, b * frames * bins, b * block, b * block, b * block]
            for b in range(batch)]
    meta = meta_tensor(rows, mask.device, cache_key=('sse-padded', frames, k, bins, tuple(lengths)))
    return SseProblem(mask, observation, target, scale, meta, batch, frames, k, bins, dual,
                      mask.numel(), [(0, mask.numel(), mask.shape)],
                      covers_all=all(n == frames for n in lengths)), mask
