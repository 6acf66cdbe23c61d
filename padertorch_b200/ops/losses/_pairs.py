"""Shared driver of the pair-statistics kernels (b2s_pair_stats_forward / b2s_pair_loss /
b2s_pair_backward): every time-domain regression loss and its PIT variant goes through here."""
import ctypes
import os

import torch

from ... import _lib
from ..._workspace import meta_tensor, workspace


def _stats_width(k):
    return k * k + 4 * k


class PairProblem:
    """Geometry of one call: `groups` groups of K estimate rows / K target rows of length
    `lengths[g]`, `inner` consecutive groups per example."""

    def __init__(self, estimate, target, meta, groups, inner, k, max_length, est_stride, tgt_stride,
                 covers_all=False):
        self.estimate, self.target, self.meta = estimate, target, meta
        self.groups, self.inner, self.k = groups, inner, k
        self.max_length, self.est_stride, self.tgt_stride = max_length, est_stride, tgt_stride
        self.covers_all = covers_all     # every estimate element lies inside some group's length

    def stats(self):
        lib = _lib.load()
        device = self.estimate.device
        out = torch.empty((self.groups, _stats_width(self.k)), dtype=torch.float64, device=device)
        if self.groups == 0:
            return out
        nbytes = lib.b2s_pair_workspace_bytes(self.groups, self.max_length, self.k)
        ws = workspace(device, nbytes, 'pair')
        with torch.cuda.device(device):
            rc = lib.b2s_pair_stats_forward(
                _lib.ptr(self.estimate), _lib.ptr(self.target), _lib.ptr(self.meta), self.groups,
                self.max_length, self.k, self.est_stride, self.tgt_stride, _lib.ptr(out),
                _lib.ptr(ws), _lib.stream_of(device))
        _lib.check(rc, 'b2s_pair_stats_forward')
        return out

    def loss(self, stats, kind, flags, tau, reduction, pit):
        lib = _lib.load()
        device = self.estimate.device
        examples = self.groups // self.inner
        if pit or kind == _lib.LOSS_SA_SDR:
            loss = torch.empty(examples, dtype=torch.float32, device=device)
        else:
            loss = torch.empty(self.groups * self.k, dtype=torch.float32, device=device)
        perm = torch.empty((examples, self.k), dtype=torch.int32, device=device) if pit else None
        if self.groups:
            with torch.cuda.device(device):
                rc = lib.b2s_pair_loss(_lib.ptr(stats), _lib.ptr(self.meta), self.groups, self.inner,
                                       self.k, kind, flags, tau, reduction, int(pit), _lib.ptr(loss),
                                       _lib.ptr(perm), _lib.stream_of(device))
            _lib.check(rc, 'b2s_pair_loss')
        return loss, perm

    def loss_set(self, stats, kinds, reductions, flags=0, tau=-1.0):
        """Several PIT losses and their batch means in one launch: (loss [n, examples],
        perm [n, examples, K], mean [n])."""
        lib = _lib.load()
        device = self.estimate.device
        examples = self.groups // self.inner
        n = len(kinds)
        loss = torch.empty((n, examples), dtype=torch.float32, device=device)
        perm = torch.empty((n, examples, self.k), dtype=torch.int32, device=device)
        mean = torch.full((n,), float('nan'), dtype=torch.float32, device=device) if not self.groups else \
            torch.empty(n, dtype=torch.float32, device=device)
        if self.groups:
            c_kinds = (ctypes.c_int * n)(*kinds)
            c_reductions = (ctypes.c_int * n)(*reductions)
            with torch.cuda.device(device):
                rc = lib.b2s_pair_loss_set(_lib.ptr(stats), _lib.ptr(self.meta), self.groups, self.inner,
                                           self.k, n, c_kinds, c_reductions, flags, tau, _lib.ptr(loss),
                                           _lib.ptr(perm), _lib.ptr(mean), _lib.stream_of(device))
            _lib.check(rc, 'b2s_pair_loss_set')
        return loss, perm, mean

    def stats_loss_set(self, kinds, reductions, flags=0, tau=-1.0, one_launch=None):
        """`stats()` + `loss_set()`: (stats, loss, perm, mean).  one_launch=True: b2s_pair_stats_loss_set (the CTA
        that completes an example evaluates its losses, the last one the batch means).  Measured at batch 64 x 2 x
        4 s: 22.6 us against 21.8 us for the two launches, whose second kernel is launched programmatically and
        overlaps the tail of the first -- so two launches are the default (B2S_PAIR_ONE_LAUNCH=1 switches)."""
        if one_launch is None:
            one_launch = os.environ.get('B2S_PAIR_ONE_LAUNCH', '0') == '1'
        if self.inner != 1 or not self.groups or not one_launch:
            stats = self.stats()
            return (stats,) + self.loss_set(stats, kinds, reductions, flags, tau)
        lib = _lib.load()
        device = self.estimate.device
        n = len(kinds)
        stats = torch.empty((self.groups, _stats_width(self.k)), dtype=torch.float64, device=device)
        loss = torch.empty((n, self.groups), dtype=torch.float32, device=device)
        perm = torch.empty((n, self.groups, self.k), dtype=torch.int32, device=device)
        mean = torch.empty(n, dtype=torch.float32, device=device)
        ws = workspace(device, lib.b2s_pair_workspace_bytes(self.groups, self.max_length, self.k), 'pair')
        c_kinds = (ctypes.c_int * n)(*kinds)
        c_reductions = (ctypes.c_int * n)(*reductions)
        with torch.cuda.device(device):
            rc = lib.b2s_pair_stats_loss_set(
                _lib.ptr(self.estimate), _lib.ptr(self.target), _lib.ptr(self.meta), self.groups,
                self.max_length, self.k, self.est_stride, self.tgt_stride, n, c_kinds, c_reductions, flags, tau,
                _lib.ptr(stats), _lib.ptr(loss), _lib.ptr(perm), _lib.ptr(mean), _lib.ptr(ws),
                _lib.stream_of(device))
        _lib.check(rc, 'b2s_pair_stats_loss_set')
        return stats, loss, perm, mean

    def backward(self, stats, kind, flags, tau, reduction, pit, perm, grad_loss, broadcast_scale=None):
        """broadcast_scale: `grad_loss` is ONE upstream value (the gradient of a batch mean) applied to
        every example times this factor."""
        lib = _lib.load()
        device = self.estimate.device
        # padding beyond each length must stay zero; dense problems are overwritten completely
        grad = torch.empty_like(self.estimate) if self.covers_all and self.groups else torch.zeros_like(self.estimate)
        if self.groups:
            grad_loss = grad_loss.to(torch.float32).contiguous()
            stride, scale = (1, 1.0) if broadcast_scale is None else (0, float(broadcast_scale))
            with torch.cuda.device(device):
                rc = lib.b2s_pair_backward(
                    _lib.ptr(self.estimate), _lib.ptr(self.target), _lib.ptr(self.meta), self.groups,
                    self.inner, self.max_length, self.k, self.est_stride, self.tgt_stride,
                    _lib.ptr(stats), kind, flags, tau, reduction, int(pit), _lib.ptr(perm),
                    _lib.ptr(grad_loss), stride, scale, _lib.ptr(grad), _lib.stream_of(device))
            _lib.check(rc, 'b2s_pair_backward')
        return grad


class PairLossFunction(torch.autograd.Function):
    """estimate (dense float32) -> loss vector (+ permutation); gradient w.r.t. the estimate only
    (targets on this path are data, never parameters)."""

    @staticmethod
    def forward(ctx, estimate, problem, kind, flags, tau, reduction, pit):
        stats = problem.stats()
        loss, perm = problem.loss(stats, kind, flags, tau, reduction, pit)
        ctx.problem, ctx.stats, ctx.perm = problem, stats, perm
        ctx.args = (kind, flags, tau, reduction, pit)
        if perm is not None:
            ctx.mark_non_differentiable(perm)
            return loss, perm
        return loss, torch.empty(0, dtype=torch.int32, device=estimate.device)

    @staticmethod
    def backward(ctx, grad_loss, _grad_perm):
        kind, flags, tau, reduction, pit = ctx.args
        grad = ctx.problem.backward(ctx.stats, kind, flags, tau, reduction, pit, ctx.perm, grad_loss)
        return grad, None, None, None, None, None, None


def rowwise_problem(estimate, target):
    """Rows = all leading axes, one group per row (K = 1): the plain regression losses."""
    length = estimate.shape[-1]
    e = estimate.reshape(-1, length).contiguous()
    t = target.reshape(-1, length).contiguous()
    rows = e.shape[0]
    meta = dense_meta(rows, length, length, e.device)
    return e, PairProblem(e, t, meta, rows, 1, 1, length, length, length, covers_all=True)


def dense_meta(groups, length, group_stride, device):
    """meta rows {length, g * group_stride, g * group_stride}, built on the device and cached."""
    from ..._workspace import _meta_cache, _lock
    key = (torch.device(device).index, ('dense', groups, length, group_stride))
    hit = _meta_cache.get(key)
    if hit is not None:
        return hit
    offsets = torch.arange(groups, dtype=torch.int64, device=device) * group_stride
    table = torch.stack([torch.full_like(offsets, length), offsets, offsets], dim=1).contiguous()
    with _lock:
        if len(_meta_cache) >= 256:
            _meta_cache.clear()
        _meta_cache[key] = table
    return table


def check_target(target):
    if target.requires_grad:
        raise NotImplementedError(
            'padertorch_b200 regression losses do not propagate gradients into `target`')
