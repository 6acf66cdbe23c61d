"""Multi-GPU glue for the hot path: one process per GPU, utterances sharded by rank, and the single
exchange step of training -- the gradient reduction of the mask network -- done as a bucketed NCCL
all-reduce over NVLink/NVSwitch.

What this replaces: ``Trainer.train(device=[0, 1, ...])`` (padertorch/train/trainer.py:396-442) runs
ONE process with N Python threads, re-broadcasts every parameter to every GPU each iteration
(``replicate``, :408), and reduces the gradients onto GPU 0 (``ReduceAddCoalesced``).  Here every rank
keeps a persistent replica, so the per-iteration broadcast disappears and the reduce becomes an
all-reduce.  Semantics that are preserved:

* gradients and losses are **summed**, never averaged, across devices and virtual-minibatch rounds
  (trainer.py:87, :426-428);
* consecutive dataset items go to consecutive devices (``islice(train_iterable, len(device))``, :359):
  rank r of G takes items r, r + G, r + 2G, ...;
* ``virtual_minibatch_size % G == 0`` (:336); each rank runs ``virtual_minibatch_size // G`` rounds.

The kernels themselves need no collective: every hot-path op is per-utterance.
"""
import itertools

import torch
import torch.distributed as dist

DEFAULT_BUCKET_BYTES = 32 << 20   # sized for launch latency / overlap, not link count (NVSwitch)


def shard_for_rank(iterable, rank=None, world_size=None):
    """Items rank, rank + G, rank + 2G, ... of `iterable` (the reference hands consecutive examples
    to consecutive devices, trainer.py:359,415-419)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    return itertools.islice(iterable, rank, None, world_size)


def rounds_per_rank(virtual_minibatch_size, world_size=None):
    """Accumulation rounds each rank runs per optimizer step (trainer.py:336,357)."""
    world_size = dist.get_world_size() if world_size is None else world_size
    assert virtual_minibatch_size % world_size == 0, (
        f'virtual_minibatch_size={virtual_minibatch_size} must be divisible by the number of devices '
        f'({world_size}), as in padertorch.Trainer')
    return virtual_minibatch_size // world_size


def _buckets(tensors, bucket_bytes):
    bucket, size = [], 0
    for t in tensors:
        nbytes = t.numel() * t.element_size()
        if bucket and (size + nbytes > bucket_bytes or t.dtype != bucket[0].dtype
                       or t.device != bucket[0].device):
            yield bucket
            bucket, size = [], 0
        bucket.append(t)
        size += nbytes
    if bucket:
        yield bucket


def allreduce_gradients(parameters, extra=(), group=None, bucket_bytes=DEFAULT_BUCKET_BYTES,
                        async_op=False):
    """Sum the gradients of `parameters` (and the tensors in `extra`, e.g. the scalar loss) over all
    ranks, in place.  Gradients are packed into flat buckets, one all-reduce per bucket, issued in
    reverse parameter order (the order backward produces them).  Parameters without a gradient on
    this rank contribute zeros (every rank must reduce the same set)."""
    params = [p for p in parameters if p.requires_grad]
    grads = []
    for p in reversed(params):
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    tensors = list(extra) + grads
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return []
    works = []
    for bucket in _buckets(tensors, bucket_bytes):
        flat = torch.cat([t.reshape(-1) for t in bucket]) if len(bucket) > 1 else bucket[0].reshape(-1)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        works.append((work, flat, bucket))
    if async_op:
        return works
    finish_allreduce(works)
    return []


def finish_allreduce(works):
    """Wait for the buckets of allreduce_gradients(async_op=True) and scatter them back."""
    for work, flat, bucket in works:
        work.wait()
        if len(bucket) > 1 or flat.data_ptr() != bucket[0].data_ptr():
            offset = 0
            for t in bucket:
                t.copy_(flat[offset:offset + t.numel()].view_as(t))
                offset += t.numel()


def broadcast_parameters(module, src=0, group=None):
    """One-time synchronisation of the replicas (instead of the reference's per-iteration replicate)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in itertools.chain(module.parameters(), module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def distributed_trainer_class(trainer_class):
    """Subclass of a padertorch-style ``Trainer`` whose optimizer step first all-reduces (sums) the
    gradients over the process group.  Run one process per GPU, give each the same model /
    optimizer / seed, a dataset sharded with `shard_for_rank`, ``device=LOCAL_RANK`` and
    ``virtual_minibatch_size = rounds_per_rank(total)``; hooks, checkpoints and `test_run` are
    untouched (keep summaries / checkpoints on rank 0 by giving the other ranks a scratch dir)."""

    class DistributedTrainer(trainer_class):

        def optimizer_step(self, *args, **kwargs):
            allreduce_gradients(self.model.parameters())
            return super().optimizer_step(*args, **kwargs)

        def train(self, *args, **kwargs):
            broadcast_parameters(self.model)
            return super().train(*args, **kwargs)

    DistributedTrainer.__name__ = f'Distributed{trainer_class.__name__}'
    return DistributedTrainer
