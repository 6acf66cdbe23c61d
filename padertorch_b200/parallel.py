"""Multi-GPU glue for the hot path: one process per GPU, utterances sharded by rank, and the single
exchange step of training -- the gradient reduction of the mask network -- as bucketed NCCL all-reduces over
NVLink / NVSwitch that are issued WHILE backward is still producing the remaining gradients.

What this replaces: ``Trainer.train(device=[0, 1, ...])`` (padertorch/train/trainer.py:396-442) runs
ONE process with N Python threads, re-broadcasts every parameter to every GPU each iteration
(``replicate``, :408), and reduces the gradients onto GPU 0 (``ReduceAddCoalesced``).  Here every rank
keeps a persistent replica, so the per-iteration broadcast disappears and the reduce becomes an
all-reduce.  Semantics that are preserved:

* gradients and losses are **summed**, never averaged, across devices and virtual-minibatch rounds
  (trainer.py:87, :426-428);
* consecutive dataset items go to consecutive devices (``islice(train_iterable, len(device))``, :359):
  rank r of G takes items r, r + G, r + 2G, ... -- of COMPLETE groups of G items only, so that every rank
  sees the same number of items and issues the same number of collectives (an incomplete last group is
  dropped; the reference would run it on fewer devices);
* ``virtual_minibatch_size % G == 0`` (:336); each rank runs ``virtual_minibatch_size // G`` rounds.

Gradient exchange (``GradientExchange``): the parameters' ``.grad`` are views into a few flat, pre-allocated
fp32 buckets (default 32 MiB: sized for launch latency and overlap, not for link count -- NVSwitch gives every
pair full bandwidth) laid out in reverse parameter order, the order in which backward finishes them.  A
post-accumulate hook per parameter counts a bucket's finished gradients; the bucket's ``all_reduce`` is issued
asynchronously the moment it is complete in the LAST accumulation round of the optimizer step, so only the
last bucket's transfer is exposed.  No ``torch.cat``, no copy back.  The scalar loss rides in one extra slot of
the last bucket.

The kernels themselves need no collective: every hot-path op is per-utterance.
"""
import itertools
import os
import tempfile

import torch
import torch.distributed as dist

DEFAULT_BUCKET_BYTES = 32 << 20


def _world(world_size=None, group=None):
    if world_size is not None:
        return world_size
    return dist.get_world_size(group) if dist.is_initialized() else 1


def shard_for_rank(iterable, rank=None, world_size=None):
    """Item ``rank`` of every COMPLETE group of ``world_size`` consecutive items of `iterable` (the reference
    hands consecutive examples to consecutive devices, trainer.py:359,415-419).  Every rank gets exactly
    ``len(iterable) // world_size`` items: ranks that ran out of data earlier than others would skip an
    optimizer step (trainer.py:360-366) the others still take, and the collectives would no longer pair up."""
    rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
    world_size = _world(world_size)
    iterator = iter(iterable)
    while True:
        group = list(itertools.islice(iterator, world_size))
        if len(group) < world_size:
            return
        yield group[rank]


class ShardedDataset:
    """Re-iterable view of a dataset for one rank (``Trainer.train`` iterates its dataset once per epoch)."""

    def __init__(self, dataset, rank=None, world_size=None):
        self.dataset, self.rank, self.world_size = dataset, rank, world_size

    def __iter__(self):
        return shard_for_rank(self.dataset, self.rank, self.world_size)

    def __len__(self):
        return len(self.dataset) // _world(self.world_size)


def rounds_per_rank(virtual_minibatch_size, world_size=None):
    """Accumulation rounds each rank runs per optimizer step (trainer.py:336,357)."""
    world_size = _world(world_size)
    assert virtual_minibatch_size % world_size == 0, (
        f'virtual_minibatch_size={virtual_minibatch_size} must be divisible by the number of devices '
        f'({world_size}), as in padertorch.Trainer')
    return virtual_minibatch_size // world_size


def rank_storage_dir(storage_dir, rank=None):
    """`storage_dir` on rank 0, a scratch directory elsewhere: summaries and checkpoints are written once."""
    rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
    return storage_dir if rank == 0 else tempfile.mkdtemp(prefix=f'b2s_rank{rank}_')


class GradientExchange:
    """Sum-all-reduce of the gradients of `parameters`, overlapped with backward.

    ``rounds``: backward passes per optimizer step on this rank (virtual minibatch rounds); only the last one
    triggers the collectives.  Usage::

        exchange = GradientExchange(model.parameters(), rounds=1)
        loss.backward()                  # buckets are all-reduced as they complete
        exchange.finish()                # wait; gradients (and exchange.loss) now hold the sums over all ranks
        optimizer.step(); exchange.zero_grad()
    """

    def __init__(self, parameters, rounds=1, group=None, bucket_bytes=DEFAULT_BUCKET_BYTES):
        self.group, self.rounds = group, int(rounds)
        self.params = [p for p in parameters if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._buckets = []            # dicts: flat, params, pending, work
        self._where = {}              # id(param) -> bucket index
        self._round, self._launched, self._complete = 0, 0, 0
        self._hooks = []
        cursor = []

        def close(extra=0):
            if not cursor:
                return
            ref = cursor[0]
            total = sum(p.numel() for p in cursor) + extra
            flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
            offset = 0
            for p in cursor:
                view = flat[offset:offset + p.numel()].view_as(p)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
                self._where[id(p)] = len(self._buckets)
                offset += p.numel()
            self._buckets.append(dict(flat=flat, params=list(cursor), pending=len(cursor), work=None, used=offset))
            cursor.clear()

        size = 0
        for p in reversed(self.params):           # the order in which backward finishes them
            nbytes = p.numel() * p.element_size()
            if cursor and (size + nbytes > bucket_bytes or p.dtype != cursor[0].dtype or p.device != cursor[0].device):
                close()
                size = 0
            cursor.append(p)
            size += nbytes
        close(extra=1)                             # one more slot in the last bucket: the scalar loss
        last = self._buckets[-1] if self._buckets else None
        self.loss = last['flat'][last['used']:last['used'] + 1] if last is not None else torch.zeros(1)
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ---- bookkeeping -------------------------------------------------------------------------------------
    def _on_grad(self, param):
        index = self._where[id(param)]
        bucket = self._buckets[index]
        view_ptr = bucket['flat'].data_ptr()
        grad = param.grad
        # autograd replaced the view (first backward after a zero_grad(set_to_none=True)): restore it
        if not (view_ptr <= grad.data_ptr() < view_ptr + bucket['flat'].numel() * bucket['flat'].element_size()):
            offset = sum(q.numel() for q in bucket['params'][:bucket['params'].index(param)])
            view = bucket['flat'][offset:offset + param.numel()].view_as(param)
            view.copy_(grad)
            param.grad = view
        bucket['pending'] -= 1
        if bucket['pending'] == 0:
            self._complete += 1
            if self._round == self.rounds - 1:
                self._launch(index)
            if self._complete == len(self._buckets):   # every gradient of this backward pass has arrived
                self.end_round()

    def _launch(self, index):
        bucket = self._buckets[index]
        if bucket['work'] is None and self.world > 1:
            bucket['work'] = dist.all_reduce(bucket['flat'], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._launched += 1

    def add_loss(self, loss):
        """Accumulate this rank's (detached) scalar loss; after finish() `exchange.loss` is the global sum.
        Call before the backward pass of the round the loss belongs to."""
        self.loss += loss.detach().to(self.loss.dtype).reshape(1)

    def end_round(self):
        """End of one backward pass.  Detected automatically when every parameter received a gradient; call it
        after ``loss.backward()`` yourself if some parameters are unused in a pass."""
        for bucket in self._buckets:
            bucket['pending'] = len(bucket['params'])
        self._complete = 0
        self._round += 1

    def finish(self):
        """Issue what is still outstanding (parameters that received no gradient on this rank contribute zeros:
        every rank must reduce the same buckets) and wait for all buckets."""
        for index, bucket in enumerate(self._buckets):
            if bucket['work'] is None:
                self._launch(index)
        for bucket in self._buckets:
            if bucket['work'] is not None:
                bucket['work'].wait()
                bucket['work'] = None
        self._round, self._launched, self._complete = 0, 0, 0
        for bucket in self._buckets:
            bucket['pending'] = len(bucket['params'])

    def zero_grad(self):
        """One memset per bucket; the gradients stay views of the buckets."""
        for bucket in self._buckets:
            bucket['flat'].zero_()
        for bucket in self._buckets:
            offset = 0
            for p in bucket['params']:
                if p.grad is None or p.grad.data_ptr() != bucket['flat'].data_ptr() + offset * bucket['flat'].element_size():
                    p.grad = bucket['flat'][offset:offset + p.numel()].view_as(p)
                offset += p.numel()

    def nbytes(self):
        return sum(b['flat'].numel() * b['flat'].element_size() for b in self._buckets)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(parameters, extra=(), group=None, bucket_bytes=DEFAULT_BUCKET_BYTES):
    """Blocking one-shot form (no persistent buckets): sum the gradients of `parameters` and the tensors in
    `extra` over all ranks, in place.  Parameters without a gradient contribute zeros."""
    params = [p for p in parameters if p.requires_grad]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    tensors = list(extra) + [p.grad for p in reversed(params)]
    works, bucket, size = [], [], 0

    def flush():
        if not bucket:
            return
        flat = torch._utils._flatten_dense_tensors(bucket) if len(bucket) > 1 else bucket[0].reshape(-1)
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, list(bucket)))
        bucket.clear()

    for t in tensors:
        nbytes = t.numel() * t.element_size()
        if bucket and (size + nbytes > bucket_bytes or t.dtype != bucket[0].dtype or t.device != bucket[0].device):
            flush()
            size = 0
        bucket.append(t)
        size += nbytes
    flush()
    for work, flat, members in works:
        work.wait()
        if len(members) > 1:
            for t, synced in zip(members, torch._utils._unflatten_dense_tensors(flat, members)):
                t.copy_(synced)


def broadcast_parameters(module, src=0, group=None):
    """One-time synchronisation of the replicas (instead of the reference's per-iteration replicate)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in itertools.chain(module.parameters(), module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def distributed_trainer_class(trainer_class):
    """Subclass of a padertorch-style ``Trainer`` for one-process-per-GPU training.  Per rank::

        Trainer = distributed_trainer_class(pt.Trainer)
        trainer = Trainer(model, rank_storage_dir(storage_dir), optimizer, ...,
                          virtual_minibatch_size=rounds_per_rank(total_virtual_minibatch_size))
        trainer.train(ShardedDataset(dataset), device=LOCAL_RANK)

    * ``to(device)`` (called by ``train`` before the loop, trainer.py:285-293) moves the model FIRST, then
      synchronises the replicas once from rank 0 and builds the gradient buckets on the device;
    * every ``loss.backward()`` of ``train`` feeds the buckets; the last round of an optimizer step launches the
      all-reduces bucket by bucket while backward is still running;
    * ``optimizer_step`` waits for them, then clips / steps exactly like the base class -- on the SUMMED gradient,
      identically on every rank; ``optimizer_zero_grad`` zeroes the buckets in place.
    Hooks, checkpoints, ``test_run`` are untouched."""

    class DistributedTrainer(trainer_class):
        _exchange = None

        def to(self, device):
            result = super().to(device)
            if device is not None and dist.is_initialized() and dist.get_world_size() > 1:
                broadcast_parameters(self.model)
                if self._exchange is not None:
                    self._exchange.close()
                self._exchange = GradientExchange(self.model.parameters(),
                                                  rounds=getattr(self, 'virtual_minibatch_size', 1))
                self._install_round_marker()
            return result

        def _install_round_marker(self):
            # the base class calls loss.backward() itself (trainer.py:390,441): the loss of every round is added to
            # the exchange before that backward pass starts (it travels in the last bucket)
            exchange, original = self._exchange, self.train_step

            def train_step(model, example, device):
                out = original(model, example, device)
                exchange.add_loss(out[0])
                return out
            self.train_step = train_step

        def optimizer_step(self, *args, **kwargs):
            if self._exchange is not None:
                self._exchange.finish()
                self.last_loss_sum = float(self._exchange.loss)
            return super().optimizer_step(*args, **kwargs)

        def optimizer_zero_grad(self):
            if self._exchange is not None:
                self._exchange.zero_grad()
            else:
                super().optimizer_zero_grad()

    DistributedTrainer.__name__ = f'Distributed{trainer_class.__name__}'
    return DistributedTrainer


def nccl_debug_to_file(directory=None):
    """Route NCCL's INFO log (rings / trees / NVLS, number of ranks) to per-rank files instead of silencing it;
    returns the file pattern.  Call before ``init_process_group``."""
    directory = directory or tempfile.gettempdir()
    pattern = os.path.join(directory, 'nccl_%h_%p.log')
    os.environ.setdefault('NCCL_DEBUG', 'INFO')
    os.environ.setdefault('NCCL_DEBUG_FILE', pattern)
    return pattern
