"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding rule, sum-semantics gradient
all-reduce == single-process run with virtual_minibatch_size = world size."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _examples():
    g = torch.Generator().manual_seed(1)
    return [(torch.randn(4, 6, generator=g), torch.randn(4, 3, generator=g)) for _ in range(8)]


class _Trainer:
    """Minimal stand-in with padertorch.Trainer's structure: summed losses over the virtual minibatch,
    optimizer_step() after the rounds (train/trainer.py:357-447)."""

    def __init__(self, model, virtual_minibatch_size):
        self.model, self.virtual_minibatch_size = model, virtual_minibatch_size
        self.optimizer = torch.optim.SGD(model.parameters(), lr=0.1)
        self.optimizer.zero_grad()

    def optimizer_step(self):
        self.optimizer.step()
        self.optimizer.zero_grad()

    def train(self, examples, steps):
        it = iter(examples)
        for _ in range(steps):
            for _ in range(self.virtual_minibatch_size):
                x, y = next(it)
                torch.nn.functional.mse_loss(self.model(x), y, reduction='sum').backward()
            self.optimizer_step()


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from padertorch_b200 import parallel
    try:
        # sharding rule
        assert list(parallel.shard_for_rank(range(7))) == list(range(rank, 7, world))
        assert parallel.rounds_per_rank(4) == 4 // world
        with pytest.raises(AssertionError):
            parallel.rounds_per_rank(3)
        # gradient all-reduce with tiny buckets (forces several), incl. an extra scalar and a missing grad
        model = _model()
        model[2].bias.requires_grad_(True)
        x, y = _examples()[rank]
        loss = torch.nn.functional.mse_loss(model(x), y, reduction='sum')
        loss.backward()
        model[2].bias.grad = None if rank == 1 else model[2].bias.grad
        total = loss.detach().clone()
        parallel.allreduce_gradients(model.parameters(), extra=[total], bucket_bytes=64)
        torch.save(dict(grads=[p.grad.clone() for p in model.parameters()], loss=total),
                   os.path.join(outdir, f'reduce_{rank}.pt'))
        # distributed trainer == single process with vmb = world
        Trainer = parallel.distributed_trainer_class(_Trainer)
        model = _model()
        if rank == 1:            # replicas start different; train() must synchronise them once
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        trainer = Trainer(model, parallel.rounds_per_rank(2))
        trainer.train(parallel.shard_for_rank(_examples()), steps=3)
        torch.save([p.detach().clone() for p in model.parameters()], os.path.join(outdir, f'train_{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, _free_port(), outdir), nprocs=world, join=True)
        # expected reduced gradients: sum of the per-rank gradients
        expect, losses = None, 0.0
        for rank in range(world):
            model = _model()
            x, y = _examples()[rank]
            loss = torch.nn.functional.mse_loss(model(x), y, reduction='sum')
            loss.backward()
            grads = [p.grad.clone() for p in model.parameters()]
            if rank == 1:
                grads[-1].zero_()          # that rank had no gradient for the last bias
            expect = grads if expect is None else [a + b for a, b in zip(expect, grads)]
            losses += float(loss)
        for rank in range(world):
            got = torch.load(os.path.join(outdir, f'reduce_{rank}.pt'))
            for g, e in zip(got['grads'], expect):
                torch.testing.assert_close(g, e, rtol=1e-6, atol=1e-6)
            assert abs(float(got['loss']) - losses) < 1e-4
        # single-process reference run: virtual minibatch of `world` consecutive examples per step
        single = _Trainer(_model(), world)
        single.train(_examples(), steps=3)
        for rank in range(world):
            params = torch.load(os.path.join(outdir, f'train_{rank}.pt'))
            for p, q in zip(params, single.model.parameters()):
                torch.testing.assert_close(p, q.detach(), rtol=1e-5, atol=1e-6)
