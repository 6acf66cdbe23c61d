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
        # item `rank` of every COMPLETE group of `world` items: equal counts on every rank
        assert list(parallel.shard_for_rank(range(7))) == list(range(rank, 6, world))
        assert len(parallel.ShardedDataset(list(range(7)))) == 3
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
        # overlapped exchange with persistent buckets: 2 accumulation rounds per step, tiny buckets, loss slot
        model = _model()
        exchange = parallel.GradientExchange(model.parameters(), rounds=2, bucket_bytes=64)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        losses = []
        examples = list(parallel.shard_for_rank(_examples()))
        for step in range(2):
            for x, y in examples[2 * step:2 * step + 2]:
                loss = torch.nn.functional.mse_loss(model(x), y, reduction='sum')
                exchange.add_loss(loss)
                loss.backward()
            exchange.finish()
            losses.append(float(exchange.loss))
            opt.step()
            exchange.zero_grad()
        torch.save(dict(params=[p.detach().clone() for p in model.parameters()], losses=losses),
                   os.path.join(outdir, f'exchange_{rank}.pt'))
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
        # overlapped exchange: 2 ranks x 2 rounds per step == one process with 4 consecutive examples per step
        model = _model()
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        want_losses = []
        for step in range(2):
            opt.zero_grad()
            total = 0.0
            for x, y in _examples()[4 * step:4 * step + 4]:
                loss = torch.nn.functional.mse_loss(model(x), y, reduction='sum')
                loss.backward()
                total += float(loss)
            want_losses.append(total)
            opt.step()
        for rank in range(world):
            got = torch.load(os.path.join(outdir, f'exchange_{rank}.pt'))
            for p, q in zip(got['params'], model.parameters()):
                torch.testing.assert_close(p, q.detach(), rtol=1e-5, atol=1e-6)
            assert all(abs(a - b) < 1e-3 for a, b in zip(got['losses'], want_losses)), (got['losses'], want_losses)


# ---------------------------------------------------------------------------------------------- the REAL Trainer
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
STANDINS = os.path.join(ROOT, 'oracle', 'ref_standins')


def _import_reference():
    import sys
    for path in (STANDINS, REF):
        if path not in sys.path:
            sys.path.insert(0, path)
    import warnings
    warnings.filterwarnings('ignore')
    import padertorch as pt
    return pt


def _toy_model(pt):
    class Toy(pt.Model):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))

        def forward(self, example):
            return self.net(example['x'])

        def review(self, example, output):
            return dict(loss=torch.nn.functional.mse_loss(output, example['y'], reduction='sum'))
    return Toy()


def _dict_examples(n=9):
    g = torch.Generator().manual_seed(1)
    return [dict(x=torch.randn(4, 6, generator=g), y=torch.randn(4, 3, generator=g)) for _ in range(n)]


def _real_trainer_worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pt = _import_reference()
        from padertorch_b200 import parallel
        Trainer = parallel.distributed_trainer_class(pt.Trainer)
        model = _toy_model(pt)
        if rank == 1:        # different replicas: to() must synchronise them AFTER moving the model
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(0.5)
        trainer = Trainer(model, os.path.join(parallel.rank_storage_dir(os.path.join(outdir, 'run')), 'x'),
                          optimizer=pt.optimizer.SGD(lr=0.1), stop_trigger=(2, 'iteration'),
                          summary_trigger=(1, 'iteration'), checkpoint_trigger=(100, 'iteration'),
                          virtual_minibatch_size=parallel.rounds_per_rank(4))
        # 9 examples over 2 ranks: the incomplete last group is dropped, both ranks see 4
        trainer.train(parallel.ShardedDataset(_dict_examples(9)), device='cpu', progress_bar=False)
        torch.save(dict(params=[p.detach().clone() for p in model.parameters()], iteration=trainer.iteration,
                        loss=trainer.last_loss_sum), os.path.join(outdir, f'real_{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'padertorch')), reason='reference not installed in baseline/_ref')
def test_real_padertorch_trainer_world_size_two_gloo():
    """DistributedTrainer around the UNMODIFIED padertorch.Trainer (baseline/_ref), 2 ranks x 2 rounds per step, must
    equal the reference Trainer alone with virtual_minibatch_size=4 on the same examples (SURVEY.md section 8e:
    sum semantics, trainer.py:336,357,396-442) -- including replicas that start different and a dataset whose
    length is not divisible by the world size."""
    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_real_trainer_worker, args=(world, _free_port(), outdir), nprocs=world, join=True)
        pt = _import_reference()
        model = _toy_model(pt)
        single = pt.Trainer(model, os.path.join(outdir, 'single'), optimizer=pt.optimizer.SGD(lr=0.1),
                            stop_trigger=(2, 'iteration'), summary_trigger=(1, 'iteration'),
                            checkpoint_trigger=(100, 'iteration'), virtual_minibatch_size=4)
        single.train(_dict_examples(8), device='cpu', progress_bar=False)
        for rank in range(world):
            got = torch.load(os.path.join(outdir, f'real_{rank}.pt'))
            assert got['iteration'] == single.iteration == 2
            assert got['loss'] > 0      # the summed loss of the last step travelled in the last bucket
            for p, q in zip(got['params'], model.parameters()):
                torch.testing.assert_close(p, q.detach(), rtol=1e-5, atol=1e-6)
