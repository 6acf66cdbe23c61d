#!/usr/bin/env python
"""NCCL check of the multi-GPU glue, run as one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/nccl_probe.py

1. Equivalence: a small mask estimator + the PIT-SSE loss kernels; every rank back-propagates its own shard
   (items rank, rank + G, ...), gradients are summed with padertorch_b200.parallel.allreduce_gradients; rank 0
   also processes ALL shards sequentially (the reference's virtual_minibatch_size = G on one GPU) and compares.
2. Cost of the exchange step: bucketed all-reduce of 101.3 MB of fp32 gradients (the PIT BLSTM's size), device
   timed, max over ranks.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402
from padertorch_b200 import parallel  # noqa: E402


def real_trainer_check(rank, world, dev):
    import tempfile
    import warnings
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'padertorch')):
        if rank == 0:
            print('[nccl_probe] reference not installed in baseline/_ref: real-Trainer check skipped', flush=True)
        return
    for path in (os.path.join(ROOT, 'oracle', 'ref_standins'), ref):
        if path not in sys.path:
            sys.path.insert(0, path)
    warnings.filterwarnings('ignore')
    import padertorch as pt
    from padertorch.contrib.examples.source_separation.pit.model import PermutationInvariantTrainingModel
    b2s.patch_padertorch(pt)
    stft = pt.ops.STFT(1024, 256)

    def examples(n):
        out = []
        g = torch.Generator().manual_seed(7)
        for i in range(n):
            T = 16000 - 800 * (i % 3)
            s = 0.1 * torch.randn(2, T, generator=g)
            Y = stft(s.sum(0).to(dev))
            X = stft(s.to(dev)).transpose(0, 1)
            out.append(dict(Y_abs=[Y.abs().cpu()], X_abs=[X.abs().cpu()],
                            cos_phase_difference=[torch.cos(torch.angle(Y[:, None, :]) - torch.angle(X)).cpu()]))
        return out

    def model():
        torch.manual_seed(3)
        return PermutationInvariantTrainingModel(F=513, recurrent_layers=1, units=32, K=2, dropout_input=0.,
                                                 dropout_hidden=0., dropout_linear=0.)

    data = examples(2 * world + 1)           # not divisible by the world size: the tail group is dropped
    steps = 2
    kwargs = dict(optimizer=pt.optimizer.SGD(lr=0.05), loss_weights={'pit_mse_loss': 1.0, 'pit_ips_loss': 0.5},
                  stop_trigger=(steps, 'iteration'), summary_trigger=(1, 'iteration'), checkpoint_trigger=(1000, 'iteration'))
    Trainer = parallel.distributed_trainer_class(pt.Trainer)
    with tempfile.TemporaryDirectory() as tmp:
        mine = model()
        trainer = Trainer(mine, os.path.join(parallel.rank_storage_dir(tmp), 'dist'),
                          virtual_minibatch_size=parallel.rounds_per_rank(world), **kwargs)
        trainer.train(parallel.ShardedDataset(data), device=dev.index, progress_bar=False)
        if rank == 0:
            single = model()
            reference_trainer = pt.Trainer(single, os.path.join(tmp, 'single'), virtual_minibatch_size=world, **kwargs)
            reference_trainer.train(data[:2 * world], device=dev.index, progress_bar=False)
            worst = max(float((p - q).abs().max() / q.abs().max().clamp_min(1e-12))
                        for p, q in zip(mine.parameters(), single.parameters()))
            print(f'[nccl_probe] real padertorch.Trainer, {world} GPU(s) x 1 round vs 1 GPU x virtual_minibatch_size {world}, '
                  f'{steps} optimizer steps: max relative parameter difference {worst:.2e}; summed loss of the last step '
                  f'{trainer.last_loss_sum:.6f}', flush=True)
            assert worst < 1e-4, worst
    b2s.unpatch_padertorch()


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dev = torch.device('cuda', torch.cuda.current_device())
    dist.init_process_group('nccl', device_id=dev)
    K, F, M, items = 2, 513, 40, 2 * world

    def make_model():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Linear(F, 256), torch.nn.ReLU(), torch.nn.Linear(256, K * F)).to(dev)

    def example(i):
        g = torch.Generator(device='cpu').manual_seed(100 + i)
        y = torch.rand(M, F, generator=g).to(dev)
        x = torch.rand(M, K, F, generator=g).to(dev)
        return y, x

    def loss_of(model, i):
        y, x = example(i)
        masks = torch.sigmoid(model(torch.log1p(y))).view(M, K, F)
        return b2s.review.pit_review_losses([masks], [y], [x])['pit_mse_loss']

    model = make_model()
    parallel.broadcast_parameters(model)
    total = torch.zeros((), device=dev)
    for i in parallel.shard_for_rank(range(items), rank, world):
        loss = loss_of(model, i)
        loss.backward()                      # gradients accumulate (sum) over the rounds, trainer.py:426-428
        total += loss.detach()
    parallel.allreduce_gradients(model.parameters(), extra=[total])
    if rank == 0:
        single = make_model()
        want = torch.zeros((), device=dev)
        for i in range(items):
            loss = loss_of(single, i)
            loss.backward()
            want += loss.detach()
        worst = max(float((p.grad - q.grad).abs().max() / q.grad.abs().max())
                    for p, q in zip(model.parameters(), single.parameters()))
        print(f'[nccl_probe] world {world}: summed loss {float(total):.6f} vs single-process {float(want):.6f}; '
              f'max relative gradient difference {worst:.2e}', flush=True)
        assert abs(float(total) - float(want)) <= 1e-5 * abs(float(want)) and worst < 1e-5

    # ---- 1b. the REAL padertorch.Trainer (baseline/_ref) under DistributedTrainer: G GPUs x 1 round per step must equal
    # the unmodified Trainer on one GPU with virtual_minibatch_size = G on the same examples
    real_trainer_check(rank, world, dev)

    # ---- cost of the exchange step
    numel = 25_324_626                       # PIT BLSTM (F = 513), SURVEY.md appendix B
    params = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in (numel // 3, numel // 3, numel - 2 * (numel // 3))]
    for p in params:
        p.grad = torch.ones_like(p)
    for _ in range(3):
        parallel.allreduce_gradients(params)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        parallel.allreduce_gradients(params)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = numel * 4 / 1e9
        print(f'[nccl_probe] all-reduce of {gb * 1e3:.1f} MB of gradients in {parallel.DEFAULT_BUCKET_BYTES >> 20} MB '
              f'buckets: {float(ms):.3f} ms (max over {world} ranks), algorithm bandwidth {gb / (float(ms) * 1e-3):.0f} GB/s',
              flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
