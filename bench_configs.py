"""BASELINE.json's configs 2-5 behind ``bench.py --config {train, dc, tasnet, pit3}`` (config 1 / the headline is
bench.py itself).  Same JSON contract as the headline line: device-timed value with inputs resident in HBM, `e2e`
through the public call with pinned host buffers, `roofline` of the dominant kernel against the measured HBM peak,
`cpu_baseline` = the unmodified reference (baseline/_ref) on the host cores on a bounded sample.

  pit3    config 5: 3 speakers (3! permutations), 8 s / 16 kHz, batch 32 per GPU: front-end + fused STFT->mask->PIT
  dc      config 3: deep-clustering affinity loss (E = 20) forward + backward, batch 16 of 2..8 s (ragged)
  tasnet  config 4: time-domain PIT losses (si-sdr, log-mse, log1p-mse) forward + si-sdr backward on [32, 2, 64000] per
          GPU, plus the sum-all-reduce of a ConvTasNet's 34.9 MB gradient (NCCL, issued right after backward)
  train   config 2: one full training step of the PIT BLSTM mask estimator (25.3 M parameters) at batch 32 x 4 s:
          |Y| front-end -> log1p -> 3 x bidirectional LSTM(600) (cuDNN) -> tcgen05 projections -> fused STFT->mask->PIT
          loss (forward + backward kernels) -> backward -> bucketed NCCL gradient exchange overlapped with it -> Adam
"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

import bench as B

SIZE, SHIFT, BINS = 1024, 256, 513


def _frames(samples):
    return -(-(samples + 2 * (SIZE - SHIFT) - SIZE + SHIFT) // SHIFT)


def _time_graphs(graphs, steps, distributed, device):
    import torch.distributed as dist
    for i in range(2 * len(graphs)):
        graphs[i % len(graphs)].replay()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        graphs[i % len(graphs)].replay()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    if distributed:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def _traffic(key):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), or None."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'traffic.json')
    try:
        with open(path) as fd:
            return json.load(fd).get(key)
    except (OSError, ValueError):
        return None


def _capture(fn, sets):
    graphs, keep = [], []
    for data in sets:
        fn(data)
    torch.cuda.synchronize()
    for data in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(fn(data))
        graphs.append(g)
    return graphs, keep


def _burst_ms(graphs, bursts=6):
    n = len(graphs)
    for i in range(2 * n):
        graphs[i % n].replay()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(bursts)]
    per = 4 * n
    for e0, e1 in pairs:
        e0.record()
        for i in range(per):
            graphs[i % n].replay()
        e1.record()
    torch.cuda.synchronize()
    return statistics.median(e0.elapsed_time(e1) for e0, e1 in pairs) / per


def _e2e(step_host, steps, distributed, device):
    import torch.distributed as dist
    for i in range(3):
        step_host(i)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step_host(i)
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = float(t.item())
    return s


def _cpu_model():
    try:
        with open('/proc/cpuinfo') as fd:
            return next(l.split(':', 1)[1].strip() for l in fd if l.startswith('model name'))
    except Exception:
        return ''


def _emit(line):
    B.emit(line)


def _base_line(args, world, value, ms_per_step, config, **extra):
    line = {'metric': 'utterances/sec', 'value': value, 'unit': 'utt/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config}
    line.update(extra)
    return line


# ================================================================================================ pit3 (config 5)
class Pit3:
    name = 'pit3'
    batch, sources, samples = 32, 3, 128000
    frames = _frames(128000)
    workload = ('fused STFT->mask->PIT-loss path, 3 speakers (3! = 6 permutations), batch 32 x 8 s x 16 kHz, '
                'STFT(1024,256) (503 frames x 513 bins); masks synthetic U(0,1), mask network excluded')

    def config(self, world):
        return {'workload': self.workload, 'batch_per_gpu': self.batch, 'samples': self.samples,
                'sources': self.sources, 'frames': self.frames, 'bins': BINS,
                'l2': 'inputs larger than L2: 3 rotating input sets of 164 MB each',
                'parallelism': f'{world} independent shard(s), no data-path collective'}

    def host_set(self, seed):
        g = torch.Generator().manual_seed(seed)
        s = 0.1 * torch.randn(self.batch, self.sources, self.samples, generator=g)
        return dict(y=s.sum(1), s=s, masks=torch.rand(self.batch, self.frames, self.sources, BINS, generator=g))

    def bytes_front(self):
        return 4 * self.samples + 4 * self.frames * BINS

    def bytes_loss(self):
        return 4 * self.samples * (1 + self.sources) + 4 * self.frames * BINS * self.sources

    def run_ours(self, args, rank, world, device, distributed):
        import padertorch_b200 as b2s
        from padertorch_b200 import review
        stft = b2s.ops.STFT(SIZE, SHIFT)
        sets = [{k: v.to(device) for k, v in self.host_set(100 * rank + i).items()} for i in range(3)]

        def step(d):
            y_abs = stft.magnitude(d['y'])
            return review.stft_mask_pit_step(None, d['s'], d['masks'], stft=stft, observation_abs=y_abs)

        graphs, _ = _capture(step, sets)
        ms = _time_graphs(graphs, args.steps, distributed, device)
        yabs = [stft.magnitude(d['y']) for d in sets]
        idx = {id(d): i for i, d in enumerate(sets)}
        g_loss, _ = _capture(lambda d: review.stft_mask_pit_step(None, d['s'], d['masks'], stft=stft,
                                                                 observation_abs=yabs[idx[id(d)]]), sets)
        g_front, _ = _capture(lambda d: stft.magnitude(d['y']), sets)
        loss_ms, front_ms = _burst_ms(g_loss), _burst_ms(g_front)
        host = [{k: v.pin_memory() for k, v in self.host_set(100 * rank + i).items()} for i in range(2)]
        out_loss = torch.empty(self.batch).pin_memory()

        def step_host(i):
            d = {k: v.to(device, non_blocking=True) for k, v in host[i % 2].items()}
            loss, _ = step(d)
            out_loss.copy_(loss, non_blocking=True)

        e2e_steps = max(5, min(args.steps, 20))
        e2e_s = _e2e(step_host, e2e_steps, distributed, device)
        if rank != 0:
            return
        peak, peak_kind = B.measured_peaks()
        achieved = self.bytes_loss() * self.batch / (loss_ms * 1e-3) / 1e9
        cpu = self.cpu_baseline(2)
        _emit(_base_line(
            args, world, world * self.batch * args.steps / (ms * 1e-3), ms / args.steps, self.config(world),
            e2e={'value': world * self.batch * e2e_steps / e2e_s, 'unit': 'utt/s',
                 'h2d_bytes_per_step': sum(v.numel() * 4 for v in host[0].values()), 'd2h_bytes_per_step': self.batch * 4,
                 'steps': e2e_steps},
            gpu_launches=2 * args.steps,
            roofline={'bound': 'hbm', 'kernel': 'stft_pit_fused_kernel<3>', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                      'frac': achieved / peak, 'traffic': None, 'peak_kind': peak_kind, 'kernel_ms': loss_ms,
                      'algorithmic_bytes_per_launch': self.bytes_loss() * self.batch},
            path_roofline={'front_end_kernel_ms': front_ms,
                           'front_end_frac': self.bytes_front() * self.batch / (front_ms * 1e-3) / 1e9 / peak,
                           'frac': (self.bytes_front() + self.bytes_loss()) * self.batch / (ms / args.steps * 1e-3) / 1e9 / peak},
            cpu_baseline=cpu, launch='CUDA graph replay, one graph (2 kernel nodes) per input set'))

    def reference_step(self, pt, d, n, stft):
        with torch.no_grad():
            y_abs = stft(d['y'][:n]).abs()
            x_abs = stft(d['s'][:n]).abs().transpose(1, 2)
            return [pt.ops.losses.pit_loss(d['masks'][b] * y_abs[b][:, None, :], x_abs[b], axis=-2) for b in range(n)]

    def cpu_baseline(self, reps, n=8):
        pt = B.import_reference()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        d = self.host_set(1)
        if pt is None:
            return {'value': None, 'unit': 'utt/s', 'cores': cores, 'kind': 'port', 'sample': 'reference not installed'}
        stft = pt.ops.STFT(SIZE, SHIFT)
        self.reference_step(pt, d, n, stft)
        t0 = time.perf_counter()
        for _ in range(reps):
            self.reference_step(pt, d, n, stft)
        per = (time.perf_counter() - t0) / reps
        return {'value': n / per, 'unit': 'utt/s', 'cores': cores, 'kind': 'reference',
                'sample': f'{reps} x {n} utterances of the batch through the unmodified reference (pt.ops.STFT + per-example '
                          f'pit_loss loop over 3! permutations), {per:.2f} s per pass, {_cpu_model()}'}


# ================================================================================================ dc (config 3)
class Dc:
    name = 'dc'
    batch, e_dim, sources = 16, 20, 2
    workload = ('deep-clustering affinity loss forward + backward (dc_gram_ring_kernel + dc_backward_frame_kernel, length-balanced chunks), '
                'batch 16 of 2..8 s (lengths uniform in [32000, 128000] samples -> 128..503 frames x 513 bins), E = 20, '
                "K = 2, embeddings in the model's 't e f' layout, binary target masks; embedding network excluded")

    def lengths(self, seed):
        rng = np.random.RandomState(seed)
        return sorted((int(n) for n in rng.randint(32000, 128001, size=self.batch)), reverse=True)

    def config(self, world):
        return {'workload': self.workload, 'batch_per_gpu': self.batch, 'embedding_dim': self.e_dim, 'sources': self.sources,
                'bins': BINS, 'l2': 'inputs larger than L2: 3 rotating ragged input sets of ~230 MB each',
                'parallelism': f'{world} independent shard(s), no data-path collective'}

    def host_set(self, seed):
        g = torch.Generator().manual_seed(seed)
        emb, tgt = [], []
        for n in self.lengths(seed):
            m = _frames(n)
            e = torch.nn.functional.normalize(torch.randn(m, self.e_dim, BINS, generator=g), dim=-2)
            hot = torch.randint(0, self.sources, (m, BINS), generator=g)
            emb.append(e)
            tgt.append(torch.nn.functional.one_hot(hot, self.sources).permute(0, 2, 1).float().contiguous())
        return emb, tgt

    def algorithmic_bytes(self, emb):
        c = self.e_dim + self.sources
        fwd = sum(4 * e.shape[0] * BINS * c for e in emb)
        bwd = fwd + sum(4 * e.shape[0] * BINS * self.e_dim for e in emb)
        return fwd, bwd

    def run_ours(self, args, rank, world, device, distributed):
        from padertorch_b200 import review
        sets = []
        for i in range(3):
            emb, tgt = self.host_set(100 * rank + i)
            sets.append(([e.to(device).requires_grad_(True) for e in emb], [t.to(device) for t in tgt]))

        def step(d):
            loss = review.dc_review_loss(d[0], d[1])
            grads = torch.autograd.grad(loss, d[0])
            return loss, grads

        graphs, _ = _capture(step, sets)
        ms = _time_graphs(graphs, args.steps, distributed, device)
        g_fwd, _ = _capture(lambda d: review.dc_review_loss([e.detach() for e in d[0]], d[1]), sets)
        fwd_ms = _burst_ms(g_fwd)
        fwd_b, bwd_b = self.algorithmic_bytes(sets[0][0])
        host = [self.host_set(100 * rank + i) for i in range(2)]
        host = [([e.pin_memory() for e in emb], [t.pin_memory() for t in tgt]) for emb, tgt in host]
        out = torch.empty(1).pin_memory()

        def step_host(i):
            emb, tgt = host[i % 2]
            d = ([e.to(device, non_blocking=True).requires_grad_(True) for e in emb],
                 [t.to(device, non_blocking=True) for t in tgt])
            loss, _ = step(d)
            out.copy_(loss.detach().reshape(1), non_blocking=True)

        e2e_steps = max(5, min(args.steps, 20))
        e2e_s = _e2e(step_host, e2e_steps, distributed, device)
        if rank != 0:
            return
        peak, peak_kind = B.measured_peaks()
        achieved = fwd_b / (fwd_ms * 1e-3) / 1e9
        step_gbs = (fwd_b + bwd_b) / (ms / args.steps * 1e-3) / 1e9
        _emit(_base_line(
            args, world, world * self.batch * args.steps / (ms * 1e-3), ms / args.steps, self.config(world),
            e2e={'value': world * self.batch * e2e_steps / e2e_s, 'unit': 'utt/s',
                 'h2d_bytes_per_step': sum(t.numel() * 4 for t in host[0][0] + host[0][1]), 'd2h_bytes_per_step': 4,
                 'steps': e2e_steps},
            gpu_launches=2 * args.steps,
            roofline={'bound': 'hbm', 'kernel': 'dc_gram_ring_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                      'frac': achieved / peak, 'traffic': _traffic('dc_gram_ring_kernel@config3'), 'peak_kind': peak_kind,
                      'kernel_ms': fwd_ms,
                      'algorithmic_bytes_per_launch': fwd_b},
            step_roofline={'achieved': step_gbs, 'frac': step_gbs / peak, 'unit': 'GB/s',
                           'algorithmic_bytes_per_step': fwd_b + bwd_b},
            cpu_baseline=self.cpu_baseline(2), launch='CUDA graph replay, forward + backward per input set'))

    def cpu_baseline(self, reps, n=4):
        pt = B.import_reference()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        if pt is None:
            return {'value': None, 'unit': 'utt/s', 'cores': cores, 'kind': 'port', 'sample': 'reference not installed'}
        import einops
        emb, tgt = self.host_set(1)
        emb = [e.requires_grad_(True) for e in emb[:n]]

        def run():
            losses = [pt.ops.losses.deep_clustering_loss(einops.rearrange(e, 't e f -> (t f) e'),
                                                         einops.rearrange(t, 't k f -> (t f) k'))
                      for e, t in zip(emb, tgt[:n])]
            torch.mean(torch.stack(losses)).backward()
        run()
        t0 = time.perf_counter()
        for _ in range(reps):
            run()
        per = (time.perf_counter() - t0) / reps
        return {'value': n / per, 'unit': 'utt/s', 'cores': cores, 'kind': 'reference',
                'sample': f'{reps} x the {n} longest utterances of the batch through the unmodified reference loop of tcl/dc.py:76-84 '
                          f'(forward + backward), {per:.2f} s per pass, {_cpu_model()}'}


# ================================================================================================ tasnet (config 4)
class Tasnet:
    name = 'tasnet'
    batch, sources, samples = 32, 2, 64000
    grad_params = 8734017          # ConvTasNet (N = 256, 8 x 4 blocks), SURVEY.md appendix B
    workload = ('TasNet.loss: PIT si-sdr / log-mse / log1p-mse forward + si-sdr backward on [32, 2, 64000] estimates per GPU '
                '(batch 256 box-wide at 8 GPUs), followed by the sum-all-reduce of the 34.9 MB ConvTasNet gradient (NCCL); '
                'separator network excluded')

    def config(self, world):
        return {'workload': self.workload, 'batch_per_gpu': self.batch, 'samples': self.samples, 'sources': self.sources,
                'l2': 'inputs larger than L2: 8 rotating input sets of 32.8 MB each',
                'parallelism': f'{world} shard(s); one gradient all-reduce ({self.grad_params * 4 / 1e6:.1f} MB, sum) per step'}

    def host_set(self, seed):
        g = torch.Generator().manual_seed(seed)
        s = 0.1 * torch.randn(self.batch, self.sources, self.samples, generator=g)
        est = s[:, torch.randperm(self.sources, generator=g)] + 0.05 * torch.randn(s.shape, generator=g)
        return dict(est=est - est.mean(-1, keepdim=True), s=s)

    def run_ours(self, args, rank, world, device, distributed):
        import torch.distributed as dist
        from padertorch_b200 import review
        sets = [{k: v.to(device) for k, v in self.host_set(100 * rank + i).items()} for i in range(8)]
        for d in sets:
            d['est'].requires_grad_(True)
        lengths = [self.samples] * self.batch
        gradient = torch.zeros(self.grad_params, device=device)       # the separator's flat gradient buckets
        buckets = list(gradient.split(8 << 20))                       # 32 MiB buckets

        def losses(d):
            out = review.tasnet_losses(d['est'], d['s'], lengths)
            grad, = torch.autograd.grad(out['si-sdr'], d['est'])
            return out, grad

        # forward + backward of one input set = one CUDA graph (5 kernel nodes); the all-reduces follow eagerly
        step_graphs, step_out = _capture(losses, sets)

        def step(i, reduce=True):
            step_graphs[i % 8].replay()
            if reduce and distributed:
                works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True) for b in buckets]
                for w in works:
                    w.wait()
            return step_out[i % 8][0]['si-sdr']

        def timed(reduce):
            for i in range(max(args.warmup, 3)):
                step(i, reduce)
            torch.cuda.synchronize()
            if distributed:
                dist.barrier()
            torch.cuda.synchronize()
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            for i in range(args.steps):
                step(i, reduce)
            end.record()
            torch.cuda.synchronize()
            ms = start.elapsed_time(end)
            if distributed:
                t = torch.tensor([ms], device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        ms = timed(True)
        ms_local = timed(False) if distributed else ms
        g_fwd, _ = _capture(lambda d: review.tasnet_losses(d['est'].detach(), d['s'], lengths), sets)
        fwd_ms = _burst_ms(g_fwd)
        host = [{k: v.pin_memory() for k, v in self.host_set(100 * rank + i).items()} for i in range(2)]
        out = torch.empty(1).pin_memory()

        def step_host(i):   # the step's inputs land in the graph's own input set, then the same graph replay
            with torch.no_grad():
                for k, v in host[i % 2].items():
                    sets[i % 8][k].copy_(v, non_blocking=True)
            out.copy_(step(i).detach().reshape(1), non_blocking=True)

        e2e_steps = max(5, min(args.steps, 20))
        e2e_s = _e2e(step_host, e2e_steps, distributed, device)
        if rank != 0:
            return
        peak, peak_kind = B.measured_peaks()
        fwd_b = 2 * 4 * self.sources * self.samples * self.batch
        achieved = fwd_b / (fwd_ms * 1e-3) / 1e9
        _emit(_base_line(
            args, world, world * self.batch * args.steps / (ms * 1e-3), ms / args.steps, self.config(world),
            e2e={'value': world * self.batch * e2e_steps / e2e_s, 'unit': 'utt/s',
                 'h2d_bytes_per_step': sum(v.numel() * 4 for v in host[0].values()), 'd2h_bytes_per_step': 4, 'steps': e2e_steps},
            gpu_launches=3 * args.steps,
            roofline={'bound': 'hbm', 'kernel': 'pair_stats_kernel + pair_loss_set_kernel (TasNet forward)', 'achieved': achieved,
                      'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None, 'peak_kind': peak_kind,
                      'kernel_ms': fwd_ms, 'algorithmic_bytes_per_launch': fwd_b},
            collective={'op': 'all_reduce(sum)', 'bytes': self.grad_params * 4, 'ranks': world,
                        'ms_per_step_with': ms / args.steps, 'ms_per_step_without': ms_local / args.steps,
                        'exposed_us': (ms - ms_local) / args.steps * 1e3},
            cpu_baseline=self.cpu_baseline(2),
            launch='CUDA graph replay per input set (forward + autograd backward), NCCL all-reduces eager; device-timed'))

    def cpu_baseline(self, reps, n=8):
        pt = B.import_reference()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        if pt is None:
            return {'value': None, 'unit': 'utt/s', 'cores': cores, 'kind': 'port', 'sample': 'reference not installed'}
        from padertorch.contrib.examples.source_separation.tasnet.model import TasNet
        d = self.host_set(1)
        est = d['est'][:n].clone().requires_grad_(True)

        def run():
            out = TasNet.loss(None, {'s': d['s'][:n], 'num_samples': [self.samples] * n}, {'out': est})
            out['si-sdr'].backward()
        run()
        t0 = time.perf_counter()
        for _ in range(reps):
            run()
        per = (time.perf_counter() - t0) / reps
        return {'value': n / per, 'unit': 'utt/s', 'cores': cores, 'kind': 'reference',
                'sample': f'{reps} x {n} utterances through the unmodified TasNet.loss (tasnet/model.py:154-176) + si-sdr backward, '
                          f'{per:.2f} s per pass, {_cpu_model()}'}


# ================================================================================================ train (config 2)
class Train:
    name = 'train'
    batch, sources, samples = 32, 2, 64000
    frames = _frames(64000)
    workload = ('one training step of the PIT BLSTM mask estimator (pit/model.py: 3 x bidirectional LSTM(600), Linear(1200,1200)+ReLU, '
                'Linear(1200,1026)+sigmoid; 25.3 M parameters) at batch 32 x 4 s x 16 kHz, 2 speakers: |Y| front-end kernel -> log1p -> '
                'cuDNN LSTM -> tcgen05 projections -> fused STFT->mask->PIT loss (forward + backward kernels) -> backward -> bucketed '
                'NCCL gradient exchange overlapped with backward (sum, 101.3 MB) -> Adam')

    def config(self, world):
        return {'workload': self.workload, 'batch_per_gpu': self.batch, 'samples': self.samples, 'sources': self.sources,
                'frames': self.frames, 'bins': BINS, 'l2': 'n/a: the step streams 101 MB of parameters, their gradients and optimizer state',
                'parallelism': f'dp{world}: one replica per GPU, gradients summed with bucketed all-reduces (32 MiB buckets) as backward produces them'}

    def host_set(self, seed):
        g = torch.Generator().manual_seed(seed)
        s = 0.1 * torch.randn(self.batch, self.sources, self.samples, generator=g)
        return dict(y=s.sum(1), s=s)

    def build(self, device):
        import padertorch_b200 as b2s
        torch.manual_seed(0)

        class MaskEstimator(torch.nn.Module):
            def __init__(self, K):
                super().__init__()
                self.K = K
                self.blstm = torch.nn.LSTM(BINS, 600, 3, bidirectional=True, batch_first=True)
                self.linear1 = b2s.ops.FusedLinear(1200, 1200, activation='relu')
                self.linear2 = b2s.ops.FusedLinear(1200, BINS * K, activation='sigmoid')

            def forward(self, y_abs):
                h, _ = self.blstm(torch.log1p(y_abs))
                b, m, _ = h.shape
                h = self.linear2(self.linear1(h.reshape(b * m, -1)))
                return h.view(b, m, self.K, BINS)
        return MaskEstimator(self.sources).to(device)

    def run_ours(self, args, rank, world, device, distributed):
        import torch.distributed as dist
        import padertorch_b200 as b2s
        from padertorch_b200 import parallel, review
        stft = b2s.ops.STFT(SIZE, SHIFT)
        model = self.build(device)
        n_params = sum(p.numel() for p in model.parameters())
        parallel.broadcast_parameters(model)
        exchange = parallel.GradientExchange(model.parameters(), rounds=1)
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
        sets = [{k: v.to(device) for k, v in self.host_set(100 * rank + i).items()} for i in range(3)]

        def step(d):
            y_abs = stft.magnitude(d['y'])
            masks = model(y_abs)
            loss, _ = review.stft_mask_pit_step(None, d['s'], masks, stft=stft, observation_abs=y_abs)
            total = loss.sum()                      # summed over examples and ranks, as the reference Trainer does
            exchange.add_loss(total)
            total.backward()
            exchange.finish()
            optimizer.step()
            exchange.zero_grad()
            return total.detach()

        for i in range(max(args.warmup, 3)):
            step(sets[i % 3])
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(args.steps):
            last = step(sets[i % 3])
        end.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(end)
        if distributed:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())

        # where the time goes: CUDA-event spans of the phases of one step (eager)
        def phases(d):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record(); y_abs = stft.magnitude(d['y'])
            ev[1].record(); masks = model(y_abs)
            ev[2].record(); loss, _ = review.stft_mask_pit_step(None, d['s'], masks, stft=stft, observation_abs=y_abs)
            total = loss.sum()
            ev[3].record(); total.backward()
            ev[4].record(); exchange.finish(); optimizer.step(); exchange.zero_grad()
            ev[5].record()
            torch.cuda.synchronize()
            return [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
        spans = np.median(np.array([phases(sets[i % 3]) for i in range(5)]), axis=0)

        host = [{k: v.pin_memory() for k, v in self.host_set(100 * rank + i).items()} for i in range(2)]
        out = torch.empty(1).pin_memory()

        def step_host(i):
            d = {k: v.to(device, non_blocking=True) for k, v in host[i % 2].items()}
            out.copy_(step(d).reshape(1), non_blocking=True)

        e2e_steps = max(5, min(args.steps, 20))
        e2e_s = _e2e(step_host, e2e_steps, distributed, device)
        if rank != 0:
            return
        m_rows = self.batch * self.frames
        gemm_flops = 2.0 * m_rows * (1200 * 1200 + 1200 * BINS * self.sources)
        _emit(_base_line(
            args, world, world * self.batch * args.steps / (ms * 1e-3), ms / args.steps, self.config(world),
            e2e={'value': world * self.batch * e2e_steps / e2e_s, 'unit': 'utt/s',
                 'h2d_bytes_per_step': sum(v.numel() * 4 for v in host[0].values()), 'd2h_bytes_per_step': 4, 'steps': e2e_steps},
            gpu_launches=6 * args.steps,
            parameters=n_params, gradient_bytes=exchange.nbytes(),
            phases_ms={'front_end': float(spans[0]), 'network_forward (cuDNN LSTM + tcgen05 projections)': float(spans[1]),
                       'fused_loss_forward': float(spans[2]), 'backward (incl. fused loss backward, all-reduces in flight)': float(spans[3]),
                       'exchange_wait + adam': float(spans[4])},
            roofline={'bound': 'tensor', 'kernel': 'linear_umma_kernel<3> (the two projections, forward)', 'achieved': None,
                      'peak': None, 'unit': 'TFLOP/s', 'frac': None, 'traffic': None,
                      'note': f'{gemm_flops / 1e9:.1f} GFLOP useful per step in the projections; per-kernel tensor-pipe utilisation in '
                              'profiles/r2_gemm_tcgen05.txt; the step is bound by the cuDNN LSTM (library code)'},
            cpu_baseline=self.cpu_baseline(), launch='python eager (autograd, cuDNN, NCCL), device-timed',
            last_loss=float(last)))

    def cpu_baseline(self):
        pt = B.import_reference()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        if pt is None:
            return {'value': None, 'unit': 'utt/s', 'cores': cores, 'kind': 'port', 'sample': 'reference not installed'}
        from padertorch.contrib.examples.source_separation.pit.model import PermutationInvariantTrainingModel
        n = 2
        d = self.host_set(1)
        stft = pt.ops.STFT(SIZE, SHIFT)
        with torch.no_grad():
            Y = stft(d['y'][:n])
            X = stft(d['s'][:n]).transpose(1, 2)
        batch = dict(Y_abs=[Y[b].abs() for b in range(n)], X_abs=[X[b].abs() for b in range(n)],
                     cos_phase_difference=[torch.cos(torch.angle(Y[b][:, None, :]) - torch.angle(X[b])) for b in range(n)])
        torch.manual_seed(0)
        model = PermutationInvariantTrainingModel(F=BINS, recurrent_layers=3, units=600, K=self.sources)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)

        def run():
            opt.zero_grad()
            review = model.review(batch, model(batch))
            review['losses']['pit_mse_loss'].backward()
            opt.step()
        run()
        t0 = time.perf_counter()
        run()
        per = time.perf_counter() - t0
        return {'value': n / per, 'unit': 'utt/s', 'cores': cores, 'kind': 'reference',
                'sample': f'1 x one training step of the unmodified PermutationInvariantTrainingModel (F=513, 3 x BLSTM(600)) on {n} '
                          f'utterances with precomputed spectra, {per:.2f} s, {_cpu_model()}'}


WORKLOADS = {w.name: w for w in (Pit3, Dc, Tasnet, Train)}


def run(args, rank, world, local_rank):
    workload = WORKLOADS[args.config]()
    if args.impl == 'reference':
        if rank != 0:
            return
        reps = max(1, min(args.steps, 3))
        cpu = workload.cpu_baseline() if args.config == 'train' else workload.cpu_baseline(reps)
        value = cpu['value']
        _emit({'impl': 'reference', 'metric': 'utterances/sec', 'value': value, 'unit': 'utt/s', 'n_gpus': args.gpus,
               'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak',
               'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload.config(world), 'cpu_baseline': cpu,
               'e2e': {'value': value, 'unit': 'utt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0})
        return
    if args.impl == 'reference-gpu':
        raise SystemExit('--impl reference-gpu exists for the headline config (pit) only')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: padertorch_b200 has no CPU fallback')
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        B.nccl_info_to_stderr()
        dist.init_process_group('nccl', device_id=device)
    try:
        workload.run_ours(args, rank, world, device, distributed)
    finally:
        if distributed:
            import torch.distributed as dist
            dist.destroy_process_group()
