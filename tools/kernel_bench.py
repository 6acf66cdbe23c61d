#!/usr/bin/env python
"""Per-kernel timing of every hot-path kernel family at BASELINE.json's shapes, with the algorithmic
bytes of SURVEY.md section 8(d) and the fraction of the measured HBM peak.  Each kernel is replayed from a
one-launch CUDA graph over rotating input sets larger than L2; CUDA events on the launching stream.

    python tools/kernel_bench.py [--out profiles/kernels.json]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402
from padertorch_b200 import review  # noqa: E402


def peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs'])
    return 6650.0


def time_graph(make_fn, n_sets, iters=60):
    """make_fn(i) -> callable running ONE kernel on input set i.  Returns mean ms per replay."""
    graphs, keep = [], []
    for i in range(n_sets):
        fn = make_fn(i)
        keep.append(fn)            # buffers owned by the closure must outlive the graph replays
        fn()                       # warm-up: plans, workspaces
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        graphs.append(g)
    for i in range(2 * n_sets):
        graphs[i % n_sets].replay()
    # CUDA event timestamps tick at ~2 us on this platform: time bursts of replays, not single ones
    burst = 4 * n_sets
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
             for _ in range(max(5, iters // burst))]
    for e0, e1 in pairs:
        e0.record()
        for i in range(burst):
            graphs[i % n_sets].replay()
        e1.record()
    torch.cuda.synchronize()
    return statistics.median(e0.elapsed_time(e1) for e0, e1 in pairs) / burst


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--out', default=None)
    args = parser.parse_args()
    dev = torch.device('cuda:0')
    peak = peak_gbs()
    results = []

    def record(name, ms, bytes_, note=''):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        results.append(dict(kernel=name, ms=ms, algorithmic_bytes=bytes_, gbs=gbs, frac_of_measured_hbm=gbs / peak,
                            note=note))
        print(f'{name:34s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f} %  {note}', flush=True)

    torch.manual_seed(0)
    B, K, T, M, F = 64, 2, 64000, 253, 513
    stft = b2s.ops.STFT(1024, 256)
    n = 3
    ys = [0.1 * torch.randn(B, T, device=dev) for _ in range(n)]
    ss = [0.1 * torch.randn(B, K, T, device=dev) for _ in range(n)]
    masks = [torch.rand(B, M, K, F, device=dev) for _ in range(n)]

    # ---- STFT front-end / transforms (config: batch 64 x 4 s)
    record('stft |Y| (abs epilogue)', time_graph(lambda i: (lambda: stft.magnitude(ys[i])), n), B * (4 * T + 4 * M * F))
    record('stft log1p|Y|', time_graph(lambda i: (lambda: stft.magnitude(ys[i], log1p=True)), n), B * (4 * T + 4 * M * F))
    record('stft complex', time_graph(lambda i: (lambda: stft(ys[i])), n), B * (4 * T + 8 * M * F))
    specs = [stft(y) for y in ys]
    record('istft', time_graph(lambda i: (lambda: stft.inverse(specs[i])), n), B * (4 * T + 8 * M * F))
    gspec = [torch.view_as_real(s).contiguous() for s in specs]
    from padertorch_b200 import _lib
    from padertorch_b200.ops import _stft as S
    plan = stft._plan(dev)

    def stft_bwd(i):
        out = torch.empty(B, T, device=dev)
        def run():
            rc = _lib.load().b2s_stft_backward(plan.handle, gspec[i].data_ptr(), B, M, 0, 768, T, out.data_ptr(),
                                               _lib.ptr(S._scratch(plan, B, M, dev)), _lib.stream_of(dev))
            assert rc == 0
        return run
    record('stft backward (adjoint)', time_graph(stft_bwd, n), B * (4 * T + 8 * M * F))

    def istft_bwd(i):
        out = torch.empty(B, M, F, 2, device=dev)
        def run():
            rc = _lib.load().b2s_istft_backward(plan.handle, ys[i].data_ptr(), B, T, 768, M, 0, out.data_ptr(),
                                                _lib.stream_of(dev))
            assert rc == 0
        return run
    record('istft backward (adjoint)', time_graph(istft_bwd, n), B * (4 * T + 8 * M * F))

    # ---- fused loss, both |Y| variants
    yabs = [stft.magnitude(y) for y in ys]
    record('fused STFT->PIT (reads |Y|)',
           time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft, observation_abs=yabs[i])), n),
           B * (4 * T * (1 + K) + 4 * M * F * K))
    record('fused STFT->PIT (recomputes |Y|)',
           time_graph(lambda i: (lambda: review.stft_mask_pit_step(ys[i], ss[i], masks[i], stft=stft)), n),
           B * (4 * T * (1 + K) + 4 * M * F * K))

    # ---- batched target preparation on the device
    record('prepare_pit_targets (|Y|, |X|, cpd)',
           time_graph(lambda i: (lambda: review.prepare_pit_targets(ys[i], ss[i], stft=stft)), n),
           B * (4 * T * (1 + K) + 4 * M * F * (1 + 2 * K)), 'one kernel: transforms, magnitudes and phase term in registers')

    # ---- un-fused PIT-SSE (targets materialised), forward and backward, single and dual
    xabs = [stft.magnitude(s).transpose(1, 2).contiguous() for s in ss]
    cpd = [torch.rand(B, M, K, F, device=dev) * 2 - 1 for _ in range(n)]
    record('pit_sse forward', time_graph(lambda i: (lambda: review.pit_losses_per_example(masks[i], yabs[i], xabs[i])), n),
           B * 4 * M * F * (2 * K + 1))
    record('pit_sse forward dual (mse+ips)',
           time_graph(lambda i: (lambda: review.pit_losses_per_example(masks[i], yabs[i], xabs[i], cpd[i])), n),
           B * 4 * M * F * (3 * K + 1))
    from padertorch_b200.ops.losses import _sse

    def sse_bwd(i, dual):
        problem, _ = _sse.padded_problem(masks[i], yabs[i], xabs[i], cpd[i] if dual else None, None, dual)
        loss, perm, _ = problem.forward()
        g = torch.ones_like(loss)
        return lambda: problem.backward(perm, g)
    record('pit_sse backward', time_graph(lambda i: sse_bwd(i, False), n), B * 4 * M * F * (3 * K + 1))
    record('pit_sse backward dual', time_graph(lambda i: sse_bwd(i, True), n), B * 4 * M * F * (4 * K + 1))

    # ---- time-domain pair statistics (config 4 shape per GPU: batch 32 x 2 x 4 s; here batch 64)
    est = [s + 0.3 * torch.randn_like(s) for s in ss]
    num = [T] * B
    record('tasnet_losses forward (3 losses)',
           time_graph(lambda i: (lambda: review.tasnet_losses(est[i], ss[i], num)), n), B * 2 * 4 * K * T,
           '1 statistics pass + 1 launch for the three PIT losses and their batch means')
    from padertorch_b200.ops.losses import _pairs
    from padertorch_b200._workspace import meta_tensor

    def pair_bwd(i):
        rows = [[T, b * K * T, b * K * T] for b in range(B)]
        meta = meta_tensor(rows, dev, cache_key=('kb', B, K, T))
        problem = _pairs.PairProblem(est[i], ss[i], meta, B, 1, K, T, T, T, covers_all=True)
        stats = problem.stats()
        loss, perm = problem.loss(stats, _lib.LOSS_SI_SDR, 0, -1.0, _lib.REDUCE_MEAN, True)
        g = torch.ones_like(loss)
        return lambda: problem.backward(stats, _lib.LOSS_SI_SDR, 0, -1.0, _lib.REDUCE_MEAN, True, perm, g)
    record('si-sdr PIT backward', time_graph(pair_bwd, n), B * 3 * 4 * K * T)

    # ---- deep clustering (config 3: batch 16, E = 20, K = 2; fixed 4 s here, ragged in tests)
    Bd, E = 16, 20
    emb = [torch.nn.functional.normalize(torch.randn(Bd, M, E, F, device=dev), dim=2) for _ in range(n)]
    tm = [torch.nn.functional.one_hot(torch.randint(0, K, (Bd, M, F), device=dev), K).permute(0, 1, 3, 2).float().contiguous()
          for _ in range(n)]
    record('dc forward (Gram)', time_graph(lambda i: (lambda: review.dc_losses_per_example(emb[i], tm[i])), n),
           Bd * 4 * M * F * (E + K))
    from padertorch_b200.ops.losses.source_separation import DcProblem

    def dc_bwd(i):
        rows = [[M, b * M * E * F, b * M * K * F, b * M * E * F] for b in range(Bd)]
        meta = meta_tensor(rows, dev, cache_key=('kbdc', Bd, M, E, K, F))
        problem = DcProblem(emb[i], tm[i], meta, Bd, M, F, E, K, (E * F, F, 1), (K * F, F, 1), emb[i].numel(),
                            [(0, emb[i].numel(), emb[i].shape)])
        loss, gram = problem.forward()
        g = torch.ones_like(loss)
        return lambda: problem.backward(gram, g)
    record('dc backward', time_graph(dc_bwd, n), Bd * (4 * M * F * (E + K) + 4 * M * F * E))

    # ---- 3 speakers, 8 s (config 5 shape), batch 32
    B5, K5, T5, M5 = 32, 3, 128000, 503
    s5 = [0.1 * torch.randn(B5, K5, T5, device=dev) for _ in range(2)]
    y5 = [s.sum(1) for s in s5]
    m5 = [torch.rand(B5, M5, K5, F, device=dev) for _ in range(2)]
    ya5 = [stft.magnitude(y) for y in y5]
    record('fused STFT->PIT K=3, 8 s, batch 32',
           time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, s5[i], m5[i], stft=stft, observation_abs=ya5[i])), 2),
           B5 * (4 * T5 * (1 + K5) + 4 * M5 * F * K5))

    out = dict(peak_gbs=peak, gpu=torch.cuda.get_device_name(0), results=results)
    if args.out:
        with open(args.out, 'w') as fd:
            json.dump(out, fd, indent=1)


if __name__ == '__main__':
    main()
