#!/usr/bin/env python
"""A/B timing of the pipeline shapes of the two headline kernels in ONE process (the libraries read
B2S_FUSED_VARIANT / B2S_FWD_VARIANT at every launch, i.e. at graph capture).  Every variant's results are
compared with the first variant's bit for bit resp. to 1e-6.

    python tools/variant_bench.py [--fused 1,0,2,3,4] [--fwd 0,1,2,3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402
from padertorch_b200 import review  # noqa: E402
from tools.kernel_bench import peak_gbs, time_graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--fused', default='0,1,2,3,4')
    ap.add_argument('--fwd', default='0,1,2,3')
    ap.add_argument('--ring', default='', help='B2S_FUSED_RING values to time (default shape), e.g. 0,1')
    ap.add_argument('--pair', default='', help='B2S_FUSED_PAIR values to time (1 = pair-transform kernel), e.g. 0,1,0,1')
    ap.add_argument('--ws', default='', help='B2S_FUSED_WS values to time (1 = warp-specialised kernel), e.g. 0,1,0,1')
    ap.add_argument('--ablate', default='', help='B2S_FUSED_ABLATE values to time (default shape)')
    ap.add_argument('--iters', type=int, default=240)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    peak = peak_gbs()
    torch.manual_seed(0)
    B, K, T, M, F = 64, 2, 64000, 253, 513
    stft = b2s.ops.STFT(1024, 256)
    n = 3
    ys = [0.1 * torch.randn(B, T, device=dev) for _ in range(n)]
    ss = [0.1 * torch.randn(B, K, T, device=dev) for _ in range(n)]
    masks = [torch.rand(B, M, K, F, device=dev) for _ in range(n)]
    os.environ['B2S_FWD_VARIANT'] = '0'
    yabs = [stft.magnitude(y) for y in ys]

    def show(name, ms, bytes_, note=''):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        print(f'{name:40s} {ms * 1e3:8.1f} us {gbs:8.1f} GB/s {100 * gbs / peak:5.1f} %  {note}', flush=True)

    ref = None
    for v in [int(t) for t in args.fwd.split(',') if t != '']:
        os.environ['B2S_FWD_VARIANT'] = str(v)
        out = stft.magnitude(ys[0])
        torch.cuda.synchronize()
        if ref is None:
            ref = out
        err = float((out - ref).abs().max() / ref.abs().max())
        ms = time_graph(lambda i: (lambda: stft.magnitude(ys[i])), n, iters=args.iters)
        show(f'front-end |Y| variant {v}', ms, B * (4 * T + 4 * M * F), f'max rel dev from first variant {err:.1e}')
    os.environ['B2S_FWD_VARIANT'] = '0'

    ref = None
    for v in [int(t) for t in args.fused.split(',') if t != '']:
        os.environ['B2S_FUSED_VARIANT'] = str(v)
        loss, perm = review.stft_mask_pit_step(None, ss[0], masks[0], stft=stft, observation_abs=yabs[0])
        torch.cuda.synchronize()
        if ref is None:
            ref = (loss.clone(), perm.clone())
        dl = float(((loss - ref[0]).abs() / ref[0].abs()).max())
        same = bool((perm == ref[1]).all())
        ms = time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                     observation_abs=yabs[i])), n, iters=args.iters)
        show(f'fused (reads |Y|) variant {v}', ms, B * (4 * T * (1 + K) + 4 * M * F * K),
             f'loss dev {dl:.1e}, permutations equal: {same}')
    os.environ.pop('B2S_FUSED_VARIANT', None)
    for r in [int(t) for t in args.ring.split(',') if t != '']:
        os.environ['B2S_FUSED_RING'] = str(r)
        loss, perm = review.stft_mask_pit_step(None, ss[0], masks[0], stft=stft, observation_abs=yabs[0])
        torch.cuda.synchronize()
        same = ref is None or (bool((perm == ref[1]).all()) and float(((loss - ref[0]).abs() / ref[0].abs()).max()) == 0.0)
        ms = time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                     observation_abs=yabs[i])), n, iters=args.iters)
        show(f'fused hop ring = {r}', ms, B * (4 * T * (1 + K) + 4 * M * F * K), f'bit-identical to the first variant: {same}')
    os.environ.pop('B2S_FUSED_RING', None)
    pair_ref = None
    for r in [int(t) for t in args.pair.split(',') if t != '']:
        os.environ['B2S_FUSED_PAIR'] = str(r)
        loss, perm = review.stft_mask_pit_step(None, ss[0], masks[0], stft=stft, observation_abs=yabs[0])
        torch.cuda.synchronize()
        if pair_ref is None:
            pair_ref = (loss.clone(), perm.clone())
        dl = float(((loss - pair_ref[0]).abs() / pair_ref[0].abs()).max())
        same = bool((perm == pair_ref[1]).all())
        ms = time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                     observation_abs=yabs[i])), n, iters=args.iters)
        show(f'fused pair transform = {r}', ms, B * (4 * T * (1 + K) + 4 * M * F * K),
             f'loss dev from first {dl:.1e}, permutations equal: {same}')
    os.environ.pop('B2S_FUSED_PAIR', None)
    ws_ref = None
    for r in [int(t) for t in args.ws.split(',') if t != '']:
        os.environ['B2S_FUSED_WS'] = str(r)
        loss, perm = review.stft_mask_pit_step(None, ss[0], masks[0], stft=stft, observation_abs=yabs[0])
        torch.cuda.synchronize()
        if ws_ref is None:
            ws_ref = (loss.clone(), perm.clone())
        dl = float(((loss - ws_ref[0]).abs() / ws_ref[0].abs()).max())
        same = bool((perm == ws_ref[1]).all())
        ms = time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                     observation_abs=yabs[i])), n, iters=args.iters)
        show(f'fused warp-specialised = {r}', ms, B * (4 * T * (1 + K) + 4 * M * F * K),
             f'loss dev from first {dl:.1e}, permutations equal: {same}')
    os.environ.pop('B2S_FUSED_WS', None)
    for a in [int(t) for t in args.ablate.split(',') if t != '']:
        os.environ['B2S_FUSED_ABLATE'] = str(a)
        ms = time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                     observation_abs=yabs[i])), n, iters=args.iters)
        show(f'fused ablate={a} (1 no SSE, 2 no sqrt, 4 no rows, 8 no frames)', ms,
             B * (4 * T * (1 + K) + 4 * M * F * K))
    os.environ.pop('B2S_FUSED_ABLATE', None)


if __name__ == '__main__':
    main()
