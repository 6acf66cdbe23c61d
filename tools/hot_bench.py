#!/usr/bin/env python
"""Timing of the two headline kernels only (front-end |Y| and fused STFT->PIT) at the north-star shape,
for quick A/B runs of kernel variants selected by environment variables (B2S_FWD_CTAS, B2S_FUSED_CTAS ...).

    B2S_FUSED_CTAS=2 python tools/hot_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402
from padertorch_b200 import review  # noqa: E402
from tools.kernel_bench import peak_gbs, time_graph  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    peak = peak_gbs()
    tag = ' '.join(f'{k}={v}' for k, v in sorted(os.environ.items()) if k.startswith('B2S_'))
    torch.manual_seed(0)
    B, K, T, M, F = 64, 2, 64000, 253, 513
    stft = b2s.ops.STFT(1024, 256)
    n = 3
    ys = [0.1 * torch.randn(B, T, device=dev) for _ in range(n)]
    ss = [0.1 * torch.randn(B, K, T, device=dev) for _ in range(n)]
    masks = [torch.rand(B, M, K, F, device=dev) for _ in range(n)]
    yabs = [stft.magnitude(y) for y in ys]

    def show(name, ms, bytes_):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        print(f'[{tag}] {name:34s} {ms * 1e3:8.1f} us {gbs:8.1f} GB/s {100 * gbs / peak:5.1f} %', flush=True)
        return ms

    a = show('stft |Y|', time_graph(lambda i: (lambda: stft.magnitude(ys[i])), n, iters=200), B * (4 * T + 4 * M * F))
    show('stft complex', time_graph(lambda i: (lambda: stft(ys[i])), n, iters=200), B * (4 * T + 8 * M * F))
    b = show('fused (reads |Y|)',
             time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                     observation_abs=yabs[i])), n, iters=200),
             B * (4 * T * (1 + K) + 4 * M * F * K))
    show('fused (recomputes |Y|)',
         time_graph(lambda i: (lambda: review.stft_mask_pit_step(ys[i], ss[i], masks[i], stft=stft)), n, iters=200),
         B * (4 * T * (1 + K) + 4 * M * F * K))
    total = B * (4 * T * (2 + K) + 4 * M * F * (1 + K))
    show('path = stft |Y| + fused', a + b, total)


if __name__ == '__main__':
    main()
