#!/usr/bin/env python
"""Front-end / fused kernel time versus batch size: separates the fixed per-launch cost from the per-frame cost."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import padertorch_b200 as b2s  # noqa: E402
from padertorch_b200 import review  # noqa: E402
from tools.kernel_bench import time_graph  # noqa: E402

dev = torch.device('cuda:0')
tag = ' '.join(f'{k}={v}' for k, v in sorted(os.environ.items()) if k.startswith('B2S_'))
stft = b2s.ops.STFT(1024, 256)
K, T, M, F = 2, 64000, 253, 513
for B in (8, 32, 64, 128, 256):
    n = 3 if B <= 128 else 2
    ys = [0.1 * torch.randn(B, T, device=dev) for _ in range(n)]
    a = time_graph(lambda i: (lambda: stft.magnitude(ys[i])), n, iters=100)
    line = f'[{tag}] B={B:4d} stft|Y| {a * 1e3:7.1f} us ({a * 1e6 / (B * M):6.3f} ns/frame)'
    if '--fused' in sys.argv:
        ss = [0.1 * torch.randn(B, K, T, device=dev) for _ in range(n)]
        masks = [torch.rand(B, M, K, F, device=dev) for _ in range(n)]
        yabs = [stft.magnitude(y) for y in ys]
        b = time_graph(lambda i: (lambda: review.stft_mask_pit_step(None, ss[i], masks[i], stft=stft,
                                                                    observation_abs=yabs[i])), n, iters=100)
        line += f'  fused {b * 1e3:7.1f} us ({b * 1e6 / (B * M):6.3f} ns/pos)'
    print(line, flush=True)
