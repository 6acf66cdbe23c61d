#!/usr/bin/env python
"""The two kernels of the bench step, three times each at the headline shape -- for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:'stft_pit_pair|stft_pit_fused|stft1024_warp' -s 2 -c 2 \
        -o gpurun_out/prof python tools/fused_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402
from padertorch_b200 import review  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    B, K, T, M, F = 64, 2, 64000, 253, 513
    stft = b2s.ops.STFT(1024, 256)
    y = 0.1 * torch.randn(B, T, device=dev)
    s = 0.1 * torch.randn(B, K, T, device=dev)
    mask = torch.rand(B, M, K, F, device=dev)
    for _ in range(3):
        yabs = stft.magnitude(y)
        loss, perm = review.stft_mask_pit_step(None, s, mask, stft=stft, observation_abs=yabs)
    torch.cuda.synchronize()
    print('loss[0:3]', loss[:3].tolist(), 'perm[0:3]', perm[:3].tolist())


if __name__ == '__main__':
    main()
