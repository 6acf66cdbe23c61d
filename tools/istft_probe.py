#!/usr/bin/env python
"""The inverse transform at the headline shape (64 rows x 253 frames -> 64 x 64000 samples), three launches -- for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'istft' -s 1 -c 1 -o gpurun_out/prof python tools/istft_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    B, T = 64, 64000
    stft = b2s.ops.STFT(1024, 256)
    y = 0.1 * torch.randn(B, T, device=dev)
    spec = stft(y)
    for _ in range(3):
        z = stft.inverse(spec)
    torch.cuda.synchronize()
    print('round trip error', float((z[..., :T] - y).abs().max()))


if __name__ == '__main__':
    main()
