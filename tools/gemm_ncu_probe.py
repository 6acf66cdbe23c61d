#!/usr/bin/env python
"""Three launches of the tcgen05 projection at the PIT model's Linear(1200, 1200) shape (batch 32 x 253 frames) for
    ncu --set full --clock-control none -k regex:linear_umma -s 1 -c 2 -o gpurun_out/prof_gemm python tools/gemm_ncu_probe.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from padertorch_b200.ops.linear import linear_forward, tf32_split  # noqa: E402

x = torch.randn(8096, 1200, device='cuda')
w = torch.randn(1200, 1200, device='cuda') / 35
b = torch.randn(1200, device='cuda')
xl = tf32_split(x)
for precision in ('fp32', 'tf32', 'fp32', 'tf32'):
    linear_forward(x, w, b, 'relu', precision, x_lo=xl if precision == 'fp32' else None)
torch.cuda.synchronize()
print('ok')
