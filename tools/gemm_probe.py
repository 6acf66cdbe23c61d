#!/usr/bin/env python
"""Parity + timing of the tcgen05 projection GEMM (b2s_linear_forward) at the PIT model's shapes.
    python tools/gemm_probe.py [--quick]
Run under `timeout`: a wrong barrier protocol hangs instead of failing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import padertorch_b200 as b2s  # noqa: E402
from tools.kernel_bench import time_graph  # noqa: E402


def main():
    quick = '--quick' in sys.argv
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = [(128, 128, 32), (256, 128, 64), (300, 200, 100), (1000, 1200, 1200)] if quick else \
        [(128, 128, 32), (300, 200, 100), (8096, 1200, 1200), (8096, 1026, 1200), (8096, 2400, 513), (8096, 2400, 1200)]
    for m, n, k in shapes:
        x = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        b = torch.randn(n, device=dev)
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        for precision in ('fp32', 'tf32'):
            for act in (None, 'relu', 'sigmoid'):
                y = b2s.ops.linear(x, w, b, activation=act, precision=precision)
                want = ref if act is None else (torch.relu(ref) if act == 'relu' else torch.sigmoid(ref))
                err = float((y.double() - want).abs().max() / want.abs().max())
                print(f'm={m} n={n} k={k} {precision} act={act}: max err / max |ref| = {err:.2e}', flush=True)
        err32 = float((torch.nn.functional.linear(x, w, b).double() - ref).abs().max() / ref.abs().max())
        print(f'   torch fp32 (cuBLAS) against float64: {err32:.2e}')
        if m >= 1000:
            flops = 2.0 * m * n * k
            for precision in ('fp32', 'tf32'):
                xl = b2s.ops.linear.__globals__['tf32_split'](x) if precision == 'fp32' else None
                from padertorch_b200.ops.linear import linear_forward
                ms = time_graph(lambda i: (lambda: linear_forward(x, w, b, 'relu', precision, x_lo=xl)), 1, iters=40)
                print(f'   b2s {precision}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s useful '
                      f'({(3 if precision == "fp32" else 1) * flops / ms / 1e9:8.1f} issued)')
            ms = time_graph(lambda i: (lambda: torch.relu(torch.nn.functional.linear(x, w, b))), 1, iters=40)
            print(f'   torch fp32 linear + relu: {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s')
            torch.backends.cuda.matmul.allow_tf32 = True
            ms = time_graph(lambda i: (lambda: torch.relu(torch.nn.functional.linear(x, w, b))), 1, iters=40)
            print(f'   torch tf32 linear + relu: {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s')
            torch.backends.cuda.matmul.allow_tf32 = False


if __name__ == '__main__':
    main()
