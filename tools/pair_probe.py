#!/usr/bin/env python
"""TasNet forward (batch 64 x 2 x 4 s): statistics + loss set as two launches against b2s_pair_stats_loss_set (one
launch); run with B2S_PAIR_PIPE=0 / 1 for the statistics loop without / with the software pipeline."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import torch  # noqa: E402

from kernel_bench import time_graph, peak_gbs  # noqa: E402
from padertorch_b200 import _lib  # noqa: E402
from padertorch_b200._workspace import meta_tensor  # noqa: E402
from padertorch_b200.ops.losses import _pairs  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
K, T = 2, 64000
kinds = [_lib.LOSS_SI_SDR, _lib.LOSS_LOG_MSE, _lib.LOSS_LOG1P_MSE]
reductions = [_lib.REDUCE_MEAN, _lib.REDUCE_SUM, _lib.REDUCE_SUM]
for B in (64, 32):
    n = 8
    ss = [0.1 * torch.randn(B, K, T, device=dev) for _ in range(n)]
    est = [s + 0.3 * torch.randn_like(s) for s in ss]
    meta = meta_tensor([[T, b * K * T, b * K * T] for b in range(B)], dev, cache_key=('pp', B, K, T))
    problems = [_pairs.PairProblem(est[i], ss[i], meta, B, 1, K, T, T, T, covers_all=True) for i in range(n)]
    nbytes = B * 2 * 4 * K * T

    def two(i):
        def fn():
            st = problems[i].stats()
            return problems[i].loss_set(st, kinds, reductions)
        return fn
    for name, make in (('stats only', lambda i: problems[i].stats), ('two launches', two),
                       ('one launch', lambda i: (lambda: problems[i].stats_loss_set(kinds, reductions, one_launch=True)))):
        ms = time_graph(make, n)
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(f'B={B} PIPE={os.environ.get("B2S_PAIR_PIPE", "1")} {name:14s} {ms * 1e3:7.1f} us {gbs:8.1f} GB/s '
              f'{100 * gbs / peak_gbs():5.1f} %', flush=True)
