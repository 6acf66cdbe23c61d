#!/usr/bin/env python
"""Deep-clustering forward / backward at geometries other than the compile-time one (513 bins, E = 20, K = 2): run with
B2S_DC_RING=0 / 1 to compare the six-warp frame kernel with the ring kernel on the run-time-geometry instances."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402
from kernel_bench import time_graph  # noqa: E402
from padertorch_b200 import review  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
for (B, M, E, K, F) in ((16, 253, 20, 2, 513), (16, 503, 20, 2, 257), (16, 253, 20, 3, 513), (16, 503, 20, 3, 257),
                        (16, 253, 16, 3, 513), (16, 400, 20, 4, 257)):
    n = 6
    emb = [torch.nn.functional.normalize(torch.randn(B, M, E, F, device=dev), dim=2) for _ in range(n)]
    tm = [torch.nn.functional.one_hot(torch.randint(0, K, (B, M, F), device=dev), K).permute(0, 1, 3, 2).float().contiguous()
          for _ in range(n)]
    ms = time_graph(lambda i: (lambda: review.dc_losses_per_example(emb[i], tm[i])), n)
    nbytes = B * M * F * (E + K) * 4
    from padertorch_b200.ops.losses.source_separation import DcProblem
    from padertorch_b200._workspace import meta_tensor

    def bwd(i):
        rows = [[M, b * M * E * F, b * M * K * F, b * M * E * F] for b in range(B)]
        meta = meta_tensor(rows, dev, cache_key=('geom', B, M, E, K, F))
        problem = DcProblem(emb[i], tm[i], meta, B, M, F, E, K, (E * F, F, 1), (K * F, F, 1), emb[i].numel(),
                            [(0, emb[i].numel(), emb[i].shape)])
        loss, gram = problem.forward()
        g = torch.ones_like(loss)
        return lambda: problem.backward(gram, g)
    ms_b = time_graph(bwd, n)
    print(f'RING={os.environ.get("B2S_DC_RING", "1")} B={B} M={M} E={E} K={K} F={F}: forward {ms * 1e3:7.1f} us '
          f'{nbytes / ms / 1e6:8.1f} GB/s   backward {ms_b * 1e3:7.1f} us {(nbytes + B * M * F * E * 4) / ms_b / 1e6:8.1f} GB/s',
          flush=True)
    del emb, tm
