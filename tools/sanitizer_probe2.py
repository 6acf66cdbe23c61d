#!/usr/bin/env python
"""The kernels of the last session of round 2 on small inputs -- for compute-sanitizer (memcheck / racecheck):
dc_gram_ring_kernel with several chunks per example and wrap-around of its four-stage ring, balanced chunk slots (ragged and
equal lengths) in the Gram and backward kernels, the cluster fold of the pair statistics, b2s_pair_stats_loss_set."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import padertorch_b200 as b2s
from padertorch_b200 import _lib
from padertorch_b200._workspace import meta_tensor
from padertorch_b200.ops.losses import _pairs

dev = torch.device('cuda:0')
torch.manual_seed(0)
E, K, F = 20, 2, 513
for lengths in ([70, 33, 9], [24, 24], [300]):
    T = max(lengths)
    emb = torch.nn.functional.normalize(torch.randn(len(lengths), T, E, F, device=dev), dim=2).requires_grad_(True)
    tm = torch.nn.functional.one_hot(torch.randint(0, K, (len(lengths), T, F), device=dev), K).permute(0, 1, 3, 2).float().contiguous()
    loss = b2s.review.dc_review_loss(emb, tm, lengths)
    loss.backward()
    print('dc', lengths, float(loss))
for K_, F_, lens in ((2, 513, [30, 11, 25]), (3, 257, [9, 5])):
    Tm = max(lens)
    masks = torch.rand(len(lens), Tm, K_, F_, device=dev, requires_grad=True)
    out = b2s.review.pit_review_losses(masks, torch.rand(len(lens), Tm, F_, device=dev), torch.rand(len(lens), Tm, K_, F_, device=dev),
                                       torch.rand(len(lens), Tm, K_, F_, device=dev) * 2 - 1, lens)
    (out['pit_mse_loss'] + 0.5 * out['pit_ips_loss']).backward()
    print('pit review', lens, float(out['pit_mse_loss']))
B, K, T = 3, 2, 40000
s = torch.randn(B, K, T, device=dev)
est = (s + 0.3 * torch.randn_like(s)).requires_grad_(True)
out = b2s.review.tasnet_losses(est, s, [T, 30000, T])
(out['si-sdr'] + out['log-mse']).backward()
meta = meta_tensor([[T, b * K * T, b * K * T] for b in range(B)], dev)
problem = _pairs.PairProblem(est.detach(), s, meta, B, 1, K, T, T, T, covers_all=True)
kinds = [_lib.LOSS_SI_SDR, _lib.LOSS_LOG_MSE, _lib.LOSS_LOG1P_MSE]
reductions = [_lib.REDUCE_MEAN, _lib.REDUCE_SUM, _lib.REDUCE_SUM]
for _ in range(2):
    stats, loss, perm, mean = problem.stats_loss_set(kinds, reductions, one_launch=True)
torch.cuda.synchronize()
print('pairs', float(out['si-sdr']), mean.tolist())
