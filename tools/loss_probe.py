#!/usr/bin/env python
"""Runs the time-domain and deep-clustering loss paths a few times (for an ncu launch list:
ncu --metrics gpu__time_duration.sum --csv python tools/loss_probe.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from padertorch_b200 import review  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
B, K, T, M, F, E = 64, 2, 64000, 253, 513, 20
s = 0.1 * torch.randn(B, K, T, device=dev)
est = (s + 0.3 * torch.randn_like(s)).requires_grad_(True)
for _ in range(3):
    out = review.tasnet_losses(est, s, [T] * B)
    out['si-sdr'].backward()
    est.grad = None
Bd = 16
emb = torch.nn.functional.normalize(torch.randn(Bd, M, E, F, device=dev), dim=2).requires_grad_(True)
tm = torch.nn.functional.one_hot(torch.randint(0, K, (Bd, M, F), device=dev), K).permute(0, 1, 3, 2).float().contiguous()
for _ in range(3):
    loss = review.dc_review_loss(emb, tm)
    loss.backward()
    emb.grad = None
torch.cuda.synchronize()
# PIT-SSE (dual) forward / backward and the target preparation at the bench shape
import padertorch_b200 as b2s  # noqa: E402
stft = b2s.ops.STFT(1024, 256)
prep = review.prepare_pit_targets(s.sum(1), s, stft=stft)
masks = torch.rand(B, M, K, F, device=dev, requires_grad=True)
for _ in range(2):
    out = review.pit_review_losses(masks, prep['Y_abs'], prep['X_abs'], prep['cos_phase_difference'])
    (out['pit_mse_loss'] + out['pit_ips_loss']).backward()
    masks.grad = None
torch.cuda.synchronize()
