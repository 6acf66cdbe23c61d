#!/usr/bin/env python
"""The deep-clustering loss at config 3's fixed-length shape (batch 16 x 253 frames, E = 20, K = 2), forward + backward,
three times -- for ncu:  ncu --set full -k regex:'dc_gram_ring|dc_backward_frame' -s 2 -c 2 -o gpurun_out/prof python tools/dc_probe.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import padertorch_b200 as b2s
dev = torch.device('cuda:0')
torch.manual_seed(0)
B, M, E, K, F = 16, 253, 20, 2, 513
emb = torch.nn.functional.normalize(torch.randn(B, M, E, F, device=dev), dim=2).requires_grad_(True)
tm = torch.nn.functional.one_hot(torch.randint(0, K, (B, M, F), device=dev), K).permute(0, 1, 3, 2).float().contiguous()
for _ in range(3):
    loss = b2s.review.dc_losses_per_example(emb, tm)
    loss.sum().backward()
torch.cuda.synchronize()
print('loss', loss[:3].tolist())
