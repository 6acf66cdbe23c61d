import os, sys
sys.path.insert(0, '/root/repo')
import torch
import padertorch_b200 as b2s
from padertorch_b200 import review
dev = torch.device('cuda:0')
B, K, T, M, F = 64, 2, 64000, 253, 513
stft = b2s.ops.STFT(1024, 256)
sets = []
for i in range(3):
    s = 0.1 * torch.randn(B, K, T, device=dev); y = s.sum(1); m = torch.rand(B, M, K, F, device=dev)
    sets.append((s, stft.magnitude(y), m))
for rep in range(4):
    for s, ya, m in sets:
        print('--- launch', rep, file=sys.stderr)
        review.stft_mask_pit_step(None, s, m, stft=stft, observation_abs=ya)
torch.cuda.synchronize()
