"""racecheck bisect: python tools/race_probe.py <case>  (pair | pair-ragged | old | old-ragged | ws | bwd | bwd-ragged)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
case = sys.argv[1]
if case.startswith('old') or case.startswith('bwd'):
    os.environ['B2S_FUSED_PAIR'] = '0'
if case.startswith('ws'):
    os.environ['B2S_FUSED_WS'] = '1'
import numpy as np, torch
import padertorch_b200 as b2s
dev = torch.device('cuda:0')
rng = np.random.RandomState(0)
stft = b2s.ops.STFT(1024, 256)
B, K, T = 5, 2, 20000
s = (0.1 * rng.randn(B, K, T)).astype(np.float32); y = s.sum(1)
M = stft.samples_to_frames(T)
lengths = [T, T - 999, T // 2, 4096, T - 4] if case.endswith('ragged') else None
yd, sd = torch.from_numpy(y).to(dev), torch.from_numpy(s).to(dev)
ya = stft.magnitude(yd)
md = torch.rand(B, M, K, 513, device=dev, requires_grad=case.startswith('bwd'))
loss, perm = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=ya, num_samples=lengths)
if case.startswith('bwd'):
    loss.sum().backward()
torch.cuda.synchronize()
print('ok', case, float(loss.sum()))
