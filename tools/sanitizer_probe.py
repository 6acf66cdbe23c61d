import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import padertorch_b200 as b2s
from oracle import path as OP
dev = torch.device('cuda:0')
rng = np.random.RandomState(0)
stft = b2s.ops.STFT(1024, 256)
for B, K, T in ((2, 2, 9000), (1, 3, 5001)):
    s = (0.1 * rng.randn(B, K, T)).astype(np.float32); y = s.sum(1)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, K, 513).astype(np.float32)
    want_loss, want_perm, want_yabs = OP.stft_mask_pit_step(torch.from_numpy(y), torch.from_numpy(s), torch.from_numpy(masks))
    yd, sd, md = (torch.from_numpy(a).to(dev) for a in (y, s, masks))
    ya = stft.magnitude(yd)
    Y = stft(yd)
    back = stft.inverse(Y)
    loss, perm = b2s.review.stft_mask_pit_step(yd, sd, md, stft=stft)
    loss2, perm2 = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=ya)
    torch.cuda.synchronize()
    print('ok', float((ya.cpu() - want_yabs).abs().max()), float((back[..., :T].cpu() - torch.from_numpy(y)).abs().max()),
          float((loss.cpu() - want_loss).abs().max()), float((loss2.cpu() - want_loss).abs().max()), perm.tolist() == [list(p) for p in want_perm])
