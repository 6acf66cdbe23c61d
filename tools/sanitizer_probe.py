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

# ---- frame-staged TMA kernels (PIT-SSE, deep clustering), pair statistics, target preparation
for K, F, lengths in ((2, 513, [11, 7, 3]), (3, 257, [9, 5])):
    masks = [torch.rand(T, K, F, device=dev, requires_grad=True) for T in lengths]
    yab = [torch.rand(T, F, device=dev) for T in lengths]
    xab = [torch.rand(T, K, F, device=dev) for T in lengths]
    cpd = [torch.rand(T, K, F, device=dev) * 2 - 1 for T in lengths]
    out = b2s.review.pit_review_losses(masks, yab, xab, cpd)
    (out['pit_mse_loss'] + out['pit_ips_loss']).backward()
    out1 = b2s.review.pit_review_losses(masks, yab, xab)
    out1['pit_mse_loss'].backward()
for E, K, F, lengths in ((20, 2, 513, [6, 4, 1]), (7, 3, 65, [5, 9])):
    emb = [torch.nn.functional.normalize(torch.randn(T, E, F, device=dev), dim=1).requires_grad_(True) for T in lengths]
    tm = [torch.nn.functional.one_hot(torch.randint(0, K, (T, F), device=dev), K).permute(0, 2, 1).float().contiguous()
          for T in lengths]
    b2s.review.dc_review_loss(emb, tm).backward()
s = torch.randn(3, 2, 4001, device=dev)
est = (s + 0.3 * torch.randn_like(s)).requires_grad_(True)
out = b2s.review.tasnet_losses(est, s, [4001, 3000, 4001])
(out['si-sdr'] + out['log-mse']).backward()
prep = b2s.review.prepare_pit_targets(s.sum(1), s, stft=stft)
torch.cuda.synchronize()
print('ok losses', float(out['si-sdr']), prep['X_abs'].shape)

# ---- round 2: pair-transform fused kernel (default, K = 2), 8 x 8 x 8 pipeline and warp-specialised form (opt-in), fused
# backward, ragged batches; ring inverse with several chunks per row, chunk boundaries, both layouts; ragged targets, K = 4
import os
B, K, T = 5, 2, 20000
s = (0.1 * rng.randn(B, K, T)).astype(np.float32); y = s.sum(1)
M = stft.samples_to_frames(T)
lengths = [T, T - 999, T // 2, 4096, T - 4]
yd, sd = torch.from_numpy(y).to(dev), torch.from_numpy(s).to(dev)
ya = stft.magnitude(yd)
ref = None
for env in ({}, {'B2S_FUSED_PAIR': '0'}, {'B2S_FUSED_WS': '1'}):
    os.environ.update(env)
    md = torch.rand(B, M, K, 513, device=dev, requires_grad=True)
    for ns in (None, lengths):
        loss, perm = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=ya, num_samples=ns)
        loss.sum().backward()
    for k in env:
        os.environ.pop(k)
torch.cuda.synchronize()
for rows, T2 in ((3, 5000), (70, 9000), (1300, 1500)):
    for rep in ('complex', 'concat'):
        st = b2s.ops.STFT(1024, 256, complex_representation=rep)
        x = (0.1 * torch.randn(rows, T2, device=dev)).requires_grad_(True)
        spec = st(x)
        z = st.inverse(spec)
        z.sum().backward()
prep = b2s.review.prepare_pit_targets(torch.randn(3, 7001, device=dev), torch.randn(3, 4, 7001, device=dev), stft=stft,
                                      num_samples=[7001, 5000, 2048])
torch.cuda.synchronize()
print('ok round 2', float(loss.sum()), float(z.abs().max()), prep['X_abs'].shape)
