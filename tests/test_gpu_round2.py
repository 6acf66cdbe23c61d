"""GPU parity of the round-2 additions (all through the C ABI): pairwise loss matrices and their autograd
from the pair statistics, the device assignment, PIT with cross entropy / opaque callables / complex inputs.
Oracle = oracle/ (CPU, float64) and the reference's doctest values."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4


def dev():
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def b2s():
    import padertorch_b200
    return padertorch_b200


def _pairs(b2s):
    from oracle import losses as ol
    return [
        ('mse', b2s.ops.mse_loss, ol.mse_loss),
        ('log_mse', b2s.ops.log_mse_loss, ol.log_mse_loss),
        ('log1p_mse', b2s.ops.log1p_mse_loss, ol.log1p_mse_loss),
        ('sdr', b2s.ops.sdr_loss, ol.sdr_loss),
        ('si_sdr', b2s.ops.si_sdr_loss, ol.si_sdr_loss),
    ]


@pytest.mark.parametrize('shape,axis', [((2, 4000), 0), ((3, 2, 1777), 0), ((5, 900), 0), ((2, 3, 1500), 1)])
def test_pairwise_losses_regression_against_oracle(b2s, shape, axis):
    """compute_pairwise_losses (source_separation.py:127-241) for every regression loss: values and the
    gradient of a random linear functional of the matrix, against the float64 oracle."""
    from oracle import losses as ol
    rng = np.random.RandomState(hash((shape, axis)) % 2 ** 31)
    e = rng.randn(*shape)
    t = 0.7 * e + 0.5 * rng.randn(*shape)
    k = shape[axis]
    weights = rng.randn(k, k)
    for name, ours, theirs in _pairs(b2s):
        ed = torch.tensor(e, dtype=torch.float32, device=dev(), requires_grad=True)
        td = torch.tensor(t, dtype=torch.float32, device=dev())
        got = b2s.ops.losses.compute_pairwise_losses(ed, td, axis=axis, loss_fn=ours)
        assert got.shape == (k, k)
        (got * torch.tensor(weights, dtype=torch.float32, device=dev())).sum().backward()
        e64 = torch.tensor(e, dtype=torch.float64, requires_grad=True)
        want = ol.compute_pairwise_losses(e64, torch.tensor(t, dtype=torch.float64), axis=axis, loss_fn=theirs)
        (want * torch.tensor(weights)).sum().backward()
        np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=LOSS_RTOL, atol=1e-6,
                                   err_msg=name)
        scale = float(e64.grad.abs().max())
        err = float((ed.grad.cpu().double() - e64.grad).abs().max())
        assert err <= 1e-4 * scale, (name, err, scale)


@pytest.mark.parametrize('shape,axis', [((50, 2, 33), -2), ((3, 40, 17), 0), ((7, 4, 5, 6), 1)])
def test_pairwise_losses_torch_mse(b2s, shape, axis):
    """loss_fn = torch.nn.functional.mse_loss (the default): mean over every element of the pair."""
    from oracle import losses as ol
    rng = np.random.RandomState(3)
    e, t = rng.rand(*shape), rng.rand(*shape)
    k = shape[axis]
    w = rng.randn(k, k)
    ed = torch.tensor(e, dtype=torch.float32, device=dev(), requires_grad=True)
    got = b2s.ops.losses.compute_pairwise_losses(ed, torch.tensor(t, dtype=torch.float32, device=dev()), axis=axis)
    (got * torch.tensor(w, dtype=torch.float32, device=dev())).sum().backward()
    e64 = torch.tensor(e, requires_grad=True)
    want = ol.compute_pairwise_losses(e64, torch.tensor(t), axis=axis)
    (want * torch.tensor(w)).sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=LOSS_RTOL)
    assert float((ed.grad.cpu().double() - e64.grad).abs().max()) <= 1e-4 * float(e64.grad.abs().max())


def test_loss_matrix_doctest_values(b2s):
    """source_separation.py:262-274: -26 (optimal), -21 and [-11, -10, -0] (greedy)."""
    score = np.array([[11., 10, 0], [4, 5, 10], [6, 0, 5]])
    m = torch.tensor(-score, dtype=torch.float32, device=dev())
    f = b2s.ops.losses.pit_loss_from_loss_matrix
    assert float(f(m, reduction='sum', algorithm='optimal')) == -26.0
    assert float(f(m, reduction='sum', algorithm='greedy')) == -21.0
    np.testing.assert_array_equal(f(m, reduction=None, algorithm='greedy').cpu().numpy(), [-11., -10., -0.])
    loss, col = f(m, reduction='mean', return_permutation=True)
    np.testing.assert_array_equal(np.asarray(col), [1, 2, 0])
    assert abs(float(loss) + 26.0 / 3) < 1e-6
    with pytest.raises(ValueError):
        f(m, algorithm='fastest')
    with pytest.raises(ValueError):
        f(m, reduction='median')


@pytest.mark.parametrize('k', [1, 2, 3, 4, 5, 6, 7, 8])
def test_device_assignment_matches_scipy(b2s, k):
    """b2s_assign against scipy.optimize.linear_sum_assignment (the reference's solver, :288) on random
    matrices -- same optimum; on matrices with exact ties the lexicographically first col_ind."""
    import scipy.optimize
    rng = np.random.RandomState(k)
    f = b2s.ops.losses.pit_loss_from_loss_matrix
    for trial in range(12):
        m = rng.randn(k, k).astype(np.float32)
        got, col = f(torch.tensor(m, device=dev(), requires_grad=True), reduction='sum', return_permutation=True)
        rows, cols = scipy.optimize.linear_sum_assignment(m.astype(np.float64))
        want = m.astype(np.float64)[rows, cols].sum()
        assert sorted(np.asarray(col).tolist()) == list(range(k))
        np.testing.assert_allclose(float(got), want, rtol=1e-5, atol=1e-6)
        assert got.requires_grad      # the gather stays in autograd, on the device
    ties = np.zeros((k, k), dtype=np.float32)
    _, col = f(torch.tensor(ties, device=dev()), reduction='sum', return_permutation=True)
    np.testing.assert_array_equal(np.asarray(col), np.arange(k))
    if k >= 2:
        # brute force over all assignments: the first minimum in lexicographic order of col_ind
        m = rng.randint(0, 3, size=(k, k)).astype(np.float32)
        best = min(itertools.permutations(range(k)), key=lambda c: (sum(m[i, c[i]] for i in range(k)), c))
        _, col = f(torch.tensor(m, device=dev()), reduction='sum', return_permutation=True)
        np.testing.assert_array_equal(np.asarray(col), best)


def test_pit_cross_entropy_and_opaque_callable(b2s):
    """pit_loss doctests :70-73 (cross entropy -> 0.6931) and a callable the kernels know nothing about,
    both against the oracle's permutation loop."""
    from oracle import losses as ol
    T, K, F = 4, 2, 5
    est = torch.ones(T, K, F, device=dev())
    tgt = torch.zeros(T, F, dtype=torch.int64, device=dev())
    got = b2s.ops.pit_loss(est, tgt, 1, loss_fn=torch.nn.functional.cross_entropy)
    assert abs(float(got) - 0.6931) < 1e-4
    rng = np.random.RandomState(0)
    for k in (2, 3, 4):
        logits = rng.randn(6, k, 7)
        labels = rng.randint(0, k, size=(6, 7))
        got, perm = b2s.ops.pit_loss(torch.tensor(logits, dtype=torch.float32, device=dev()),
                                     torch.tensor(labels, device=dev()), 1,
                                     loss_fn=torch.nn.functional.cross_entropy, return_permutation=True)
        want, want_perm = ol.pit_loss(torch.tensor(logits), torch.tensor(labels), 1,
                                      loss_fn=torch.nn.functional.cross_entropy, return_permutation=True)
        assert tuple(perm) == tuple(want_perm)
        np.testing.assert_allclose(float(got), float(want), rtol=1e-5)

    def l1(a, b):
        return (a - b).abs().mean()
    e, t = rng.randn(3, 50), rng.randn(3, 50)
    got, perm = b2s.ops.pit_loss(torch.tensor(e, dtype=torch.float32, device=dev()),
                                 torch.tensor(t, dtype=torch.float32, device=dev()), 0, loss_fn=l1,
                                 return_permutation=True)
    want, want_perm = ol.pit_loss(torch.tensor(e), torch.tensor(t), 0, loss_fn=l1, return_permutation=True)
    assert tuple(perm) == tuple(want_perm)
    np.testing.assert_allclose(float(got), float(want), rtol=1e-5)


def test_complex_inputs_follow_the_reference_doctests(b2s):
    """regression.py:153-156: sdr_loss on complex signals (-11.9498, -20); mse_loss / SA-SDR via abs()."""
    from oracle import losses as ol
    a = torch.tensor([1, 2 + 3j, 4j], device=dev())
    b = torch.tensor([2, 3 + 3j, 5j], device=dev())
    assert abs(float(b2s.ops.sdr_loss(a, b)) + 11.9498) < 1e-3
    assert abs(float(b2s.ops.sdr_loss(a, a, soft_sdr_max=20)) + 20.0) < 1e-4
    rng = np.random.RandomState(1)
    e = rng.randn(2, 300) + 1j * rng.randn(2, 300)
    t = rng.randn(2, 300) + 1j * rng.randn(2, 300)
    ed, td = (torch.tensor(x, dtype=torch.complex64, device=dev()) for x in (e, t))
    np.testing.assert_allclose(float(b2s.ops.mse_loss(ed, td)), float(ol.mse_loss(torch.tensor(e), torch.tensor(t))),
                               rtol=LOSS_RTOL)
    np.testing.assert_allclose(float(b2s.ops.sdr_loss(ed, td)), float(ol.sdr_loss(torch.tensor(e), torch.tensor(t))),
                               rtol=LOSS_RTOL)


def test_sa_sdr_broadcast_and_empty(b2s):
    """ADVICE (round 1): a broadcastable target must not read out of bounds; empty input -> NaN, no crash."""
    from oracle import losses as ol
    rng = np.random.RandomState(2)
    e, t = rng.randn(3, 500), rng.randn(1, 500)
    got = b2s.ops.source_aggregated_sdr_loss(torch.tensor(e, dtype=torch.float32, device=dev()),
                                             torch.tensor(t, dtype=torch.float32, device=dev()))
    want = ol.source_aggregated_sdr_loss(torch.tensor(e), torch.tensor(t).expand(3, 500))
    np.testing.assert_allclose(float(got), float(want), rtol=LOSS_RTOL)
    empty = torch.zeros(0, 10, device=dev())
    assert torch.isnan(b2s.ops.source_aggregated_sdr_loss(empty, empty))
    assert b2s.ops.mse_loss(empty, empty, reduction=None).shape == (0,)


def test_pit_regression_loss_along_inner_axis(b2s):
    """pit_loss(axis != 0) with a regression loss: [B, K, T] permuted along axis 1."""
    from oracle import losses as ol
    rng = np.random.RandomState(5)
    e, t = rng.randn(4, 3, 800), rng.randn(4, 3, 800)
    got, perm = b2s.ops.pit_loss(torch.tensor(e, dtype=torch.float32, device=dev()),
                                 torch.tensor(t, dtype=torch.float32, device=dev()), 1,
                                 loss_fn=b2s.ops.si_sdr_loss, return_permutation=True)
    want, want_perm = ol.pit_loss(torch.tensor(e), torch.tensor(t), 1, loss_fn=ol.si_sdr_loss,
                                  return_permutation=True)
    assert tuple(perm) == tuple(want_perm)
    np.testing.assert_allclose(float(got), float(want), rtol=LOSS_RTOL)


# ------------------------------------------------------------------------------------------------ fused step
def _step_inputs(B, K, T, seed, eps=None):
    rng = np.random.RandomState(seed)
    s = (0.1 * rng.randn(B, K, T)).astype(np.float32)
    if eps is not None:   # near-tie stress of SURVEY.md section 8(d): s_k = s_1 + eps * N(0, 1)
        for k in range(1, K):
            s[:, k] = s[:, 0] + (eps * rng.randn(B, T)).astype(np.float32)
    y = s.sum(1)
    return y, s, rng


@pytest.mark.parametrize('k,seconds,recompute', [(2, 1.0, False), (2, 1.0, True), (3, 0.7, False), (1, 0.5, False),
                                                 (4, 0.4, True)])
def test_fused_step_backward_against_oracle(b2s, k, seconds, recompute):
    """b2s_stft_pit_backward: d (sum_b w_b loss_b) / d mask against torch autograd through the float64 oracle
    (pit/model.py:117-128 + source_separation.py:112-119), full-length and ragged batches."""
    from oracle import path as OP
    B, T = 3, int(16000 * seconds)
    y, s, rng = _step_inputs(B, k, T, 31 + k)
    stft = b2s.ops.STFT(1024, 256)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, k, 513).astype(np.float32)
    w = rng.rand(B).astype(np.float32) + 0.5
    yd, sd = torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev())
    md = torch.from_numpy(masks).to(dev()).requires_grad_(True)
    kwargs = dict(stft=stft) if recompute else dict(stft=stft, observation_abs=stft.magnitude(yd))
    loss, perm = b2s.review.stft_mask_pit_step(yd if recompute else None, sd, md, **kwargs)
    (loss * torch.from_numpy(w).to(dev())).sum().backward()
    m64 = torch.from_numpy(masks).double().requires_grad_(True)
    want_loss, want_perm, _ = OP.stft_mask_pit_step(torch.from_numpy(y).double(), torch.from_numpy(s).double(), m64)
    (want_loss * torch.from_numpy(w).double()).sum().backward()
    np.testing.assert_array_equal(perm.cpu().numpy(), np.asarray(want_perm))
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want_loss.detach().numpy(), rtol=LOSS_RTOL)
    scale = float(m64.grad.abs().max())
    err = float((md.grad.cpu().double() - m64.grad).abs().max())
    assert err <= 1e-4 * scale, (err, scale)
    # ragged: frames beyond an example's length get a zero gradient
    num_samples = [T, T - 1234, T - 4000]
    md2 = torch.from_numpy(masks).to(dev()).requires_grad_(True)
    loss_r, _ = b2s.review.stft_mask_pit_step(yd, sd, md2, stft=stft, num_samples=num_samples)
    loss_r.sum().backward()
    for b, n in enumerate(num_samples):
        m_b = stft.samples_to_frames(n)
        mb = torch.from_numpy(masks[b:b + 1, :m_b]).double().requires_grad_(True)
        w_loss, _, _ = OP.stft_mask_pit_step(torch.from_numpy(y[b:b + 1, :n]).double(),
                                             torch.from_numpy(s[b:b + 1, :, :n]).double(), mb)
        w_loss.sum().backward()
        got = md2.grad[b].cpu().double()
        assert float((got[:m_b] - mb.grad[0]).abs().max()) <= 1e-4 * float(mb.grad.abs().max())
        assert float(got[m_b:].abs().max() if m_b < M else 0.0) == 0.0


@pytest.mark.parametrize('B,K,T', [(64, 2, 64000), (32, 3, 128000)])
def test_fused_step_full_size_against_oracle(b2s, B, K, T):
    """The headline shapes of BASELINE.json (batch 64 x 4 s x 2 speakers; batch 32 x 8 s x 3 speakers) against
    the oracle DIRECTLY: permutations bit exact, losses and the front-end |Y| within 1e-4, and the gradient of
    the mean loss w.r.t. the masks."""
    from oracle import path as OP
    y, s, rng = _step_inputs(B, K, T, 7)
    stft = b2s.ops.STFT(1024, 256)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, K, 513).astype(np.float32)
    yd, sd = torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev())
    md = torch.from_numpy(masks).to(dev()).requires_grad_(True)
    y_abs = stft.magnitude(yd)
    loss, perm = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=y_abs)
    loss.mean().backward()
    mo = torch.from_numpy(masks).requires_grad_(True)
    want_loss, want_perm, want_yabs = OP.stft_mask_pit_step(torch.from_numpy(y), torch.from_numpy(s), mo)
    want_loss.mean().backward()
    np.testing.assert_array_equal(perm.cpu().numpy(), np.asarray(want_perm))
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want_loss.detach().numpy(), rtol=LOSS_RTOL)
    assert float((y_abs.cpu() - want_yabs).abs().max()) <= 1e-4 * float(want_yabs.abs().max())
    assert float((md.grad.cpu() - mo.grad).abs().max()) <= 1e-4 * float(mo.grad.abs().max())


@pytest.mark.parametrize('B,K,T,ragged', [(5, 2, 16000, False), (7, 2, 24000, True), (3, 1, 9000, False),
                                          (64, 2, 64000, False)])
def test_warp_specialised_fused_kernel_against_oracle(b2s, B, K, T, ragged, monkeypatch):
    """The opt-in warp-specialised form of the fused kernel (csrc/fused_ws.cuh, B2S_FUSED_WS=1: transform warps and
    SSE warps with re-split register file) against the oracle and against the default one-role kernel: permutations
    bit exact, losses within 1e-4; ragged batches; bit-identical reruns."""
    from oracle import path as OP
    y, s, rng = _step_inputs(B, K, T, 11)
    stft = b2s.ops.STFT(1024, 256)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, K, 513).astype(np.float32)
    lengths = [int(T - 1000 * (b % 4) - 3 * b) for b in range(B)] if ragged else None
    if ragged:   # |Y| comes from the padded batch: the padding must be silence, as in a collated batch
        for b, n in enumerate(lengths):
            y[b, n:] = 0
            s[b, :, n:] = 0
    yd, sd = torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev())
    md = torch.from_numpy(masks).to(dev())
    y_abs = stft.magnitude(yd)
    monkeypatch.delenv('B2S_FUSED_WS', raising=False)
    monkeypatch.setenv('B2S_FUSED_PAIR', '0')    # reference: the one-role 8 x 8 x 8 pipeline
    loss0, perm0 = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=y_abs, num_samples=lengths)
    monkeypatch.delenv('B2S_FUSED_PAIR', raising=False)
    lossp, permp = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=y_abs, num_samples=lengths)
    np.testing.assert_array_equal(permp.cpu().numpy(), perm0.cpu().numpy())   # default (pair kernel for K = 2)
    np.testing.assert_allclose(lossp.cpu().numpy(), loss0.cpu().numpy(), rtol=1e-5)
    monkeypatch.setenv('B2S_FUSED_WS', '2' if K == 2 and B % 2 == 1 else '1')   # 2: pair transform in the transform warps
    loss1, perm1 = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=y_abs, num_samples=lengths)
    loss2, perm2 = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=y_abs, num_samples=lengths)
    torch.cuda.synchronize()
    assert torch.equal(loss1, loss2) and torch.equal(perm1, perm2)
    np.testing.assert_array_equal(perm1.cpu().numpy(), perm0.cpu().numpy())
    np.testing.assert_allclose(loss1.cpu().numpy(), loss0.cpu().numpy(), rtol=1e-5)
    if B <= 8:
        for b in range(B):
            n = lengths[b] if ragged else T
            m_b = stft.samples_to_frames(n)
            want_loss, want_perm, _ = OP.stft_mask_pit_step(torch.from_numpy(y[b:b + 1, :n]),
                                                           torch.from_numpy(s[b:b + 1, :, :n]),
                                                           torch.from_numpy(masks[b:b + 1, :m_b]))
            np.testing.assert_array_equal(perm1[b].cpu().numpy(), np.asarray(want_perm)[0])
            np.testing.assert_allclose(float(loss1[b]), float(want_loss[0]), rtol=LOSS_RTOL)


@pytest.mark.parametrize('eps', [1e-3, 1e-5])
def test_near_tie_permutations(b2s, eps, capsys):
    """SURVEY.md section 8(d) near-tie stress: s_2 = s_1 + eps N(0, 1).  The two candidate losses then differ by
    a tiny relative gap; bit-exact argmin is only defined where the gap exceeds the reduction error of the
    implementation it is compared with.  The kernels accumulate in fp64 -> they must agree with the float64
    oracle wherever the relative gap is above 1e-9; the float32 oracle (= the reference's arithmetic) is what may
    flip.  The histogram of gaps and both disagreement counts are printed (pytest -s / the captured log)."""
    from oracle import path as OP
    B, K, T = 48, 2, 16000
    y, s, rng = _step_inputs(B, K, T, 11, eps=eps)
    stft = b2s.ops.STFT(1024, 256)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, K, 513).astype(np.float32)
    _, perm = b2s.review.stft_mask_pit_step(torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev()),
                                            torch.from_numpy(masks).to(dev()), stft=stft)
    perm = perm.cpu().numpy()
    # float64 truth with both candidate losses (K = 2: identity and swap)
    st = OP.ReferenceSTFT(1024, 256)
    y_abs = st(torch.from_numpy(y).double()).abs()
    x_abs = st(torch.from_numpy(s).double()).abs().transpose(1, 2)
    est = torch.from_numpy(masks).double() * y_abs[:, :, None, :]
    c0 = ((est - x_abs) ** 2).mean(dim=(1, 2, 3))
    c1 = ((est[:, :, [1, 0]] - x_abs) ** 2).mean(dim=(1, 2, 3))
    truth = (c1 < c0).numpy().astype(int)           # ties -> identity (first minimum)
    gap = ((c0 - c1).abs() / torch.minimum(c0, c1)).numpy()
    _, perm32, _ = OP.stft_mask_pit_step(torch.from_numpy(y), torch.from_numpy(s), torch.from_numpy(masks))
    swap32 = np.array([p[0] for p in perm32])
    ours = perm[:, 0]
    edges = [0, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, np.inf]
    hist, _ = np.histogram(gap, bins=edges)
    with capsys.disabled():
        print(f'\n[near-tie eps={eps:g}] relative gap histogram {dict(zip([f"<{e:g}" for e in edges[1:]], hist.tolist()))}; '
              f'disagreements with the float64 oracle: kernels {int((ours != truth).sum())}/{B}, '
              f'float32 oracle (reference arithmetic) {int((swap32 != truth).sum())}/{B}')
    decided = gap > 1e-9
    np.testing.assert_array_equal(ours[decided], truth[decided])


# ------------------------------------------------------------------------------------------------ features
@pytest.mark.parametrize('power,scale_spec,log_base', [(1., False, False), (2., False, 10), (1., True, None),
                                                        (0.6, False, 2), (2., True, False)])
def test_feature_epilogues_against_reference_chain(b2s, power, scale_spec, log_base):
    """to_spectrogram (timefreq.py:171-183) + Logarithm (:37-77) fused behind the STFT, against the same chain
    of torch ops on the oracle's float64 spectrum.  Tolerance: 1e-4 of the largest value for linear features;
    log features are compared where the argument is above the clamp (their absolute error is the relative
    error of the argument)."""
    from oracle.stft import ReferenceSTFT
    rng = np.random.RandomState(17)
    x = rng.randn(3, 2, 9000).astype(np.float32)
    stft = b2s.features.SpectrogramSTFT(1024, 256, power=power, scale_spec=scale_spec, log_base=log_base,
                                        sequence_last=False)
    got, frames = stft(torch.from_numpy(x).to(dev()), sequence_lengths=[9000, 8000, 5000])
    spec = ReferenceSTFT(1024, 256)(torch.from_numpy(x).double()).abs() ** power
    if scale_spec:
        spec = spec / 1024
    assert got.shape == spec.shape
    np.testing.assert_array_equal(frames, [stft.samples_to_frames(n) for n in (9000, 8000, 5000)])
    got = got.cpu().double()
    if log_base is False:
        assert float((got - spec).abs().max()) <= 1e-4 * float(spec.abs().max())
    else:
        log_fn = {None: torch.log, 10: torch.log10, 2: torch.log2}[log_base]
        want = log_fn(torch.clamp(spec, min=1e-5))
        # d log(x) = dx / x: 1e-4 of the utterance maximum, seen from a value x, is 1e-4 max / x in log units
        bound = 1e-4 * float(spec.max()) / torch.clamp(spec, min=1e-5) + 1e-5
        scale = {None: 1.0, 10: 1 / np.log(10), 2: 1 / np.log(2)}[log_base]
        assert bool(((got - want).abs() <= bound * scale).all())
    # sequence_last=True transposes
    stft.sequence_last = True
    got_t, _ = stft(torch.from_numpy(x).to(dev()))
    assert got_t.shape == spec.transpose(-2, -1).shape


@pytest.mark.parametrize('filters,log_base', [(80, 10), (40, None), (23, False)])
def test_log_mel_against_matmul(b2s, filters, log_base):
    """MelTransform.forward (timefreq.py:398-470): (|Y|^power) @ mel_basis, then the logarithm -- the fused kernel
    against torch.matmul with the SAME basis on the oracle's float64 spectrum (this pins the kernel; the basis
    itself is paderbox's get_fbanks restated, parity unpinned)."""
    from oracle.stft import ReferenceSTFT
    rng = np.random.RandomState(5)
    x = (0.3 * rng.randn(4, 12000)).astype(np.float32)
    mel = b2s.features.MelTransform(16000, 1024, number_of_filters=filters, lowest_frequency=80,
                                    highest_frequency=7600, log_base=log_base, sequence_last=False)
    got, _ = mel(torch.from_numpy(x).to(dev()))
    spec = ReferenceSTFT(1024, 256)(torch.from_numpy(x).double()).abs()
    want = spec @ torch.from_numpy(mel.mel_basis).double()
    assert got.shape == want.shape == (4, spec.shape[1], filters)
    got = got.cpu().double()
    if log_base is False:
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())
    else:
        log_fn = {None: torch.log, 10: torch.log10}[log_base]
        ref = log_fn(torch.clamp(want, min=1e-5))
        bound = (1e-4 * float(want.max()) / torch.clamp(want, min=1e-5) + 1e-5) * (1.0 if log_base is None else 1 / np.log(10))
        assert bool(((got - ref).abs() <= bound).all())
    # an arbitrary dense basis (no triangular structure): still exact
    dense = rng.rand(513, 7).astype(np.float32)
    mel2 = b2s.features.MelTransform(16000, 1024, mel_basis=dense, log_base=False, sequence_last=False)
    got2, _ = mel2(torch.from_numpy(x).to(dev()))
    want2 = spec @ torch.from_numpy(dense).double()
    assert float((got2.cpu().double() - want2).abs().max()) <= 1e-4 * float(want2.abs().max())
    with pytest.raises(NotImplementedError):
        b2s.features.stft_features(b2s.ops.STFT(512, 128), torch.from_numpy(x).to(dev()))


# ------------------------------------------------------------------------------------------------ evaluation
@pytest.mark.parametrize('k', [2, 3, 5])
def test_evaluation_path_against_oracle(b2s, k):
    """review.evaluate_separation (pit/evaluate.py:144-176 on the device): mask * STFT(y) -> iSTFT -> SI-SDR /
    SDR of the best assignment, against the same chain through the oracle (reference iSTFT, regression losses,
    brute-force assignment by SI-SDR)."""
    from oracle import losses as ol
    from oracle.stft import ReferenceSTFT
    rng = np.random.RandomState(40 + k)
    B, T = 3, 12000
    s = (0.1 * rng.randn(B, k, T)).astype(np.float32)
    y = s.sum(1)
    stft = b2s.ops.STFT(1024, 256)
    ref = ReferenceSTFT(1024, 256)
    M = stft.samples_to_frames(T)
    # oracle-like masks: ideal ratio masks of a random permutation of the speakers + noise
    S = ref(torch.from_numpy(s).double()).abs()                       # [B, K, M, F]
    order = [rng.permutation(k) for _ in range(B)]
    irm = torch.stack([S[b, order[b]] for b in range(B)]) / (S.sum(1, keepdim=True) + 1e-9)
    masks = (irm.permute(0, 2, 1, 3) + 0.05 * torch.from_numpy(rng.rand(B, M, k, 513))).float().contiguous()
    out = b2s.review.evaluate_separation(masks.to(dev()), torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev()),
                                         stft=stft)
    Y = ref(torch.from_numpy(y).double())                              # [B, M, F] complex
    Z = masks.double().permute(0, 2, 1, 3) * Y[:, None]               # [B, K, M, F]
    z = ref.inverse(Z)[..., :T]
    assert float((out['estimates'].cpu().double() - z).abs().max()) <= 1e-4 * float(z.abs().max())
    for b in range(B):
        pair = torch.stack([torch.stack([ol.si_sdr_loss(z[b, i], torch.from_numpy(s[b, j]).double()) for j in range(k)])
                            for i in range(k)])
        best = min(itertools.permutations(range(k)), key=lambda p: (sum(float(pair[p[j], j]) for j in range(k)), p))
        assert tuple(out['permutation'][b].tolist()) == best
        for j in range(k):
            e, t = z[b, best[j]], torch.from_numpy(s[b, j]).double()
            np.testing.assert_allclose(float(out['si_sdr'][b, j]), -float(ol.si_sdr_loss(e, t)), rtol=1e-3, atol=2e-3)
            np.testing.assert_allclose(float(out['sdr'][b, j]), -float(ol.sdr_loss(e, t)), rtol=1e-3, atol=2e-3)
            obs = torch.from_numpy(y[b]).double()
            np.testing.assert_allclose(float(out['input_si_sdr'][b, j]), -float(ol.si_sdr_loss(obs, t)), rtol=1e-3, atol=2e-3)
    assert bool((out['si_sdr_improvement'] > 0).all())      # ratio masks do separate


# ------------------------------------------------------------------------------------------------ projections
@pytest.mark.parametrize('m,n,k', [(1, 8, 4), (77, 130, 36), (253, 1026, 1200), (640, 2400, 513), (1000, 1200, 1200)])
def test_tcgen05_linear_against_float64(b2s, m, n, k):
    """b2s_linear_forward (tcgen05 / TMEM / TMA) == activation(F.linear(x, w, b)) of pit/model.py:96-102.
    precision='fp32' (3-term TF32 split) must be as close to the float64 result as cuBLAS fp32 is, within a small
    factor -- the reference runs these GEMMs in fp32; precision='tf32' is held to 2e-3 of the largest output."""
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k, device=dev())
    w = torch.randn(n, k, device=dev()) / k ** 0.5
    b = torch.randn(n, device=dev())
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    scale = float(ref.abs().max())
    err_cublas = float((torch.nn.functional.linear(x, w, b).double() - ref).abs().max())
    for act, fn in ((None, lambda t: t), ('relu', torch.relu), ('sigmoid', torch.sigmoid)):
        want = fn(ref)
        got = b2s.ops.linear(x, w, b, activation=act)
        err = float((got.double() - want).abs().max())
        assert err <= max(10 * err_cublas, 2e-6 * scale), (act, err, err_cublas, scale)
        got_tf32 = b2s.ops.linear(x, w, b, activation=act, precision='tf32')
        assert float((got_tf32.double() - want).abs().max()) <= 2e-3 * max(scale, 1.0)
    # no bias, leading axes, chained lo output
    got = b2s.ops.linear(x.view(1, m, k), w)
    assert got.shape == (1, m, n)
    assert float((got[0].double() - (ref - b.double())).abs().max()) <= max(10 * err_cublas, 2e-6 * scale)


def test_tcgen05_linear_autograd_and_chaining(b2s):
    """Gradients (cuBLAS backward around the tensor-core forward) equal torch's; the `c_lo` output of one
    projection is a valid lo operand of the next (Linear + ReLU -> Linear + sigmoid, pit/model.py:98-102)."""
    from padertorch_b200.ops.linear import linear_forward, tf32_split
    torch.manual_seed(0)
    x = torch.randn(300, 1200, device=dev(), requires_grad=True)
    l1 = torch.nn.Linear(1200, 1200).to(dev())
    l2 = torch.nn.Linear(1200, 1026).to(dev())
    y = b2s.ops.linear(b2s.ops.linear(x, l1.weight, l1.bias, 'relu'), l2.weight, l2.bias, 'sigmoid')
    y.square().mean().backward()
    grads = [x.grad.clone(), l1.weight.grad.clone(), l2.bias.grad.clone()]
    x.grad = None; l1.zero_grad(); l2.zero_grad()
    y_ref = torch.sigmoid(l2(torch.relu(l1(x))))
    y_ref.square().mean().backward()
    torch.testing.assert_close(y, y_ref, rtol=1e-4, atol=1e-5)
    for g, r in zip(grads, [x.grad, l1.weight.grad, l2.bias.grad]):
        torch.testing.assert_close(g, r, rtol=1e-3, atol=1e-5 * float(r.abs().max()) + 1e-9)
    with torch.no_grad():
        h, h_lo = linear_forward(x.detach(), l1.weight, l1.bias, 'relu', want_lo=True)
        torch.testing.assert_close(h_lo, tf32_split(h), rtol=0, atol=0)
        out, _ = linear_forward(h, l2.weight, l2.bias, 'sigmoid', x_lo=h_lo)
        torch.testing.assert_close(out, y_ref, rtol=1e-4, atol=1e-5)
    m = b2s.ops.FusedLinear(1200, 600, activation='relu').to(dev())
    torch.testing.assert_close(m(x.detach()), torch.relu(torch.nn.functional.linear(x.detach(), m.weight, m.bias)),
                               rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('rows,T', [(1, 256), (1, 700), (2, 1024), (3, 5000), (5, 17001), (130, 9000), (1300, 2100),
                                    (64, 64000)])
@pytest.mark.parametrize('layout', ['complex', 'concat'])
def test_ring_inverse_against_chunked_kernel_and_oracle(b2s, rows, T, layout, monkeypatch):
    """The halo-free ring inverse (istft_ring_kernel: per-warp chunks, register overlap-add, ticketed chunk boundaries)
    against the chunked kernel it replaces (B2S_INV_RING=0) and against the oracle's iSTFT; also as the adjoint of the
    STFT (autograd of the forward transform).  Rows beyond the warp count (several units per warp), rows with fewer
    than four frames (one chunk), odd chunk lengths, both spectrum layouts; bit-identical reruns."""
    from oracle.stft import ReferenceSTFT
    torch.manual_seed(rows + T)
    stft = b2s.ops.STFT(1024, 256, complex_representation=layout)
    x = 0.1 * torch.randn(rows, T, device=dev())
    xr = x.clone().requires_grad_(True)
    spec = stft(xr)
    g = torch.randn_like(spec) if not spec.is_complex() else torch.randn_like(torch.view_as_real(spec))
    monkeypatch.setenv('B2S_INV_RING', '2')
    z1 = stft.inverse(spec.detach())
    z1b = stft.inverse(spec.detach())
    (grad1,) = torch.autograd.grad(spec if not spec.is_complex() else torch.view_as_real(spec), xr, g, retain_graph=True)
    monkeypatch.setenv('B2S_INV_RING', '1')
    z2 = stft.inverse(spec.detach())
    monkeypatch.setenv('B2S_INV_RING', '0')
    z0 = stft.inverse(spec.detach())
    (grad0,) = torch.autograd.grad(spec if not spec.is_complex() else torch.view_as_real(spec), xr, g)
    torch.cuda.synchronize()
    assert torch.equal(z1, z1b)
    scale = float(z0.abs().max())
    assert float((z1 - z0).abs().max()) <= 2e-6 * scale and float((z2 - z0).abs().max()) <= 2e-6 * scale
    assert float((grad1 - grad0).abs().max()) <= 2e-6 * float(grad0.abs().max())
    assert float((z1[..., :T] - x).abs().max()) <= 1e-4 * float(x.abs().max())     # round trip
    if rows <= 5:
        ref = ReferenceSTFT(1024, 256, complex_representation=layout)
        want = ref.inverse(spec.detach().cpu())
        assert float((z1.cpu() - want).abs().max()) <= 1e-4 * float(want.abs().max())


@pytest.mark.parametrize('B,T,shift,ragged', [(1, 3000, 256, False), (3, 9001, 256, False), (4, 12002, 256, True),
                                              (2, 8000, 512, False), (3, 7003, 128, True), (70, 2500, 256, False)])
def test_pair_kernel_edge_cases_against_oracle(b2s, B, T, shift, ragged, monkeypatch):
    """The two-source pair-transform kernel (csrc/fused_pair.cuh; the default when |Y| is given) on the cases its fast
    path does not cover: rows that are not 16-byte aligned (odd lengths: 4-byte zero-filling copies), other shifts,
    ragged batches whose ranges start or end in the padding of a shorter example, more examples than pipelines have
    positions; with and without the opt-in hop ring.  Permutations bit exact, losses within 1e-4 of the oracle."""
    from oracle import path as OP
    from oracle.stft import ReferenceSTFT
    y, s, rng = _step_inputs(B, 2, T, 23 + B)
    lengths = [int(T - 700 * (b % 3) - b) for b in range(B)] if ragged else None
    if ragged:
        for b, n in enumerate(lengths):
            y[b, n:] = 0
            s[b, :, n:] = 0
    stft = b2s.ops.STFT(1024, shift)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, 2, 513).astype(np.float32)
    yd, sd, md = torch.from_numpy(y).to(dev()), torch.from_numpy(s).to(dev()), torch.from_numpy(masks).to(dev())
    y_abs = stft.magnitude(yd)
    results = []
    for ring in ('0', '1'):
        monkeypatch.setenv('B2S_PAIR_RING', ring)
        loss, perm = b2s.review.stft_mask_pit_step(None, sd, md, stft=stft, observation_abs=y_abs, num_samples=lengths)
        results.append((loss.cpu().numpy(), perm.cpu().numpy()))
    monkeypatch.delenv('B2S_PAIR_RING', raising=False)
    np.testing.assert_array_equal(results[0][1], results[1][1])
    np.testing.assert_allclose(results[0][0], results[1][0], rtol=1e-6)
    ref = ReferenceSTFT(1024, shift)
    for b in range(min(B, 6)):
        n = lengths[b] if ragged else T
        m_b = stft.samples_to_frames(n)
        want_loss, want_perm, _ = OP.stft_mask_pit_step(torch.from_numpy(y[b:b + 1, :n]), torch.from_numpy(s[b:b + 1, :, :n]),
                                                       torch.from_numpy(masks[b:b + 1, :m_b]), stft=ref)
        np.testing.assert_array_equal(results[0][1][b], np.asarray(want_perm)[0])
        np.testing.assert_allclose(results[0][0][b], float(want_loss[0]), rtol=LOSS_RTOL)
