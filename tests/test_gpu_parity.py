"""GPU parity: the sm_100a kernels (called through the C ABI via padertorch_b200) against

* the golden fixtures recorded from the UNMODIFIED reference (tests/golden, oracle/make_golden.py),
* the CPU oracle (oracle/) on fresh seeded inputs,
* size-independent properties at BASELINE.json's full sizes.

Tolerances follow BASELINE.json's north star: integer results (frame counts, permutations) bit exact;
fp32 spectra within 1e-4 of the per-tensor maximum; losses within 1e-4 relative.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SPEC_RTOL = 1e-4      # relative to max |reference| of the tensor
LOSS_RTOL = 1e-4


def dev():
    return torch.device('cuda:0')


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def assert_spec_close(got, want, rtol=SPEC_RTOL, what=''):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(float(np.abs(want).max()), 1e-30) if want.size else 1.0
    err = float(np.abs(got - want).max()) if want.size else 0.0
    assert err <= rtol * scale, f'{what}: max err {err:.3e} > {rtol} * {scale:.3e}'


def assert_loss_close(got, want, rtol=LOSS_RTOL, what=''):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * 1e-2, err_msg=what)


@pytest.fixture(scope='module')
def b2s():
    import padertorch_b200
    return padertorch_b200


def stft_cases(golden):
    return sorted(golden.index['stft'])


# ------------------------------------------------------------------------------------------------ STFT
def _stft_case_names():
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'index.json')) as fd:
        return sorted(json.load(fd)['stft'])


@pytest.mark.parametrize('name', _stft_case_names())
def test_stft_forward_inverse_golden(b2s, golden, name):
    entry = golden.index['stft'][name]
    kwargs = dict(entry['kwargs'])
    x = golden(f'stft/{name}/x').astype(np.float32)
    reps = [r for r in ('complex', 'concat', 'stacked') if golden.has(f'stft/{name}/{r}')]
    assert reps
    for rep in reps:
        stft = b2s.ops.STFT(complex_representation=rep, **kwargs)
        out = stft(cuda(x))
        want = golden(f'stft/{name}/{rep}')
        assert tuple(out.shape) == want.shape
        got = out.cpu().numpy()
        if rep == 'complex':
            assert out.dtype == torch.complex64
            got = np.stack([got.real, got.imag], -1)
            want_cmp = np.stack([want.real, want.imag], -1)
        else:
            want_cmp = want
        assert_spec_close(got, want_cmp, what=f'{name}/{rep}')
        # inverse of the reference's own spectrum
        spec = torch.from_numpy(want.astype(np.complex64 if rep == 'complex' else np.float32)).to(dev())
        back = stft.inverse(spec)
        assert_spec_close(back.cpu().numpy(), golden(f'stft/{name}/{rep}_inverse'),
                          what=f'{name}/{rep}_inverse')


@pytest.mark.parametrize('name', _stft_case_names())
def test_stft_frame_arithmetic_golden(b2s, golden, name):
    entry = golden.index['stft'][name]
    stft = b2s.ops.STFT(**entry['kwargs'])
    for samples, frames in entry['frames'].items():
        assert stft.samples_to_frames(int(samples)) == frames      # bit exact
    for frames, samples in entry['frames_to_samples'].items():
        assert stft.frames_to_samples(int(frames)) == samples
    arr = np.array([int(s) for s in entry['frames']])
    np.testing.assert_array_equal(stft.samples_to_frames(arr), [entry['frames'][str(s)] for s in arr])


@pytest.mark.parametrize('name', _stft_case_names())
def test_stft_autograd_golden(b2s, golden, name):
    entry = golden.index['stft'][name]
    stft = b2s.ops.STFT(complex_representation='stacked', **entry['kwargs'])
    x = cuda(golden(f'stft/{name}/x').astype(np.float32)).requires_grad_(True)
    g = cuda(golden(f'stft/{name}/grad_out_stacked').astype(np.float32))
    out = stft(x)
    (grad_x,) = torch.autograd.grad((out * g).sum(), x)
    assert_spec_close(grad_x.cpu().numpy(), golden(f'stft/{name}/grad_x'), what=f'{name}/grad_x')
    spec = cuda(golden(f'stft/{name}/inv_in_stacked').astype(np.float32)).requires_grad_(True)
    sig = stft.inverse(spec)
    assert_spec_close(sig.detach().cpu().numpy(), golden(f'stft/{name}/inv_out'), what=f'{name}/inv_out')
    gs = cuda(golden(f'stft/{name}/inv_grad_out').astype(np.float32))
    (grad_spec,) = torch.autograd.grad((sig * gs).sum(), spec)
    assert_spec_close(grad_spec.cpu().numpy(), golden(f'stft/{name}/inv_grad_in'),
                      what=f'{name}/inv_grad_in')


def test_stft_literal_vector(b2s):
    # padertorch/contrib/cb/transform.py:219-232
    expect = np.array([[0.5 + 0j, 0 + 0.5j, -0.5 + 0j],
                       [4 + 0j, -2 + 1j, 0 + 0j],
                       [8 + 0j, -4 + 1j, 0 + 0j],
                       [12 + 0j, -6 + 1j, 0 + 0j],
                       [3.5 + 0j, 0 - 3.5j, -3.5 + 0j]])
    out = b2s.ops.STFT(4, 2, window='hann')(cuda(np.arange(8, dtype=np.float32))).cpu().numpy()
    np.testing.assert_allclose(out, expect, atol=1e-5)


@pytest.mark.parametrize('size,shift,wl,samples,plain,faded', [
    (1024, 256, 1024, (1023, 1024, 1025), (1, 1, 2), (7, 7, 8)),       # tests/test_ops/test_stft.py:44-70
    (512, 20, 40, (1019, 1020, 1021), (50, 50, 51), (52, 52, 53)),     # :139-165
])
def test_stft_frame_counts_reference_tests(b2s, size, shift, wl, samples, plain, faded):
    stft = b2s.ops.STFT(size, shift, window_length=wl, complex_representation='concat')
    for fading, expect in ((False, plain), (True, faded)):
        stft.fading = fading       # mutated after construction, like the reference test
        for n, m in zip(samples, expect):
            assert stft(torch.rand(n, device=dev())).shape == (m, 2 * (size // 2 + 1))
            assert stft.samples_to_frames(n) == m


@pytest.mark.parametrize('kwargs', [
    dict(size=1024, shift=256), dict(size=1024, shift=512, window='hann'),
    dict(size=1024, shift=128, window='hamming'), dict(size=1024, shift=256, window_length=800),
    dict(size=1024, shift=300), dict(size=1024, shift=255), dict(size=1024, shift=256, fading='half'),
    dict(size=1024, shift=256, fading=None, pad=False), dict(size=512, shift=128),
    dict(size=2048, shift=512), dict(size=400, shift=160, window='hann'),
])
def test_stft_against_oracle_fp64(b2s, kwargs):
    """Fresh inputs, float64 rfft formulation of the oracle as the truth; ragged lengths."""
    from oracle import stft as OS
    rng = np.random.RandomState(7)
    for samples in (5000, 4097, 1000 if kwargs.get('pad', True) else 2100):
        x = rng.randn(3, samples).astype(np.float32)
        want = OS.stft_rfft(x.astype(np.float64), **kwargs)
        stft = b2s.ops.STFT(**kwargs)
        got = stft(cuda(x)).cpu().numpy()
        assert_spec_close(np.stack([got.real, got.imag], -1), np.stack([want.real, want.imag], -1),
                          what=f'{kwargs} T={samples}')
        mag = stft.magnitude(cuda(x)).cpu().numpy()
        assert_spec_close(mag, np.abs(want), what=f'abs {kwargs}')
        lmag = stft.magnitude(cuda(x), log1p=True).cpu().numpy()
        assert_spec_close(lmag, np.log1p(np.abs(want)), what=f'log1p {kwargs}')
        back = stft.inverse(torch.from_numpy(want.astype(np.complex64)).to(dev())).cpu().numpy()
        want_back = OS.istft_rfft(want, **{k: v for k, v in kwargs.items() if k != 'pad'})
        assert_spec_close(back, want_back, what=f'inverse {kwargs} T={samples}')


def test_stft_unaligned_rows_and_views(b2s):
    """Odd row length / offset views take the scalar-load path; results must not change."""
    from oracle import stft as OS
    rng = np.random.RandomState(3)
    base = rng.randn(4, 7001).astype(np.float32)
    x = cuda(base)[:, 1:]                      # rows start at odd element offsets
    want = OS.stft_rfft(base[:, 1:].astype(np.float64), 1024, 256)
    got = b2s.ops.STFT(1024, 256)(x).cpu().numpy()
    assert_spec_close(np.stack([got.real, got.imag], -1), np.stack([want.real, want.imag], -1))


def test_stft_errors(b2s):
    with pytest.raises(AssertionError):
        b2s.ops.STFT(1023, 256)
    with pytest.raises(AssertionError):
        b2s.ops.STFT(1024, 256, complex_representation='polar')
    with pytest.raises(AssertionError):
        b2s.ops.STFT(1024, 256, fading='quarter')
    with pytest.raises(RuntimeError):
        b2s.ops.STFT(1024, 256)(torch.zeros(2, 4000))            # CPU tensor: no fallback
    with pytest.raises(TypeError):
        b2s.ops.STFT(1024, 256)(torch.zeros(2, 4000, dtype=torch.float64, device=dev()))
    with pytest.raises(RuntimeError):
        b2s.ops.STFT(1024, 256, fading=None, pad=False)(torch.zeros(2, 100, device=dev()))


# ------------------------------------------------------------------------------------------------ regression
REGRESSION_FNS = ['mse_loss', 'log_mse_loss', 'log1p_mse_loss', 'sdr_loss', 'si_sdr_loss',
                  'source_aggregated_sdr_loss']


@pytest.mark.parametrize('shape_name', ['k2_t4000', 'k3_t1000', 'b4_k2_t501', 'vec_t100'])
@pytest.mark.parametrize('fname', REGRESSION_FNS)
def test_regression_golden(b2s, golden, shape_name, fname):
    fn = getattr(b2s.ops.losses, fname)
    for v, kwargs in golden.index['regression'][fname].items():
        e = cuda(golden(f'regression/{shape_name}/estimate')).requires_grad_(True)
        t = cuda(golden(f'regression/{shape_name}/target'))
        out = fn(e, t, **kwargs)
        want = golden(f'regression/{shape_name}/{fname}/{v}')
        assert tuple(out.shape) == want.shape, (fname, kwargs, out.shape, want.shape)
        assert_loss_close(out.detach().cpu().numpy(), want, what=f'{fname} {kwargs}')
        (grad,) = torch.autograd.grad(out.sum(), e)
        assert_spec_close(grad.cpu().numpy(), golden(f'regression/{shape_name}/{fname}/{v}/grad'),
                          rtol=2e-4, what=f'{fname} {kwargs} grad')


def test_regression_known_answers(b2s):
    L = b2s.ops.losses
    e = torch.tensor([[1., 2, 3], [4, 5, 6]], device=dev())
    t = torch.tensor([[2., 3, 4], [4, 0, 6]], device=dev())
    close = lambda a, b: np.testing.assert_allclose(a.cpu().numpy(), b, atol=5e-5)  # noqa: E731
    close(L.mse_loss(e, t), 9.3333)                                  # regression.py:61-66
    close(L.mse_loss(e, t, reduction=None), [1.0, 8.3333])
    close(L.log_mse_loss(e, t), 0.9208)                              # :113-120
    close(L.log_mse_loss(t, t, soft_sdr_max=20), -1.7758)
    close(L.sdr_loss(e, t), -6.5167)                                 # :144-155
    close(L.sdr_loss(e, t, reduction=None), [-9.8528, -3.1806])
    close(L.sdr_loss(t, t, soft_sdr_max=20), -20.)
    close(L.si_sdr_loss(e, t), -10.7099)                             # :202-271
    close(L.si_sdr_loss(e, t, reduction=None), [-18.2391, -3.1806])
    close(L.log1p_mse_loss(e, t), 1.2711)                            # :331-336
    close(L.source_aggregated_sdr_loss(e, t), -4.6133)               # :354-366
    e2 = torch.tensor([[1., 2, 3], [4, 2, 6]], device=dev())
    t2 = torch.tensor([[2., 3, 4], [6, 4, 8]], device=dev())
    close(L.source_aggregated_sdr_loss(e2, t2), -9.8528)
    zero = torch.zeros(2, 3, device=dev())
    assert torch.isnan(L.si_sdr_loss(zero, t)).all()                 # NaN cases :248-269
    assert torch.isnan(L.si_sdr_loss(e, zero)).all()
    assert float(L.si_sdr_loss(t, t)) < -130                         # perfect estimate bound


# ------------------------------------------------------------------------------------------------ PIT / DC ops
def _pit_fn(b2s, name):
    L = b2s.ops.losses
    return {'mse': torch.nn.functional.mse_loss, 'pt_mse': L.mse_loss, 'log_mse': L.log_mse_loss,
            'log1p_mse': L.log1p_mse_loss, 'sdr': L.sdr_loss, 'si_sdr': L.si_sdr_loss}[name]


def _pit_cases():
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'index.json')) as fd:
        index = json.load(fd)['pit']
    return [(case, fn) for case in sorted(index) for fn in index[case]['loss_fns']]


@pytest.mark.parametrize('case,lname', _pit_cases())
def test_pit_golden(b2s, golden, case, lname):
    axis = golden.index['pit'][case]['axis']
    e = cuda(golden(f'pit/{case}/estimate')).requires_grad_(True)
    t = cuda(golden(f'pit/{case}/target'))
    loss, perm = b2s.ops.pit_loss(e, t, axis=axis, loss_fn=_pit_fn(b2s, lname), return_permutation=True)
    assert loss.dim() == 0
    assert tuple(perm) == tuple(int(v) for v in golden(f'pit/{case}/{lname}/perm'))     # bit exact
    assert_loss_close(loss.detach().cpu().numpy(), golden(f'pit/{case}/{lname}/loss'), what=f'{case}/{lname}')
    (grad,) = torch.autograd.grad(loss, e)
    assert_spec_close(grad.cpu().numpy(), golden(f'pit/{case}/{lname}/grad'), rtol=2e-4,
                      what=f'{case}/{lname} grad')
    if golden.has(f'pit/{case}/{lname}/pairwise'):
        matrix = b2s.ops.compute_pairwise_losses(e.detach(), t, axis=axis, loss_fn=_pit_fn(b2s, lname))
        assert_spec_close(matrix.cpu().numpy(), golden(f'pit/{case}/{lname}/pairwise'), rtol=2e-4,
                          what=f'{case}/{lname} pairwise')
        for red in ('mean', 'sum'):
            val, cols = b2s.ops.pit_loss_from_loss_matrix(matrix, reduction=red, return_permutation=True)
            assert_loss_close(val.cpu().numpy(), golden(f'pit/{case}/{lname}/matrix_{red}'), rtol=2e-4)
        np.testing.assert_array_equal(cols, golden(f'pit/{case}/{lname}/matrix_cols'))


def test_pit_known_answers(b2s):
    pit = b2s.ops.pit_loss
    d = dev()
    # source_separation.py:64-93
    assert float(pit(torch.ones(5, 3, 10, device=d), torch.zeros(5, 3, 10, device=d), axis=-2)) == 1.0
    t = torch.randn(100, 2, 257, device=d)
    loss, perm = pit(torch.stack([t[:, 1], t[:, 0]], 1), t, axis=-2, return_permutation=True)
    assert float(loss) == 0.0 and perm == (1, 0)
    assert float(pit(torch.ones(2, 3, 4, 5, 6, device=d), torch.zeros(2, 3, 4, 5, 6, device=d), axis=-3)) == 1.0
    # exact tie: first permutation in itertools order wins (torch.min on CPU)
    loss, perm = pit(torch.zeros(7, 3, 9, device=d), torch.zeros(7, 3, 9, device=d), axis=1,
                     return_permutation=True)
    assert perm == (0, 1, 2) and float(loss) == 0.0
    # tests/test_ops/test_losses.py:137-150
    e = torch.tensor([[0.], [2.]], device=d).view(1, 2, 1)
    assert float(pit(e, torch.tensor([[2.], [0.]], device=d).view(1, 2, 1), axis=1)) == 0.0
    assert float(pit(e, torch.tensor([[-1.], [0.]], device=d).view(1, 2, 1), axis=1)) == 2.5
    # generic callable path (cross entropy variant, source_separation.py:86-93)
    logits = torch.zeros(4, 2, 6, device=d)
    target = torch.zeros(4, 6, dtype=torch.long, device=d)
    ce = pit(logits, target, axis=1, loss_fn=torch.nn.functional.cross_entropy)
    np.testing.assert_allclose(float(ce), 0.6931, atol=1e-4)
    with pytest.raises(AssertionError):
        pit(torch.zeros(3, 2, 4, device=d), torch.zeros(3, 2, 5, device=d), axis=1)
    with pytest.raises(RuntimeError):
        pit(torch.zeros(3, 2, 4), torch.zeros(3, 2, 4), axis=1)          # CPU: no fallback


@pytest.mark.parametrize('k', [1, 2, 3, 4, 5])
def test_pit_against_oracle_random(b2s, k):
    from oracle import losses as OL
    rng = np.random.RandomState(10 + k)
    for trial in range(3):
        t = np.abs(rng.randn(41, k, 129)).astype(np.float32)
        order = rng.permutation(k)
        e = (t[:, order] + 0.3 * rng.randn(*t.shape)).astype(np.float32)
        want, want_perm = OL.pit_loss(torch.from_numpy(e).double(), torch.from_numpy(t).double(), axis=-2,
                                      return_permutation=True)
        got, perm = b2s.ops.pit_loss(cuda(e), cuda(t), axis=-2, return_permutation=True)
        assert tuple(perm) == tuple(want_perm)
        assert_loss_close(got.cpu().numpy(), want.numpy())
        for fn_name in ('si_sdr_loss', 'log_mse_loss', 'sdr_loss'):
            tt = rng.randn(k, 3001).astype(np.float32)
            ee = (tt[rng.permutation(k)] + 0.5 * rng.randn(k, 3001)).astype(np.float32)
            want, want_perm = OL.pit_loss(torch.from_numpy(ee).double(), torch.from_numpy(tt).double(),
                                          axis=0, loss_fn=getattr(OL, fn_name), return_permutation=True)
            got, perm = b2s.ops.pit_loss(cuda(ee), cuda(tt), axis=0,
                                         loss_fn=getattr(b2s.ops.losses, fn_name), return_permutation=True)
            assert tuple(perm) == tuple(want_perm), fn_name
            assert_loss_close(got.cpu().numpy(), want.numpy(), what=fn_name)


@pytest.mark.parametrize('case', ['n100_e20_k3', 'n4104_e20_k2', 'n999_e7_k4'])
def test_dc_golden(b2s, golden, case):
    x = cuda(golden(f'dc/{case}/x')).requires_grad_(True)
    t = cuda(golden(f'dc/{case}/t'))
    loss = b2s.ops.deep_clustering_loss(x, t)
    assert loss.dim() == 0
    assert_loss_close(loss.detach().cpu().numpy(), golden(f'dc/{case}/loss'), what=case)
    (grad,) = torch.autograd.grad(loss, x)
    assert_spec_close(grad.cpu().numpy(), golden(f'dc/{case}/grad'), rtol=2e-4, what=f'{case} grad')


def test_dc_analytic(b2s):
    # tests/test_ops/test_losses.py:105-116: perfect embedding -> 0; single swapped point -> 4 / N^2 scale
    n, k = 64, 2
    labels = np.arange(n) % k
    t = np.eye(k, dtype=np.float32)[labels]
    loss = b2s.ops.deep_clustering_loss(cuda(t.copy()), cuda(t))
    np.testing.assert_allclose(float(loss), 0.0, atol=1e-6)


# ------------------------------------------------------------------------------------------------ model-level
def test_pit_review_golden(b2s, golden):
    lengths = golden.index['models']['pit']['lengths']
    masks = [cuda(golden(f'models/pit/mask_{b}')).requires_grad_(True) for b in range(len(lengths))]
    y = [cuda(golden(f'models/pit/y_abs_{b}')) for b in range(len(lengths))]
    x = [cuda(golden(f'models/pit/x_abs_{b}')) for b in range(len(lengths))]
    cpd = [cuda(golden(f'models/pit/cpd_{b}')) for b in range(len(lengths))]
    out = b2s.review.pit_review_losses(masks, y, x, cpd)
    assert_loss_close(out['pit_mse_loss'].detach().cpu().numpy(), golden('models/pit/pit_mse_loss'))
    assert_loss_close(out['pit_ips_loss'].detach().cpu().numpy(), golden('models/pit/pit_ips_loss'))
    grads = torch.autograd.grad(out['pit_mse_loss'] + out['pit_ips_loss'], masks)
    for b, g in enumerate(grads):
        assert_spec_close(g.cpu().numpy(), golden(f'models/pit/grad_mask_{b}'), rtol=2e-4, what=f'grad {b}')
    # padded-batch entry point gives the same numbers
    T = max(lengths)
    pad = lambda seq: torch.stack([torch.nn.functional.pad(a, (0, 0) * (a.dim() - 1) + (0, T - a.shape[0]))  # noqa: E731
                                   for a in seq])
    out2 = b2s.review.pit_review_losses(pad([m.detach() for m in masks]), pad(y), pad(x), pad(cpd), lengths)
    assert_loss_close(out2['pit_mse_loss'].cpu().numpy(), golden('models/pit/pit_mse_loss'))
    assert_loss_close(out2['pit_ips_loss'].cpu().numpy(), golden('models/pit/pit_ips_loss'))
    # minibatch loss == mean of single-example losses (tests/test_models/test_bss.py:57-83)
    singles = [b2s.review.pit_review_losses([m.detach()], [yy], [xx], [cc])['pit_mse_loss']
               for m, yy, xx, cc in zip(masks, y, x, cpd)]
    np.testing.assert_allclose(float(torch.stack(singles).mean()), float(out['pit_mse_loss']), atol=1e-6)


def test_dc_review_golden(b2s, golden):
    lengths = golden.index['models']['dc']['lengths']
    emb = [cuda(golden(f'models/dc/embedding_{b}')).requires_grad_(True) for b in range(len(lengths))]
    tm = [cuda(golden(f'models/dc/target_mask_{b}')) for b in range(len(lengths))]
    loss = b2s.review.dc_review_loss(emb, tm)
    assert_loss_close(loss.detach().cpu().numpy(), golden('models/dc/dc_loss'))
    grads = torch.autograd.grad(loss, emb)
    for b, g in enumerate(grads):
        assert_spec_close(g.cpu().numpy(), golden(f'models/dc/grad_embedding_{b}'), rtol=2e-4, what=f'grad {b}')
    T = max(lengths)
    pad = lambda seq: torch.stack([torch.nn.functional.pad(a.detach(), (0, 0, 0, 0, 0, T - a.shape[0])) for a in seq])  # noqa: E731
    loss2 = b2s.review.dc_review_loss(pad(emb), pad(tm), lengths)
    assert_loss_close(loss2.cpu().numpy(), golden('models/dc/dc_loss'))


def test_dc_review_reference_geometry(b2s):
    """E = 20, K = 2, F = 513 (the geometry the frame-tiled kernels are specialised for at compile time),
    ragged lengths, list and padded entry points, against the oracle in float64."""
    from oracle import path as oracle_path
    rng = np.random.RandomState(7)
    lengths, E, K, F = [9, 4, 1, 6], 20, 2, 513
    emb_np = [rng.randn(T, E, F).astype(np.float32) for T in lengths]
    emb_np = [e / np.linalg.norm(e, axis=1, keepdims=True) for e in emb_np]
    tm_np = [np.eye(K, dtype=np.float32)[rng.randint(0, K, (T, F))].transpose(0, 2, 1).copy() for T in lengths]
    ref_emb = [torch.from_numpy(e).double().requires_grad_(True) for e in emb_np]
    want, _ = oracle_path.dc_review_loss(ref_emb, [torch.from_numpy(t).double() for t in tm_np])
    want_grads = torch.autograd.grad(want, ref_emb)
    emb = [cuda(e).requires_grad_(True) for e in emb_np]
    tm = [cuda(t) for t in tm_np]
    loss = b2s.review.dc_review_loss(emb, tm)
    assert_loss_close(loss.detach().cpu().numpy(), want.detach().numpy())
    for b, (g, w) in enumerate(zip(torch.autograd.grad(loss, emb), want_grads)):
        assert_spec_close(g.cpu().numpy(), w.numpy(), rtol=2e-4, what=f'grad {b}')
    T = max(lengths)
    pad = lambda seq: torch.stack([torch.nn.functional.pad(a.detach(), (0, 0, 0, 0, 0, T - a.shape[0])) for a in seq])  # noqa: E731
    padded = pad(emb).requires_grad_(True)
    loss2 = b2s.review.dc_review_loss(padded, pad(tm), lengths)
    assert_loss_close(loss2.detach().cpu().numpy(), want.detach().numpy())
    (g2,) = torch.autograd.grad(loss2, padded)
    for b, w in enumerate(want_grads):
        assert_spec_close(g2[b, :lengths[b]].cpu().numpy(), w.numpy(), rtol=2e-4, what=f'padded grad {b}')
        assert float(g2[b, lengths[b]:].abs().max()) == 0.0 if lengths[b] < T else True
    # an unaligned view: frames start at odd float offsets
    base = torch.zeros(1 + emb_np[0].size, device=dev())
    view = base[1:].view(emb_np[0].shape)
    view.copy_(cuda(emb_np[0]))
    single = b2s.review.dc_review_loss([view], [tm[0]])
    want0, _ = oracle_path.dc_review_loss([torch.from_numpy(emb_np[0]).double()], [torch.from_numpy(tm_np[0]).double()])
    assert_loss_close(single.cpu().numpy(), want0.numpy())


@pytest.mark.parametrize('lengths', [[40], [3], [1, 2, 300, 7, 64], [17] * 37, [5 + (11 * b) % 23 for b in range(148)],
                                     [4 + b % 9 for b in range(151)], [120, 8, 9, 500, 33, 16, 250]])
def test_dc_balanced_chunks(b2s, lengths):
    """Length-balanced chunk slots of the ring Gram kernel and the frame backward kernel (balanced_chunk_slot): one
    example, more examples than SMs (the uniform grid takes over), equal lengths (closed form), chunks shorter than
    eight frames, extreme spreads -- per-example losses and gradients against the oracle in float64, and bit-identical
    to the uniform grid (B2S_DC_BALANCE is read once per process, so the uniform grid is checked through the oracle)."""
    from oracle import path as oracle_path
    rng = np.random.RandomState(len(lengths) * 7 + lengths[0])
    E, K, F = 20, 2, 513
    T = max(lengths)
    emb_np = rng.randn(len(lengths), T, E, F).astype(np.float32)
    emb_np /= np.linalg.norm(emb_np, axis=2, keepdims=True)
    tm_np = np.eye(K, dtype=np.float32)[rng.randint(0, K, (len(lengths), T, F))].transpose(0, 1, 3, 2).copy()
    emb = cuda(emb_np).requires_grad_(True)
    losses = b2s.review.dc_losses_per_example(emb, cuda(tm_np), lengths)
    weights = torch.linspace(0.5, 1.5, len(lengths), device=dev())
    (grad,) = torch.autograd.grad((losses * weights).sum(), emb)
    again = b2s.review.dc_losses_per_example(emb.detach(), cuda(tm_np), lengths)
    assert torch.equal(again, losses.detach())   # fixed summation order
    check = sorted(set([0, len(lengths) - 1, int(np.argmax(lengths)), int(np.argmin(lengths))]))
    for b in check:
        e = torch.from_numpy(emb_np[b, :lengths[b]]).double().requires_grad_(True)
        want, _ = oracle_path.dc_review_loss([e], [torch.from_numpy(tm_np[b, :lengths[b]]).double()])
        assert_loss_close(losses[b].detach().cpu().numpy(), want.detach().numpy(), what=f'loss {b}')
        (wg,) = torch.autograd.grad(want * float(weights[b]), e)
        assert_spec_close(grad[b, :lengths[b]].cpu().numpy(), wg.numpy(), rtol=2e-4, what=f'grad {b}')
        if lengths[b] < T:
            assert float(grad[b, lengths[b]:].abs().max()) == 0.0


@pytest.mark.parametrize('K,F,lengths,dual', [(2, 513, [30, 11, 25], True), (2, 513, [17] * 5, False), (3, 257, [9, 5], True),
                                              (2, 40000, [3], False)])
def test_pit_review_losses_folded_means(b2s, K, F, lengths, dual):
    """pit_review_losses (means from b2s_pit_sse_forward_mean, backward through b2s_pit_sse_backward_scaled with one upstream
    value per loss) == means of pit_losses_per_example and their autograd: padded and list entry points, ragged lengths,
    frame-staged and direct kernels, repeated calls."""
    g = torch.Generator().manual_seed(K * 100 + len(lengths))
    T = max(lengths)
    B = len(lengths)
    masks = torch.rand(B, T, K, F, generator=g).to(dev()).requires_grad_(True)
    yab = torch.rand(B, T, F, generator=g).to(dev())
    xab = torch.rand(B, T, K, F, generator=g).to(dev())
    cpd = (torch.rand(B, T, K, F, generator=g) * 2 - 1).to(dev()) if dual else None
    mse, _, ips, _ = b2s.review.pit_losses_per_example(masks, yab, xab, cpd, lengths)
    want = 1.5 * mse.mean() + (0.25 * ips.mean() if dual else 0.0)
    (want_grad,) = torch.autograd.grad(want, masks)
    atol = 1e-6 * float(want_grad.abs().max())
    for _ in range(2):
        out = b2s.review.pit_review_losses(masks, yab, xab, cpd, lengths)
        torch.testing.assert_close(out['pit_mse_loss'], mse.mean(), rtol=1e-6, atol=0)
        got = 1.5 * out['pit_mse_loss']
        if dual:
            torch.testing.assert_close(out['pit_ips_loss'], ips.mean(), rtol=1e-6, atol=0)
            got = got + 0.25 * out['pit_ips_loss']
        (grad,) = torch.autograd.grad(got, masks)
        torch.testing.assert_close(grad, want_grad, rtol=1e-5, atol=atol)
    mask_list = [masks[b, :n].detach().clone().requires_grad_(True) for b, n in enumerate(lengths)]
    out = b2s.review.pit_review_losses(mask_list, [yab[b, :n].contiguous() for b, n in enumerate(lengths)],
                                       [xab[b, :n].contiguous() for b, n in enumerate(lengths)],
                                       [cpd[b, :n].contiguous() for b, n in enumerate(lengths)] if dual else None)
    got = 1.5 * out['pit_mse_loss'] + (0.25 * out['pit_ips_loss'] if dual else 0.0)
    torch.testing.assert_close(got, want.detach(), rtol=1e-6, atol=0)
    grads = torch.autograd.grad(got, mask_list)
    for b, n in enumerate(lengths):
        torch.testing.assert_close(grads[b], want_grad[b, :n], rtol=1e-5, atol=atol)


@pytest.mark.parametrize('F,E,K,lengths', [(513, 20, 2, [40, 12, 33, 7]), (513, 20, 2, [25] * 6), (65, 7, 3, [5, 9, 4]),
                                            (513, 30, 2, [6, 11]), (513, 20, 2, [19])])
def test_dc_review_loss_folded_mean(b2s, F, E, K, lengths):
    """dc_review_loss (batch mean folded by the Gram launch, b2s_dc_forward_mean; backward with the broadcast upstream
    gradient, b2s_dc_backward_scaled) == mean of dc_losses_per_example and its autograd, for the ring kernel, the
    six-warp kernel and the generic kernels (mean by the fallback kernel); repeated calls (the ticket returns to zero)."""
    rng = np.random.RandomState(F + E + len(lengths))
    T = max(lengths)
    emb_np = rng.randn(len(lengths), T, E, F).astype(np.float32)
    emb_np /= np.linalg.norm(emb_np, axis=2, keepdims=True)
    tm = cuda(np.eye(K, dtype=np.float32)[rng.randint(0, K, (len(lengths), T, F))].transpose(0, 1, 3, 2).copy())
    emb = cuda(emb_np).requires_grad_(True)
    want = b2s.review.dc_losses_per_example(emb, tm, lengths).mean()
    (want_grad,) = torch.autograd.grad(2.5 * want, emb)
    # the upstream factor 2.5 / B is rounded to float32 on one route and applied in float64 on the other
    atol = 1e-6 * float(want_grad.abs().max())
    for _ in range(3):
        got = b2s.review.dc_review_loss(emb, tm, lengths)
        assert got.shape == ()
        torch.testing.assert_close(got, want, rtol=1e-6, atol=0)
        (grad,) = torch.autograd.grad(2.5 * got, emb)
        torch.testing.assert_close(grad, want_grad, rtol=1e-5, atol=atol)
    # list entry point (the model's own call)
    emb_list = [cuda(emb_np[b, :n]).requires_grad_(True) for b, n in enumerate(lengths)]
    got = b2s.review.dc_review_loss(emb_list, [tm[b, :n].contiguous() for b, n in enumerate(lengths)])
    torch.testing.assert_close(got, want, rtol=1e-6, atol=0)
    grads = torch.autograd.grad(2.5 * got, emb_list)
    for b, n in enumerate(lengths):
        torch.testing.assert_close(grads[b], want_grad[b, :n], rtol=1e-5, atol=atol)


@pytest.mark.parametrize('F,E,K', [(257, 20, 2), (513, 20, 3), (257, 20, 3), (129, 20, 2), (257, 16, 2)])
def test_dc_compile_time_geometries(b2s, F, E, K):
    """The compile-time instances of the ring Gram / frame backward kernels beyond 513 / 20 / 2 (the reference model's
    F = 257, three speakers) and two run-time geometries next to them, ragged batch, against the oracle in float64."""
    from oracle import path as oracle_path
    rng = np.random.RandomState(F + E + K)
    lengths = [37, 9, 64, 21, 50]
    T = max(lengths)
    emb_np = rng.randn(len(lengths), T, E, F).astype(np.float32)
    emb_np /= np.linalg.norm(emb_np, axis=2, keepdims=True)
    tm_np = np.eye(K, dtype=np.float32)[rng.randint(0, K, (len(lengths), T, F))].transpose(0, 1, 3, 2).copy()
    emb = cuda(emb_np).requires_grad_(True)
    losses = b2s.review.dc_losses_per_example(emb, cuda(tm_np), lengths)
    (grad,) = torch.autograd.grad(losses.sum(), emb)
    for b in range(len(lengths)):
        e = torch.from_numpy(emb_np[b, :lengths[b]]).double().requires_grad_(True)
        want, _ = oracle_path.dc_review_loss([e], [torch.from_numpy(tm_np[b, :lengths[b]]).double()])
        assert_loss_close(losses[b].detach().cpu().numpy(), want.detach().numpy(), what=f'loss {b}')
        (wg,) = torch.autograd.grad(want, e)
        assert_spec_close(grad[b, :lengths[b]].cpu().numpy(), wg.numpy(), rtol=2e-4, what=f'grad {b}')
        if lengths[b] < T:
            assert float(grad[b, lengths[b]:].abs().max()) == 0.0


@pytest.mark.parametrize('K', [3, 5])
def test_tasnet_losses_more_sources(b2s, K):
    """The one-launch loss set (K <= 4: thread per example) and its K > 4 fallback against the oracle."""
    from oracle import path as oracle_path
    rng = np.random.RandomState(K)
    B, T = 5, 1500
    num_samples = [1500, 1203, 1500, 777, 1001]
    s = rng.randn(B, K, T).astype(np.float32)
    est = (s[:, ::-1] + 0.4 * rng.randn(B, K, T)).astype(np.float32).copy()
    ref_est = torch.from_numpy(est).double().requires_grad_(True)
    want = oracle_path.tasnet_losses(ref_est, torch.from_numpy(s).double(), num_samples)
    e = cuda(est).requires_grad_(True)
    out = b2s.review.tasnet_losses(e, cuda(s), num_samples)
    for name in ('si-sdr', 'log-mse', 'log1p-mse'):
        assert_loss_close(out[name].detach().cpu().numpy(), want[name].detach().numpy(), what=name)
    (grad,) = torch.autograd.grad(out['si-sdr'] + 0.5 * out['log1p-mse'], e)
    (want_grad,) = torch.autograd.grad(want['si-sdr'] + 0.5 * want['log1p-mse'], ref_est)
    assert_spec_close(grad.cpu().numpy(), want_grad.numpy(), rtol=2e-4, what='combined gradient')
    for b, n in enumerate(num_samples):
        assert float(grad[b, :, n:].abs().max()) == 0.0 if n < T else True


@pytest.mark.parametrize('K,B,T,aligned', [(2, 64, 64000, True), (2, 5, 3001, True), (3, 7, 4096, True),
                                          (4, 3, 2500, True), (2, 6, 3001, False), (5, 4, 1999, True),
                                          (1, 9, 1200, True)])
def test_pair_stats_loss_set_one_launch(b2s, K, B, T, aligned):
    """b2s_pair_stats_loss_set (statistics + loss set + batch means in ONE launch) == b2s_pair_stats_forward followed by
    b2s_pair_loss_set: identical statistics, losses and permutations, repeated calls (the tickets return to zero),
    ragged lengths, rows that take the internal two-launch route (unaligned, K > 4)."""
    from padertorch_b200 import _lib
    from padertorch_b200._workspace import meta_tensor
    from padertorch_b200.ops.losses import _pairs
    g = torch.Generator(device='cpu').manual_seed(K * 1000 + B)
    length = T if aligned else T + 1
    s = torch.randn(B, K, length, generator=g).to(dev())
    est = (s.flip(1) + 0.5 * torch.randn(B, K, length, generator=g).to(dev()))
    if not aligned:   # rows start at odd float offsets
        s, est = s[..., 1:], est[..., 1:]
        assert s.data_ptr() % 16 != 0
    num = [T - (37 * b) % (T // 3) for b in range(B)]
    num[0] = T
    stride_e, stride_t = est.stride(1), s.stride(1)
    rows = [[num[b], b * est.stride(0), b * s.stride(0)] for b in range(B)]
    meta = meta_tensor(rows, est.device)
    problem = _pairs.PairProblem(est, s, meta, B, 1, K, max(num), stride_e, stride_t)
    kinds = [_lib.LOSS_SI_SDR, _lib.LOSS_LOG_MSE, _lib.LOSS_LOG1P_MSE, _lib.LOSS_MSE, _lib.LOSS_SDR]
    reductions = [_lib.REDUCE_MEAN, _lib.REDUCE_SUM, _lib.REDUCE_SUM, _lib.REDUCE_MEAN, _lib.REDUCE_SUM]
    stats2 = problem.stats()
    loss2, perm2, mean2 = problem.loss_set(stats2, kinds, reductions)
    for _ in range(3):
        stats1, loss1, perm1, mean1 = problem.stats_loss_set(kinds, reductions, one_launch=True)
        assert torch.equal(stats1, stats2)
        assert torch.equal(loss1, loss2) and torch.equal(perm1, perm2)
        torch.testing.assert_close(mean1, mean2, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(mean1, loss1.double().mean(1).float(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('size,shift,T,K', [(1024, 256, 5000, 2), (1024, 256, 5001, 2), (1024, 256, 9000, 3),
                                           (1024, 256, 4100, 1), (1024, 512, 6000, 2), (512, 128, 3001, 3)])
def test_prepare_pit_targets(b2s, size, shift, T, K):
    """pre_batch_transform on the device (pit/data.py:49-77) against the oracle: |Y|, |X| in 't k f' layout and
    cos(angle(Y) - angle(X)); the phase term is compared where both magnitudes are well above rounding noise."""
    from oracle import path as oracle_path
    from oracle.stft import ReferenceSTFT
    rng = np.random.RandomState(size + K)
    B = 3
    s = (0.1 * rng.randn(B, K, T)).astype(np.float32)
    if K > 1:
        s[1, 0, :] = 0.0                   # a silent source: angle(0) = 0
    y = s.sum(1)
    stft = b2s.ops.STFT(size, shift)
    out = b2s.review.prepare_pit_targets(cuda(y), cuda(s), stft=stft)
    ref = ReferenceSTFT(size, shift)
    for b in range(B):
        y_abs, x_abs, cpd = oracle_path.prepare_pit_example(torch.from_numpy(y[b]).double(), torch.from_numpy(s[b]).double(), ref)
        assert out['num_frames'] == y_abs.shape[0]
        assert_spec_close(out['Y_abs'][b].cpu().numpy(), y_abs.numpy(), what='Y_abs')
        assert_spec_close(out['X_abs'][b].cpu().numpy(), x_abs.numpy(), what='X_abs')
        got = out['cos_phase_difference'][b].cpu().numpy()
        assert got.shape == tuple(cpd.shape) and np.abs(got).max() <= 1.0
        strong = ((y_abs[:, None, :] > 1e-2 * y_abs.max()) & (x_abs > 1e-2 * x_abs.max().clamp_min(1e-30))).numpy()
        assert strong.mean() > 0.3
        np.testing.assert_allclose(got[strong], cpd.numpy()[strong], atol=2e-4)
    if K > 1:   # silent source: angle(X) = 0 -> cos(angle(Y))
        Y1 = ref(torch.from_numpy(y[1]).double())
        want = torch.cos(torch.angle(Y1)).numpy()
        strong = (Y1.abs() > 1e-2 * Y1.abs().max()).numpy()
        np.testing.assert_allclose(out['cos_phase_difference'][1, :, 0].cpu().numpy()[strong], want[strong], atol=2e-4)
        assert float(out['X_abs'][1, :, 0].abs().max()) == 0.0
    # the prepared tensors drive the un-fused review exactly like the reference's
    masks = cuda(rng.rand(B, out['num_frames'], K, size // 2 + 1).astype(np.float32))
    got = b2s.review.pit_review_losses(masks, out['Y_abs'], out['X_abs'], out['cos_phase_difference'])
    want = oracle_path.pit_review_losses(
        [m.double() for m in masks.cpu()],
        [a.double() for a in out['Y_abs'].cpu()], [a.double() for a in out['X_abs'].cpu()],
        [a.double() for a in out['cos_phase_difference'].cpu()])
    assert_loss_close(got['pit_mse_loss'].cpu().numpy(), want['pit_mse_loss'].numpy())
    assert_loss_close(got['pit_ips_loss'].cpu().numpy(), want['pit_ips_loss'].numpy())


@pytest.mark.parametrize('size,shift,T,K', [(1024, 256, 9000, 2), (1024, 256, 7001, 4), (1024, 256, 6000, 3),
                                           (512, 128, 3001, 2)])
def test_prepare_pit_targets_ragged(b2s, size, shift, T, K):
    """Ragged batches (num_samples) and K = 4 through the one-kernel target preparation (b2s_stft_pit_targets with a
    meta table; the two-transform path for other plans): every example equals the oracle on its own length, rows of
    frames beyond an example's frame count are zeros, and garbage beyond an example's samples is never read."""
    from oracle import path as oracle_path
    from oracle.stft import ReferenceSTFT
    rng = np.random.RandomState(size + 7 * K)
    B = 4
    s = (0.1 * rng.randn(B, K, T)).astype(np.float32)
    lengths = [T, T - 1234, T // 2 + 3, 2048]
    y = s.sum(1)
    s_dirty, y_dirty = s.copy(), y.copy()
    for b, n in enumerate(lengths):       # NaN beyond the lengths: must not reach the outputs
        s_dirty[b, :, n:] = np.nan
        y_dirty[b, n:] = np.nan
    stft = b2s.ops.STFT(size, shift)
    out = b2s.review.prepare_pit_targets(cuda(y_dirty), cuda(s_dirty), stft=stft, num_samples=lengths)
    ref = ReferenceSTFT(size, shift)
    assert len(out['num_frames']) == B
    for b, n in enumerate(lengths):
        y_abs, x_abs, cpd = oracle_path.prepare_pit_example(torch.from_numpy(y[b, :n]).double(),
                                                            torch.from_numpy(s[b, :, :n]).double(), ref)
        m_b = y_abs.shape[0]
        assert out['num_frames'][b] == m_b
        assert_spec_close(out['Y_abs'][b, :m_b].cpu().numpy(), y_abs.numpy(), what='Y_abs')
        assert_spec_close(out['X_abs'][b, :m_b].cpu().numpy(), x_abs.numpy(), what='X_abs')
        got = out['cos_phase_difference'][b, :m_b].cpu().numpy()
        strong = ((y_abs[:, None, :] > 1e-2 * y_abs.max()) & (x_abs > 1e-2 * x_abs.max().clamp_min(1e-30))).numpy()
        np.testing.assert_allclose(got[strong], cpd.numpy()[strong], atol=2e-4)
        for name in ('Y_abs', 'X_abs', 'cos_phase_difference'):
            tail = out[name][b, m_b:]
            assert tail.numel() == 0 or float(tail.abs().max()) == 0.0, name


def test_tasnet_losses_golden(b2s, golden):
    num_samples = golden.index['models']['tasnet']['num_samples']
    s = cuda(golden('models/tasnet/s'))
    est = cuda(golden('models/tasnet/estimate')).requires_grad_(True)
    out = b2s.review.tasnet_losses(est, s, num_samples)
    for name in ('si-sdr', 'log-mse', 'log1p-mse'):
        assert_loss_close(out[name].detach().cpu().numpy(), golden(f'models/tasnet/{name}'), what=name)
        (grad,) = torch.autograd.grad(out[name], est, retain_graph=True)
        assert_spec_close(grad.cpu().numpy(), golden(f'models/tasnet/{name}/grad'), rtol=2e-4, what=name)


# ------------------------------------------------------------------------------------------------ fused step
def test_fused_step_golden(b2s, golden):
    y, s, masks = cuda(golden('step/y')), cuda(golden('step/s')), cuda(golden('step/masks'))
    stft = b2s.ops.STFT(1024, 256)
    y_abs = stft.magnitude(y)
    assert_spec_close(y_abs.cpu().numpy(), golden('step/Y_abs'))
    for kwargs in (dict(mixture=y), dict(mixture=None, observation_abs=y_abs)):
        loss, perm = b2s.review.stft_mask_pit_step(sources=s, masks=masks, stft=stft, **kwargs)
        np.testing.assert_array_equal(perm.cpu().numpy(), golden('step/perm'))          # bit exact
        assert_loss_close(loss.cpu().numpy(), golden('step/loss'))
    # un-fused composition of the same kernels agrees
    x_abs = stft.magnitude(s).transpose(1, 2).contiguous()
    mse, perm2, _, _ = b2s.review.pit_losses_per_example(masks, y_abs, x_abs)
    np.testing.assert_array_equal(perm2.cpu().numpy(), golden('step/perm'))
    assert_loss_close(mse.cpu().numpy(), golden('step/loss'))


@pytest.mark.parametrize('k,seconds', [(2, 1.0), (3, 0.7)])
def test_fused_step_against_oracle(b2s, k, seconds):
    from oracle import path as OP
    rng = np.random.RandomState(21)
    B, T = 3, int(16000 * seconds)
    s = (0.1 * rng.randn(B, k, T)).astype(np.float32)
    y = s.sum(1)
    stft = b2s.ops.STFT(1024, 256)
    M = stft.samples_to_frames(T)
    masks = rng.rand(B, M, k, 513).astype(np.float32)
    want_loss, want_perm, want_yabs = OP.stft_mask_pit_step(torch.from_numpy(y), torch.from_numpy(s),
                                                            torch.from_numpy(masks))
    loss, perm = b2s.review.stft_mask_pit_step(cuda(y), cuda(s), cuda(masks), stft=stft)
    np.testing.assert_array_equal(perm.cpu().numpy(), np.asarray(want_perm))
    assert_loss_close(loss.cpu().numpy(), want_loss.numpy())
    # ragged: per-example lengths
    num_samples = [T, T - 1234, T - 4000]
    loss_r, perm_r = b2s.review.stft_mask_pit_step(cuda(y), cuda(s), cuda(masks), stft=stft,
                                                   num_samples=num_samples)
    for b, n in enumerate(num_samples):
        m_b = stft.samples_to_frames(n)
        w_loss, w_perm, _ = OP.stft_mask_pit_step(torch.from_numpy(y[b:b + 1, :n]),
                                                  torch.from_numpy(s[b:b + 1, :, :n]),
                                                  torch.from_numpy(masks[b:b + 1, :m_b]))
        assert tuple(perm_r[b].tolist()) == tuple(w_perm[0])
        assert_loss_close(loss_r[b].cpu().numpy(), w_loss[0].numpy())


def test_fused_step_unaligned_views_and_tiny_batches(b2s):
    """The fused kernel's TMA pipelines fall back to zero-filling cp.async for rows that are not 16-byte
    aligned; batches with fewer frame positions than warps leave pipelines idle.  Results must not change."""
    from oracle import path as OP
    rng = np.random.RandomState(5)
    stft = b2s.ops.STFT(1024, 256)
    for B, T in ((1, 1300), (2, 5000)):
        k = 2
        base = (0.1 * rng.randn(B, k, T + 3)).astype(np.float32)
        s_np = np.ascontiguousarray(base[:, :, 1:T + 1])
        y_np = s_np.sum(1)
        M = stft.samples_to_frames(T)
        masks = rng.rand(B, M, k, 513).astype(np.float32)
        want_loss, want_perm, _ = OP.stft_mask_pit_step(torch.from_numpy(y_np), torch.from_numpy(s_np),
                                                        torch.from_numpy(masks))
        # aligned tensor and a contiguous tensor whose storage starts 4 bytes off a 16-byte boundary
        flat = torch.zeros(s_np.size + 1, dtype=torch.float32, device=dev())
        flat[1:] = cuda(s_np).reshape(-1)
        s_off = flat[1:].view(B, k, T)
        assert s_off.is_contiguous() and s_off.data_ptr() % 16 != 0
        for s_dev in (cuda(s_np), s_off):
            loss, perm = b2s.review.stft_mask_pit_step(cuda(y_np), s_dev, cuda(masks), stft=stft)
            np.testing.assert_array_equal(perm.cpu().numpy(), np.asarray(want_perm))
            assert_loss_close(loss.cpu().numpy(), want_loss.numpy())
        # odd sample count: frames end inside a 16-byte unit
        T2 = T - 1
        s2, y2 = np.ascontiguousarray(s_np[:, :, :T2]), np.ascontiguousarray(y_np[:, :T2])
        M2 = stft.samples_to_frames(T2)
        w_loss, w_perm, _ = OP.stft_mask_pit_step(torch.from_numpy(y2), torch.from_numpy(s2),
                                                  torch.from_numpy(masks[:, :M2]))
        loss, perm = b2s.review.stft_mask_pit_step(cuda(y2), cuda(s2), cuda(np.ascontiguousarray(masks[:, :M2])),
                                                   stft=stft)
        np.testing.assert_array_equal(perm.cpu().numpy(), np.asarray(w_perm))
        assert_loss_close(loss.cpu().numpy(), w_loss.numpy())


# ------------------------------------------------------------------------------------------------ full size
def test_full_size_properties(b2s):
    """BASELINE.json's headline shape: batch 64 x 4 s x 16 kHz, 2 speakers, STFT(1024, 256)."""
    torch.manual_seed(0)
    d = dev()
    B, K, T = 64, 2, 64000
    s = 0.1 * torch.randn(B, K, T, device=d)
    y = s.sum(1)
    stft = b2s.ops.STFT(1024, 256)
    Y = stft(y)
    assert Y.shape == (B, 253, 513) and Y.dtype == torch.complex64
    # round trip (tests/test_ops/test_stft.py:36-42)
    back = stft.inverse(Y)
    assert back.shape == (B, T)
    assert float((back - y).abs().max()) <= 1e-4 * float(y.abs().max())
    # linearity: STFT(y) == sum_k STFT(s_k)
    S = stft(s)
    assert float((S.sum(1) - Y).abs().max()) <= 1e-4 * float(Y.abs().max())
    # Parseval-style checksum against the time domain (window energy is folded in by the round trip)
    y_abs = stft.magnitude(y)
    assert float((y_abs - Y.abs()).abs().max()) <= 1e-4 * float(y_abs.max())
    # ideal-ratio masks with a known swap: permutation must follow the swap, bit exact
    x_abs = S.abs().transpose(1, 2).contiguous()                  # [B, M, K, F]
    ideal = x_abs / (y_abs[:, :, None, :] + 1e-3)
    swap = torch.arange(B, device=d) % 3 == 0
    masks = torch.where(swap[:, None, None, None], ideal.flip(2), ideal).contiguous()
    loss, perm = b2s.review.stft_mask_pit_step(y, s, masks, stft=stft)
    want = torch.where(swap[:, None], torch.tensor([1, 0], device=d), torch.tensor([0, 1], device=d))
    assert torch.equal(perm.long(), want)
    loss2, perm2 = b2s.review.stft_mask_pit_step(None, s, masks, stft=stft, observation_abs=y_abs)
    assert torch.equal(perm2, perm)
    mse, perm3, _, _ = b2s.review.pit_losses_per_example(masks, y_abs, x_abs)
    assert torch.equal(perm3, perm)
    # (ideal masks leave residuals of 1e-3 |X|: the loss is a near-cancellation, and rounding differences of 1e-7 |X|
    # between two correct fp32 transforms -- the |Y|-recomputing kernel uses the 8 x 8 x 8 transform, the front-end and
    # the two-source kernel the pair transform -- show up amplified in it; the absolute tolerance is therefore stated
    # relative to the targets' energy, 1e-7 of it)
    atol = 1e-7 * float((x_abs ** 2).mean())
    torch.testing.assert_close(loss, mse, rtol=1e-4, atol=atol)
    torch.testing.assert_close(loss2, mse, rtol=1e-4, atol=atol)
    # determinism (Trainer.test_run compares two runs at 1e-5 / 1e-6): bit identical here
    loss_again, _ = b2s.review.stft_mask_pit_step(y, s, masks, stft=stft)
    assert torch.equal(loss, loss_again)
    assert torch.equal(torch.view_as_real(stft(y)), torch.view_as_real(Y))


def test_three_speaker_eight_seconds(b2s):
    """BASELINE config 5 shape: 3 speakers, 8 s (503 frames), 3! permutations."""
    from oracle import losses as OL
    torch.manual_seed(1)
    d = dev()
    B, K, T = 4, 3, 128000
    s = 0.1 * torch.randn(B, K, T, device=d)
    y = s.sum(1)
    stft = b2s.ops.STFT(1024, 256)
    y_abs, x_abs = stft.magnitude(y), stft.magnitude(s).transpose(1, 2).contiguous()
    assert y_abs.shape == (B, 503, 513)
    masks = torch.rand(B, 503, K, 513, device=d)
    order = [(2, 0, 1), (0, 1, 2), (1, 2, 0), (2, 1, 0)]
    for b, o in enumerate(order):
        masks[b] = (x_abs[b] / (y_abs[b][:, None] + 1e-3))[:, list(o)]
    loss, perm = b2s.review.stft_mask_pit_step(y, s, masks, stft=stft)
    for b in range(B):
        w, wp = OL.pit_loss((masks[b] * y_abs[b][:, None]).cpu().double(), x_abs[b].cpu().double(), axis=-2,
                            return_permutation=True)
        assert tuple(perm[b].tolist()) == tuple(wp)
        assert_loss_close(loss[b].cpu().numpy(), w.numpy())


def test_deep_clustering_full_size_properties(b2s):
    """BASELINE config 3 shape: batch 16, variable length 2-8 s (128..503 frames), E = 20, K = 2, F = 513.
    Size-independent properties of deep_clustering_loss (source_separation.py:13-31):
    L(aV) = a^4 |V^T V|^2 - 2 a^2 |V^T Y|^2 + |Y^T Y|^2 (over N^2), <grad L, V> = dL(aV)/da at a = 1."""
    torch.manual_seed(3)
    d = dev()
    B, E, K, F = 16, 20, 2, 513
    lengths = [int(v) for v in torch.randint(128, 504, (B,))]
    lengths[0], lengths[1] = 503, 128
    T = max(lengths)
    emb = torch.nn.functional.normalize(torch.randn(B, T, E, F, device=d), dim=2)
    labels = torch.randint(0, K, (B, T, F), device=d)
    tm = torch.nn.functional.one_hot(labels, K).permute(0, 1, 3, 2).float().contiguous()
    per_example = lambda e: b2s.review.dc_losses_per_example(e, tm, lengths)  # noqa: E731
    # perfect embedding: the target itself in the first K channels -> V V^T == Y Y^T -> 0
    perfect = torch.zeros_like(emb)
    perfect[:, :, :K] = tm
    assert float(per_example(perfect).abs().max()) <= 1e-6
    # invariance under a permutation of the embedding channels
    base = per_example(emb)
    shuffled = per_example(emb[:, :, torch.randperm(E, device=d)].contiguous())
    torch.testing.assert_close(shuffled, base, rtol=1e-5, atol=0)
    # list entry point == padded entry point; reruns are bit identical
    as_list = b2s.review.dc_losses_per_example([emb[b, :n] for b, n in enumerate(lengths)],
                                               [tm[b, :n] for b, n in enumerate(lengths)])
    torch.testing.assert_close(as_list, base, rtol=1e-5, atol=0)
    assert torch.equal(per_example(emb), base)
    # quartic in the scale of V: fit A, B, C from a = 1, 0.5, 2 and predict a = 1.5; <grad, V> = 4A - 4B
    l1, lh, l2, l15 = (per_example(a * emb).double() for a in (1.0, 0.5, 2.0, 1.5))
    # l1 = A - 2B + C, lh = A/16 - B/2 + C, l2 = 16A - 8B + C
    A = (l2 - l1 - 4.0 * (l1 - lh)) / 11.25
    Bc = (15.0 * A - (l2 - l1)) / 6.0
    C = l1 - A + 2.0 * Bc
    torch.testing.assert_close(1.5 ** 4 * A - 2 * 1.5 ** 2 * Bc + C, l15, rtol=2e-4, atol=0)
    v = emb.clone().requires_grad_(True)
    b2s.review.dc_losses_per_example(v, tm, lengths).sum().backward()
    along = (v.grad.double() * emb.double()).sum(dim=(1, 2, 3))
    torch.testing.assert_close(along, 4.0 * A - 4.0 * Bc, rtol=2e-3, atol=0)
    for b, n in enumerate(lengths):
        if n < T:
            assert float(v.grad[b, n:].abs().max()) == 0.0


def test_tasnet_losses_full_size_properties(b2s):
    """BASELINE config 4 shape per GPU: batch 32 x 2 speakers x 4 s, PIT over si-sdr / log-mse / log1p-mse
    (tasnet/model.py:154-176): planted permutations, scale invariance, known SNR, batch mean == mean of
    single-example losses (tests/test_models/test_bss.py:57-83), an independent float64 evaluation on the GPU."""
    torch.manual_seed(4)
    d = dev()
    B, K, T = 32, 2, 64000
    s = torch.randn(B, K, T, device=d)
    swap = torch.arange(B, device=d) % 2 == 1
    noise = 0.1 * torch.randn(B, K, T, device=d)
    est = torch.where(swap[:, None, None], s.flip(1), s) + noise
    num = [T] * B
    out = b2s.review.tasnet_losses(est, s, num)
    # independent float64 evaluation with the planted assignment
    e64 = torch.where(swap[:, None, None], est.flip(1), est).double()
    s64 = s.double()
    alpha = (e64 * s64).sum(-1, keepdim=True) / (s64 * s64).sum(-1, keepdim=True)
    si_sdr = -10 * torch.log10((alpha * s64).pow(2).sum(-1) / (e64 - alpha * s64).pow(2).sum(-1))
    mse = (e64 - s64).pow(2).mean(-1)
    assert_loss_close(out['si-sdr'].cpu().numpy(), si_sdr.mean(-1).mean().cpu().numpy(), what='si-sdr')
    assert_loss_close(out['log-mse'].cpu().numpy(), torch.log10(mse).sum(-1).mean().cpu().numpy(), what='log-mse')
    assert_loss_close(out['log1p-mse'].cpu().numpy(), torch.log10(1 + mse).sum(-1).mean().cpu().numpy(), what='log1p')
    # known SNR: 20 dB of independent noise -> si-sdr within 0.1 dB of -20
    assert abs(float(out['si-sdr']) + 20.0) < 0.1
    # planted permutation, bit exact; scale invariance of si-sdr
    for b in (0, 1, 30, 31):
        value, perm = b2s.ops.pit_loss(est[b], s[b], axis=0, loss_fn=b2s.ops.si_sdr_loss, return_permutation=True)
        assert tuple(perm) == ((1, 0) if b % 2 else (0, 1))
        scaled = b2s.ops.pit_loss(3.7 * est[b], s[b], axis=0, loss_fn=b2s.ops.si_sdr_loss)
        torch.testing.assert_close(scaled, value, rtol=1e-5, atol=1e-5)
    # minibatch loss == mean of single-example losses
    singles = torch.stack([b2s.review.tasnet_losses(est[b:b + 1], s[b:b + 1], [T])['si-sdr'] for b in range(0, B, 8)])
    eight = b2s.review.tasnet_losses(est[::8].contiguous(), s[::8].contiguous(), [T] * 4)['si-sdr']
    np.testing.assert_allclose(float(singles.mean()), float(eight), atol=1e-5)
    # gradient: descends the loss, zero beyond the length of a shortened example
    e = est.clone().requires_grad_(True)
    short = [T] * (B - 1) + [T // 2]
    b2s.review.tasnet_losses(e, s, short)['si-sdr'].backward()
    assert torch.isfinite(e.grad).all() and float(e.grad[-1, :, T // 2:].abs().max()) == 0.0
    stepped = b2s.review.tasnet_losses((e - 100.0 * e.grad).detach(), s, short)['si-sdr']
    assert float(stepped) < float(b2s.review.tasnet_losses(est, s, short)['si-sdr'])
