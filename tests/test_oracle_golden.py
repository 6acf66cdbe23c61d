"""Pin the CPU oracle (oracle/) to the reference: known-answer vectors the reference's own
tests/doctests hold for the path and fixtures recorded from the
unmodified reference by oracle/make_golden.py."""
import math

import numpy as np
import pytest
import torch

from oracle import losses as L
from oracle import path as P
from oracle import stft as OS


# ----------------------------------------------------------------- known answers (reference tests / doctests)
def test_literal_stft_vector():
    # padertorch/contrib/cb/transform.py:219-232
    expect = np.array([[0.5 + 0j, 0 + 0.5j, -0.5 + 0j],
                       [4 + 0j, -2 + 1j, 0 + 0j],
                       [8 + 0j, -4 + 1j, 0 + 0j],
                       [12 + 0j, -6 + 1j, 0 + 0j],
                       [3.5 + 0j, 0 - 3.5j, -3.5 + 0j]])
    x = np.arange(8).astype(np.float32)
    np.testing.assert_allclose(OS.stft_rfft(x, 4, 2, window='hann'), expect, atol=1e-12)
    got = OS.ReferenceSTFT(4, 2, window='hann')(torch.from_numpy(x).double()).numpy()
    np.testing.assert_allclose(got, expect, atol=1e-12)


@pytest.mark.parametrize('size,shift,wl,samples,plain,faded', [
    # tests/test_ops/test_stft.py:44-70
    (1024, 256, 1024, (1023, 1024, 1025), (1, 1, 2), (7, 7, 8)),
    # tests/test_ops/test_stft.py:139-165
    (512, 20, 40, (1019, 1020, 1021), (50, 50, 51), (52, 52, 53)),
])
def test_frame_counts(size, shift, wl, samples, plain, faded):
    stft = OS.ReferenceSTFT(size, shift, window_length=wl, complex_representation='concat')
    for fading, expect in ((False, plain), (True, faded)):
        stft.fading = fading       # mutated after construction, like the reference test
        for n, m in zip(samples, expect):
            assert stft(torch.rand(n)).shape == (m, 2 * (size // 2 + 1))
            assert stft.samples_to_frames(n) == m
            assert OS.stft_rfft(np.zeros(n), size, shift, window_length=wl,
                                fading=fading).shape[0] == m


def test_doctest_shapes():
    # ops/_stft.py:109-128, 190-223; tas_coders.py:140-155, 197-208
    x = torch.rand(2, 6, 203)
    assert OS.ReferenceSTFT(512, 20, window_length=40,
                            complex_representation='concat')(x).shape == (2, 6, 12, 514)
    assert OS.ReferenceSTFT(512, 20, window_length=40)(x).shape == (2, 6, 12, 257)
    inv = OS.ReferenceSTFT(512, 20, window_length=40, complex_representation='concat')
    assert inv.inverse(torch.rand(2, 4, 10, 514)).shape == (2, 4, 180)
    enc = OS.ReferenceSTFT(256, 10, window_length=20, fading=False,
                           complex_representation='concat')
    assert enc(x).shape == (2, 6, 20, 258)
    assert [enc.samples_to_frames(n) for n in (203, 150)] == [20, 14]
    assert enc.inverse(torch.rand(2, 4, 10, 258)).shape == (2, 4, 110)


@pytest.mark.parametrize('kwargs', [dict(size=1024, shift=256), dict(size=512, shift=20, window_length=40, window='hamming'),
                                    dict(size=512, shift=128, fading='half'), dict(size=256, shift=64, fading=None)])
def test_round_trip(kwargs):
    # tests/test_ops/test_stft.py:36-42
    x = np.random.RandomState(0).randn(2, 5000)
    # without full fading the first / last window is only partially overlapped
    edge = 0 if kwargs.get('fading', 'full') == 'full' else kwargs['size']
    keep = slice(edge, x.shape[-1] - edge)
    back = OS.istft_rfft(OS.stft_rfft(x, **kwargs), **kwargs)[..., :x.shape[-1]]
    np.testing.assert_allclose(back[..., keep], x[..., keep], atol=1e-10)
    stft = OS.ReferenceSTFT(**kwargs)
    back = stft.inverse(stft(torch.from_numpy(x)))[..., :x.shape[-1]].numpy()
    np.testing.assert_allclose(back[..., keep], x[..., keep], atol=1e-10)


def test_regression_known_answers():
    e = torch.tensor([[1., 2, 3], [4, 5, 6]])
    t = torch.tensor([[2., 3, 4], [4, 0, 6]])
    close = lambda a, b: np.testing.assert_allclose(np.asarray(a), b, atol=5e-5)  # noqa: E731
    close(L.mse_loss(e, t), 9.3333)                                  # regression.py:61-66
    close(L.mse_loss(e, t, reduction=None), [1.0, 8.3333])
    close(L.log_mse_loss(e, t), 0.9208)                              # :113-120
    close(L.log_mse_loss(e, t, reduction=None), [0.0, 0.9208])
    close(L.log_mse_loss(t, t, soft_sdr_max=20), -1.7758)
    close(L.sdr_loss(e, t), -6.5167)                                 # :144-155
    close(L.sdr_loss(e, t, reduction=None), [-9.8528, -3.1806])
    close(L.sdr_loss(t, t, soft_sdr_max=20), -20.)
    close(L.sdr_loss(torch.tensor([1, 2 + 3j, 4j]), torch.tensor([2, 3 + 3j, 5j])), -11.9498)
    close(L.si_sdr_loss(e, t), -10.7099)                             # :202-205
    close(L.si_sdr_loss(e, t, reduction=None), [-18.2391, -3.1806])
    close(L.si_sdr_loss(t, t, soft_sdr_max=20), -20.)
    close(L.log1p_mse_loss(e, t), 1.2711)                            # :331-336
    close(L.log1p_mse_loss(e, t, reduction=None), [0.3010, 0.9700])
    close(L.source_aggregated_sdr_loss(e, t), -4.6133)               # :354-366
    e2, t2 = torch.tensor([[1., 2, 3], [4, 2, 6]]), torch.tensor([[2., 3, 4], [6, 4, 8]])
    close(L.source_aggregated_sdr_loss(e2, t2), -9.8528)
    ref = torch.tensor(np.random.RandomState(0).randn(100))          # :207-243
    close(L.si_sdr_loss(ref, torch.flip(ref, (-1,))), 25.1277)
    close(L.si_sdr_loss(ref, ref + torch.flip(ref, (-1,))), -0.4811)
    close(L.si_sdr_loss(ref, ref + 0.5), -6.3705)
    close(L.si_sdr_loss(ref, ref * 2 + 1), -6.3705)
    assert L.si_sdr_loss(ref, ref) < -300
    assert L.si_sdr_loss(ref.float(), ref.float()) < -130
    for a, b in (([1., 0], [0., 0]), ([0., 0], [0., 0]), ([0., 0], [1., 0])):  # :248-269
        assert torch.isnan(L.si_sdr_loss(torch.tensor(a).double(), torch.tensor(b).double()))


def test_pit_known_answers():
    # source_separation.py:64-93
    T, K, F = 4, 2, 5
    assert L.pit_loss(torch.ones(T, K, F), torch.zeros(T, K, F), 1) == 1
    ce = L.pit_loss(torch.ones(T, K, F), torch.zeros(T, F, dtype=torch.int64), 1,
                    loss_fn=torch.nn.functional.cross_entropy)
    np.testing.assert_allclose(ce, 0.6931, atol=5e-5)
    assert L.pit_loss(torch.ones(K, F, T), torch.zeros(K, F, T), 0) == 1
    est = torch.stack([torch.ones(F, T), torch.zeros(F, T)])
    loss, perm = L.pit_loss(est, est[(1, 0), :, :], axis=0, return_permutation=True)
    assert loss == 0 and perm == (1, 0)
    assert L.pit_loss(torch.ones(5), torch.zeros(5), axis=0) == 1
    assert L.pit_loss(torch.ones(4, 5, 3, 100, 128), torch.zeros(4, 5, 3, 100, 128), axis=-3) == 1
    # tests/test_ops/test_losses.py:137-150
    for e, t, expect in (([[[0], [2]]], [[[0], [2]]], 0), ([[[0], [2]]], [[[2], [0]]], 0),
                         ([[[0], [2]]], [[[-1], [0]]], 2.5), ([[[0], [1]]], [[[0], [1]]], 0)):
        got = L.pit_loss(torch.tensor(e, dtype=torch.float32), torch.tensor(t, dtype=torch.float32), axis=-2)
        np.testing.assert_allclose(got, expect, rtol=1e-4)
    # source_separation.py:262-274
    score = torch.tensor(-np.array([[11., 10, 0], [4, 5, 10], [6, 0, 5]]))
    assert L.pit_loss_from_loss_matrix(score, reduction='sum') == -26
    # pairwise route == permutation route (source_separation.py:168-199)
    m = L.compute_pairwise_losses(torch.ones(T, K, F), torch.zeros(T, K, F), 1)
    assert L.pit_loss_from_loss_matrix(m) == 1


def test_dc_known_answers():
    # tests/test_ops/test_losses.py:87-127
    def numpy_reference(embedding, target_mask):
        n = embedding.shape[0]
        e, t = embedding / np.sqrt(n), target_mask / np.sqrt(n)
        return (np.sum(np.einsum('ne,nE->eE', e, e) ** 2)
                - 2 * np.sum(np.einsum('ne,nE->eE', e, t) ** 2)
                + np.sum(np.einsum('ne,nE->eE', t, t) ** 2))
    same = np.array([[1., 0], [1, 0], [0, 1]])
    assert abs(float(L.deep_clustering_loss(torch.tensor(same), torch.tensor(same)))) < 1e-6
    emb = np.array([[1., 0], [1, 0], [1, 0]])
    tgt = np.array([[1., 0], [0, 1], [1, 0]])
    np.testing.assert_allclose(L.deep_clustering_loss(torch.tensor(emb), torch.tensor(tgt)),
                               4 / 9, atol=1e-6)
    rng = np.random.RandomState(0)
    embedding = rng.normal(size=(100, 20))
    mask = rng.choice([0, 1], size=(100, 3)).astype(np.float64)
    got = L.deep_clustering_loss(torch.tensor(embedding, dtype=torch.float32),
                                 torch.tensor(mask, dtype=torch.float32))
    np.testing.assert_allclose(got, numpy_reference(embedding, mask), atol=1e-4)


# ----------------------------------------------------------------- fixtures recorded from the reference
def _tol(dtype):
    return 1e-11 if dtype == 'float64' else 2e-5


def test_golden_stft(golden):
    for name, entry in golden.index['stft'].items():
        kwargs, dtype = entry['kwargs'], entry['dtype']
        x = golden(f'stft/{name}/x')
        ref = golden(f'stft/{name}/complex')
        scale = np.abs(ref).max()
        port = OS.ReferenceSTFT(**kwargs)
        got = port(torch.from_numpy(x)).numpy()
        assert got.shape == ref.shape and got.dtype == ref.dtype, name
        np.testing.assert_allclose(got, ref, atol=_tol(dtype) * scale, err_msg=name)
        truth = OS.stft_rfft(x, **kwargs)
        np.testing.assert_allclose(truth, ref, atol=_tol(dtype) * scale, err_msg=name)
        inv_ref = golden(f'stft/{name}/complex_inverse')
        inv = port.inverse(torch.from_numpy(ref)).numpy()
        np.testing.assert_allclose(inv, inv_ref, atol=_tol(dtype) * np.abs(inv_ref).max(), err_msg=name)
        truth = OS.istft_rfft(ref, **{k: v for k, v in kwargs.items() if k != 'pad'})
        np.testing.assert_allclose(truth, inv_ref, atol=_tol(dtype) * np.abs(inv_ref).max(), err_msg=name)
        for rep in ('concat', 'stacked'):
            if golden.has(f'stft/{name}/{rep}'):
                got = OS.ReferenceSTFT(complex_representation=rep, **kwargs)(torch.from_numpy(x))
                np.testing.assert_allclose(got.numpy(), golden(f'stft/{name}/{rep}'),
                                           atol=_tol(dtype) * scale, err_msg=name)
        for samples, frames in entry['frames'].items():
            assert port.samples_to_frames(int(samples)) == frames, (name, samples)
        for frames, samples in entry['frames_to_samples'].items():
            assert port.frames_to_samples(int(frames)) == samples, (name, frames)


def test_golden_regression(golden):
    fns = dict(mse_loss=L.mse_loss, log_mse_loss=L.log_mse_loss, log1p_mse_loss=L.log1p_mse_loss,
               sdr_loss=L.sdr_loss, si_sdr_loss=L.si_sdr_loss,
               source_aggregated_sdr_loss=L.source_aggregated_sdr_loss)
    for sname in ('k2_t4000', 'k3_t1000', 'b4_k2_t501', 'vec_t100'):
        e = torch.from_numpy(golden(f'regression/{sname}/estimate'))
        t = torch.from_numpy(golden(f'regression/{sname}/target'))
        for fname, variants in golden.index['regression'].items():
            for v, kwargs in variants.items():
                ref = golden(f'regression/{sname}/{fname}/{v}')
                got = fns[fname](e, t, **kwargs).numpy()
                np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6, err_msg=f'{sname} {fname} {kwargs}')
                # float64 truth agrees with the float32 reference to float32 noise
                got64 = fns[fname](e.double(), t.double(), **kwargs).numpy()
                np.testing.assert_allclose(got64, ref, rtol=1e-4, atol=1e-5)


def _pit_fns():
    return dict(mse=torch.nn.functional.mse_loss, pt_mse=L.mse_loss, log_mse=L.log_mse_loss,
                log1p_mse=L.log1p_mse_loss, sdr=L.sdr_loss, si_sdr=L.si_sdr_loss)


def test_golden_pit(golden):
    fns = _pit_fns()
    for name, entry in golden.index['pit'].items():
        e = torch.from_numpy(golden(f'pit/{name}/estimate'))
        t = torch.from_numpy(golden(f'pit/{name}/target'))
        for lname in entry['loss_fns']:
            loss, perm = L.pit_loss(e, t, axis=entry['axis'], loss_fn=fns[lname], return_permutation=True)
            np.testing.assert_allclose(loss.numpy(), golden(f'pit/{name}/{lname}/loss'), rtol=1e-5, atol=1e-6)
            assert list(perm) == list(golden(f'pit/{name}/{lname}/perm')), (name, lname)
            if golden.has(f'pit/{name}/{lname}/pairwise'):
                m = L.compute_pairwise_losses(e, t, axis=entry['axis'], loss_fn=fns[lname])
                np.testing.assert_allclose(m.numpy(), golden(f'pit/{name}/{lname}/pairwise'), rtol=1e-5, atol=1e-6)
                val, cols = L.pit_loss_from_loss_matrix(m, reduction='sum', return_permutation=True)
                np.testing.assert_allclose(val.numpy(), golden(f'pit/{name}/{lname}/matrix_sum'), rtol=1e-5, atol=1e-6)
                assert list(cols) == list(golden(f'pit/{name}/{lname}/matrix_cols'))


def test_golden_dc(golden):
    for name in golden.index['dc']:
        x = torch.from_numpy(golden(f'dc/{name}/x'))
        t = torch.from_numpy(golden(f'dc/{name}/t'))
        np.testing.assert_allclose(L.deep_clustering_loss(x, t).numpy(), golden(f'dc/{name}/loss'), rtol=1e-5)


def test_golden_models(golden):
    meta = golden.index['models']
    n = len(meta['pit']['lengths'])
    load = lambda fmt: [torch.from_numpy(golden(fmt.format(b))) for b in range(n)]  # noqa: E731
    out = P.pit_review_losses(load('models/pit/mask_{}'), load('models/pit/y_abs_{}'),
                              load('models/pit/x_abs_{}'), load('models/pit/cpd_{}'))
    np.testing.assert_allclose(out['pit_mse_loss'].numpy(), golden('models/pit/pit_mse_loss'), rtol=1e-6)
    np.testing.assert_allclose(out['pit_ips_loss'].numpy(), golden('models/pit/pit_ips_loss'), rtol=1e-6)
    mean, _ = P.dc_review_loss(load('models/dc/embedding_{}'), load('models/dc/target_mask_{}'))
    np.testing.assert_allclose(mean.numpy(), golden('models/dc/dc_loss'), rtol=1e-5)
    out = P.tasnet_losses(torch.from_numpy(golden('models/tasnet/estimate')),
                          torch.from_numpy(golden('models/tasnet/s')), meta['tasnet']['num_samples'])
    for key, value in out.items():
        np.testing.assert_allclose(value.numpy(), golden(f'models/tasnet/{key}'), rtol=1e-5, atol=1e-6)


def test_golden_step(golden):
    y, s = torch.from_numpy(golden('step/y')), torch.from_numpy(golden('step/s'))
    masks = torch.from_numpy(golden('step/masks'))
    loss, perms, y_abs = P.stft_mask_pit_step(y, s, masks)
    np.testing.assert_allclose(loss.numpy(), golden('step/loss'), rtol=1e-5)
    assert [list(p) for p in perms] == golden('step/perm').tolist()
    assert golden('step/perm').tolist()[1] == [1, 0]          # the swapped example is in the fixture
    np.testing.assert_allclose(y_abs.numpy(), golden('step/Y_abs'), atol=2e-5 * golden('step/Y_abs').max())


def test_fading_and_tail_pad_arithmetic():
    assert OS.fading_pad_widths(1024, 256, 'full') == (768, 768)
    assert OS.fading_pad_widths(40, 21, 'half') == (9, math.ceil(19 / 2))
    assert OS.fading_pad_widths(40, 20, None) == (0, 0)
    assert OS.tail_pad(300, 1024, 256, True) == 724
    assert OS.tail_pad(1025, 1024, 256, True) == 255
    assert OS.tail_pad(1025, 1024, 256, False) == 0
    assert OS.samples_to_frames(64000, 1024, 256) == 253
    assert OS.samples_to_frames(128000, 1024, 256) == 503
    np.testing.assert_array_equal(OS.samples_to_frames(np.array([203, 150]), 20, 10, True, False), [20, 14])
