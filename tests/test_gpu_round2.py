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
