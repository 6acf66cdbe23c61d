"""Drop-in integration with the UNMODIFIED reference package (pip-installed into baseline/_ref by
`pip install --no-deps --target baseline/_ref`; its un-vendored dependencies come from the
stand-ins in oracle/ref_standins).  Skipped when the reference is not installed.

* CPU: patch_padertorch() swaps the documented attributes, routes CPU tensors to the reference's own
  functions (bit-identical results), unpatch restores everything.
* GPU: the reference's PermutationInvariantTrainingModel and its unmodified Trainer.test_run drive our
  kernels on cuda:0 (BASELINE.json config 1 on the device), and a training step through the patched ops
  equals the same step through the reference ops.
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
STANDINS = os.path.join(ROOT, 'oracle', 'ref_standins')


def import_reference():
    if not os.path.isdir(os.path.join(REF, 'padertorch')):
        pytest.skip('reference not installed in baseline/_ref')
    for path in (STANDINS, REF):
        if path not in sys.path:
            sys.path.insert(0, path)
    import warnings
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    import padertorch as pt
    return pt


@pytest.fixture()
def patched():
    pt = import_reference()
    import padertorch_b200 as b2s
    names = b2s.patch_padertorch(pt)
    yield pt, b2s, names
    b2s.unpatch_padertorch()


def make_examples(n, stft, K=2, seconds=(1.0, 0.8, 0.9, 0.7, 1.0, 0.6), device='cpu', seed=0):
    """Single-mixture examples in the format of pit/data.py:49-77 (lists of per-utterance tensors)."""
    rng = np.random.RandomState(seed)
    examples = []
    for i in range(n):
        T = int(16000 * seconds[i % len(seconds)])
        s = (0.1 * rng.randn(K, T)).astype(np.float32)
        y = s.sum(0)
        Y = stft(torch.from_numpy(y).to(device))
        X = stft(torch.from_numpy(s).to(device)).transpose(0, 1)
        cpd = torch.cos(torch.angle(Y[:, None, :]) - torch.angle(X))
        examples.append(dict(Y_abs=[Y.abs().cpu()], X_abs=[X.abs().cpu()], cos_phase_difference=[cpd.cpu()]))
    return examples


def test_patch_routes_and_restores(patched):
    pt, b2s, names = patched
    assert 'padertorch.ops.losses.source_separation.pit_loss' in names
    assert 'padertorch.ops._stft.STFT' in names and 'padertorch.ops.STFT' in names
    assert any(n.endswith('regression.si_sdr_loss') for n in names)
    # CPU tensors -> the reference's own implementation, bit identical
    rng = np.random.RandomState(0)
    e, t = torch.from_numpy(rng.randn(2, 500).astype(np.float32)), torch.from_numpy(rng.randn(2, 500).astype(np.float32))
    routed = pt.ops.losses.si_sdr_loss
    assert routed.__wrapped_reference__ is not routed
    assert torch.equal(routed(e, t), routed.__wrapped_reference__(e, t))
    est, tgt = torch.rand(20, 2, 17), torch.rand(20, 2, 17)
    assert torch.equal(pt.ops.losses.pit_loss(est, tgt, axis=-2),
                       pt.ops.losses.pit_loss.__wrapped_reference__(est, tgt, axis=-2))
    x = torch.from_numpy(rng.randn(2, 3000).astype(np.float32))
    stft = pt.ops.STFT(512, 128)
    assert isinstance(stft, b2s.ops.STFT)
    from oracle.stft import ReferenceSTFT
    np.testing.assert_allclose(torch.view_as_real(stft(x)).numpy(),
                               torch.view_as_real(ReferenceSTFT(512, 128)(x)).numpy(), atol=1e-4)
    assert stft.inverse(stft(x)).shape[-1] >= 3000
    b2s.unpatch_padertorch()
    assert not hasattr(pt.ops.losses.pit_loss, '__wrapped_reference__')
    assert pt.ops.STFT is pt.ops._stft.STFT and not issubclass(pt.ops.STFT, b2s.ops.STFT)


@pytest.mark.gpu
def test_reference_trainer_test_run_on_gpu(patched):
    """BASELINE.json config 1, on the device: ops.STFT(1024, 256) -> 2-speaker BLSTM mask estimator ->
    pit_loss, driven by the reference's own Trainer.test_run."""
    pt, b2s, _ = patched
    from padertorch.contrib.examples.source_separation.pit.model import PermutationInvariantTrainingModel
    stft = pt.ops.STFT(1024, 256)
    examples = make_examples(6, stft, device='cuda:0')
    assert examples[0]['Y_abs'][0].shape[-1] == 513
    calls = dict(n=0)
    from padertorch_b200.ops.losses import _sse
    original = _sse.SseProblem.forward

    def counting(self, *args, **kwargs):
        calls['n'] += 1
        return original(self, *args, **kwargs)
    _sse.SseProblem.forward = counting
    try:
        torch.manual_seed(0)
        model = PermutationInvariantTrainingModel(F=513, recurrent_layers=2, units=64, K=2)
        with tempfile.TemporaryDirectory() as tmp:
            trainer = pt.Trainer(model, tmp, optimizer=pt.optimizer.Adam(),
                                 loss_weights={'pit_ips_loss': 1.0, 'pit_mse_loss': 0.0},
                                 stop_trigger=(2, 'iteration'))
            trainer.test_run(examples[:4], examples[4:], device=0)
    finally:
        _sse.SseProblem.forward = original
    # the kernels, not the reference loop, computed the losses: with the model patch ONE dual launch per review
    # (2 train + 2 validation reviews, twice for the determinism checks of test_run), not 2 per example
    assert calls['n'] >= 4, calls


@pytest.mark.gpu
def test_patched_training_step_matches_reference_ops(patched):
    pt, b2s, _ = patched
    from padertorch.contrib.examples.source_separation.pit.model import PermutationInvariantTrainingModel
    stft = pt.ops.STFT(1024, 256)
    example = make_examples(1, stft, device='cuda:0')[0]
    torch.manual_seed(1)
    model = PermutationInvariantTrainingModel(F=513, recurrent_layers=1, units=32, K=2).cuda()
    batch = model.example_to_device(example, 0)

    def step():
        model.zero_grad()
        review = model.review(batch, model(batch))
        loss = review['losses']['pit_mse_loss'] + review['losses']['pit_ips_loss']
        loss.backward()
        return loss.detach().clone(), [p.grad.detach().clone() for p in model.parameters()]

    loss_ours, grads_ours = step()
    b2s.unpatch_padertorch()
    loss_ref, grads_ref = step()            # same model, reference ATen ops on the GPU
    torch.testing.assert_close(loss_ours, loss_ref, rtol=1e-4, atol=1e-7)
    for a, b in zip(grads_ours, grads_ref):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=1e-6 + 1e-4 * float(b.abs().max()))


def _launch_counter():
    """Counts C-ABI kernel entry points called through padertorch_b200._lib (every call = >= 1 launch)."""
    from padertorch_b200 import _lib
    lib = _lib.load()
    counts = {}

    class Counting:
        def __getattr__(self, name):
            fn = getattr(lib, name)
            if not name.startswith('b2s_') or 'workspace' in name or 'plan' in name or 'last_error' in name:
                return fn

            def wrapped(*args):
                counts[name] = counts.get(name, 0) + 1
                return fn(*args)
            return wrapped
    return Counting(), counts


@pytest.mark.gpu
def test_model_patch_pit_review_batched(patched):
    """patch_padertorch(models=True): PermutationInvariantTrainingModel.review on a 4-utterance batch equals
    the reference's per-example loop (losses + gradients w.r.t. the masks) and costs ONE loss launch in the
    forward and one in the backward instead of 2 * B."""
    pt, b2s, names = patched
    from padertorch.contrib.examples.source_separation.pit.model import PermutationInvariantTrainingModel
    from padertorch_b200 import _lib
    assert any(n.endswith('PermutationInvariantTrainingModel.review') for n in names)
    stft = pt.ops.STFT(1024, 256)
    parts = make_examples(4, stft, device='cuda:0', seed=3)
    parts.sort(key=lambda ex: -ex['Y_abs'][0].shape[0])      # pack_sequence wants decreasing lengths
    batch = {key: [ex[key][0].cuda() for ex in parts] for key in ('Y_abs', 'X_abs', 'cos_phase_difference')}
    torch.manual_seed(2)
    model = PermutationInvariantTrainingModel(F=513, recurrent_layers=1, units=32, K=2).cuda()

    def step():
        model.zero_grad()
        out = model(batch)
        review = model.review(batch, out)
        (review['losses']['pit_mse_loss'] + 2 * review['losses']['pit_ips_loss']).backward()
        return ({k: v.detach().clone() for k, v in review['losses'].items()}, sorted(review['images']),
                [p.grad.detach().clone() for p in model.parameters()])

    counting, counts = _launch_counter()
    real_load = _lib.load
    _lib.load = lambda: counting
    try:
        losses_ours, images_ours, grads_ours = step()
    finally:
        _lib.load = real_load
    # one loss launch (+ the one-warp mean launch behind it) forward, one backward
    assert counts.get('b2s_pit_sse_forward_mean') == 1 and counts.get('b2s_pit_sse_backward_scaled') == 1, counts
    assert 'b2s_pit_sse_forward' not in counts and 'b2s_pit_sse_backward' not in counts, counts
    b2s.unpatch_padertorch()
    losses_ref, images_ref, grads_ref = step()          # the unmodified reference on the GPU
    assert images_ours == images_ref
    for key in losses_ref:
        torch.testing.assert_close(losses_ours[key], losses_ref[key], rtol=1e-4, atol=1e-7)
    for a, b in zip(grads_ours, grads_ref):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=1e-6 + 1e-4 * float(b.abs().max()))


@pytest.mark.gpu
def test_model_patch_dc_review_and_tasnet_loss(patched):
    """DeepClusteringModel.review (tcl/dc.py:76-84) and TasNet.loss (tasnet/model.py:154-176) through the
    model patch against the unpatched reference methods on the same GPU tensors."""
    pt, b2s, names = patched
    from padertorch.contrib.tcl.dc import DeepClusteringModel
    from padertorch.contrib.examples.source_separation.tasnet.model import TasNet
    rng = np.random.RandomState(4)
    frames = [40, 31, 25]
    emb = [torch.nn.functional.normalize(torch.from_numpy(rng.randn(t, 20, 513).astype(np.float32)).cuda(), dim=-2)
           .requires_grad_(True) for t in frames]
    tgt = [torch.nn.functional.one_hot(torch.from_numpy(rng.randint(0, 2, size=(t, 513))), 2)
           .permute(0, 2, 1).float().contiguous().cuda() for t in frames]
    ours = DeepClusteringModel.review(None, {'target_mask': tgt}, emb)['losses']['dc_loss']
    ours.backward()
    grads_ours = [e.grad.clone() for e in emb]
    for e in emb:
        e.grad = None
    est = torch.from_numpy(rng.randn(3, 2, 6000).astype(np.float32)).cuda().requires_grad_(True)
    src = torch.from_numpy(rng.randn(3, 2, 6000).astype(np.float32)).cuda()
    inputs, outputs = {'s': src, 'num_samples': [6000, 5000, 4321]}, {'out': est}
    tas_ours = TasNet.loss(None, inputs, outputs)
    sum(tas_ours.values()).backward()
    tas_grad = est.grad.clone()
    est.grad = None
    b2s.unpatch_padertorch()
    ref = DeepClusteringModel.review(None, {'target_mask': tgt}, emb)['losses']['dc_loss']
    ref.backward()
    torch.testing.assert_close(ours.detach(), ref.detach(), rtol=1e-4, atol=1e-9)
    for a, e in zip(grads_ours, emb):
        torch.testing.assert_close(a, e.grad, rtol=1e-3, atol=1e-4 * float(e.grad.abs().max()))
    tas_ref = TasNet.loss(None, inputs, outputs)
    sum(tas_ref.values()).backward()
    assert sorted(tas_ours) == sorted(tas_ref)
    for key in tas_ref:
        torch.testing.assert_close(tas_ours[key].detach(), tas_ref[key].detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(tas_grad, est.grad, rtol=1e-3, atol=1e-4 * float(est.grad.abs().max()))


@pytest.mark.gpu
def test_reference_stft_coders_through_the_patch(patched):
    """The reference's own StftEncoder / IstftDecoder (tasnet/tas_coders.py:175-192, 229-240; doctests :140-155,
    :197-208) constructed AFTER the patch run the kernels (size 256 / shift 10 / window 20: the table-driven DFT
    path) and agree with the unpatched modules, values and autograd."""
    pt, b2s, names = patched
    from padertorch.contrib.examples.source_separation.tasnet import tas_coders
    assert issubclass(tas_coders.STFT, b2s.ops.STFT)
    torch.manual_seed(0)
    mixture = torch.rand((2, 6, 203))
    spec = torch.rand((2, 4, 258, 10))
    enc, dec = tas_coders.StftEncoder(feature_size=258), tas_coders.IstftDecoder(feature_size=258)
    x = mixture.cuda().requires_grad_(True)
    encoded, num_frames = enc(x, [203, 150])
    assert encoded.shape == (2, 6, 258, 20) and num_frames.tolist() == [20, 14]
    encoded.square().sum().backward()
    z = spec.cuda().requires_grad_(True)
    decoded = dec(z)
    assert decoded.shape == (2, 4, 110)
    decoded.square().sum().backward()
    b2s.unpatch_padertorch()
    enc_ref, dec_ref = tas_coders.StftEncoder(feature_size=258), tas_coders.IstftDecoder(feature_size=258)
    x_ref = mixture.clone().requires_grad_(True)
    encoded_ref, frames_ref = enc_ref(x_ref, [203, 150])
    encoded_ref.square().sum().backward()
    z_ref = spec.clone().requires_grad_(True)
    decoded_ref = dec_ref(z_ref)
    decoded_ref.square().sum().backward()
    assert frames_ref.tolist() == num_frames.tolist()
    torch.testing.assert_close(encoded.detach().cpu(), encoded_ref.detach(), rtol=0, atol=1e-4 * float(encoded_ref.abs().max()))
    torch.testing.assert_close(decoded.detach().cpu(), decoded_ref.detach(), rtol=0, atol=1e-4 * float(decoded_ref.abs().max()))
    torch.testing.assert_close(x.grad.cpu(), x_ref.grad, rtol=0, atol=1e-4 * float(x_ref.grad.abs().max()))
    torch.testing.assert_close(z.grad.cpu(), z_ref.grad, rtol=0, atol=1e-4 * float(z_ref.grad.abs().max()))
