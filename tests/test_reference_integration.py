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

    def counting(self):
        calls['n'] += 1
        return original(self)
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
    assert calls['n'] >= 8, calls       # the kernels, not the reference loop, computed the losses


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
