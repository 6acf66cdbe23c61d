#!/usr/bin/env python
"""Record golden fixtures from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz + index.json

The reference is imported from ``/root/reference`` with the stand-ins under
``oracle/ref_standins`` for its un-vendored dependencies.  Everything recorded here is an
output of reference code (``padertorch.ops.STFT``, ``padertorch.ops.losses.*`` and the
``review`` / ``loss`` methods of the three hot-path models); inputs are seeded and stored
next to the outputs so the fixtures are self-contained.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get('PADERTORCH_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'ref_standins'))
sys.path.insert(0, REFERENCE)
warnings.filterwarnings('ignore', category=SyntaxWarning)

import torch  # noqa: E402
import padertorch as pt  # noqa: E402
from padertorch.ops import STFT  # noqa: E402
from padertorch.ops.losses import regression as R  # noqa: E402
from padertorch.ops.losses import source_separation as S  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

STFT_CASES = [
    # name, ctor kwargs, input shape, dtype
    ('hann4', dict(size=4, shift=2, window='hann', fading='full'), None, 'float32'),
    ('default_1024_256', dict(size=1024, shift=256), (2, 3000), 'float32'),
    ('default_1024_256_f64', dict(size=1024, shift=256), (2, 2500), 'float64'),
    ('small_512_20_40_hamming', dict(size=512, shift=20, window_length=40, window='hamming'),
     (2, 3, 203), 'float32'),
    ('encoder_256_10_20_nofade', dict(size=256, shift=10, window_length=20, fading=False),
     (2, 203), 'float32'),
    ('half_512_128', dict(size=512, shift=128, fading='half'), (3, 1000), 'float32'),
    ('nopad_512_128', dict(size=512, shift=128, fading=None, pad=False), (2, 1111), 'float32'),
    ('symmetric_256_64', dict(size=256, shift=64, symmetric_window=True, window='hann'),
     (2, 777), 'float32'),
    ('short_input_1024', dict(size=1024, shift=256, fading=False), (2, 300), 'float32'),
    ('wl_odd_shift_512_48_96', dict(size=512, shift=48, window_length=96, window='hann'),
     (2, 1000), 'float32'),
    ('size2048_512', dict(size=2048, shift=512), (1, 6000), 'float32'),
    ('size64_16', dict(size=64, shift=16, window='hamming'), (4, 500), 'float32'),
    ('size8_shift1', dict(size=8, shift=1, window='hann', fading=None), (2, 40), 'float32'),
    ('nonpow2_size100_shift25', dict(size=100, shift=25, window='hann'), (2, 333), 'float32'),
]


ALL_REPRESENTATIONS = ('hann4', 'small_512_20_40_hamming', 'encoder_256_10_20_nofade')


def record_stft(store, index):
    rng = np.random.RandomState(0)
    for name, kwargs, shape, dtype in STFT_CASES:
        if shape is None:
            x = np.arange(8).astype(dtype)
        else:
            x = rng.randn(*shape).astype(dtype)
        entry = dict(kwargs=kwargs, dtype=dtype, frames={})
        xt = torch.from_numpy(x).requires_grad_(True)
        reps = ('complex', 'concat', 'stacked') if name in ALL_REPRESENTATIONS else ('complex',)
        for rep in reps:
            stft = STFT(complex_representation=rep, **kwargs)
            out = stft(xt)
            store[f'stft/{name}/{rep}'] = out.detach().numpy()
            back = stft.inverse(out.detach())
            store[f'stft/{name}/{rep}_inverse'] = back.numpy()
        # autograd of the reference: d/dx sum(Re(Y) Gr + Im(Y) Gi), d/dY sum(istft(Y) g)
        stft = STFT(complex_representation='stacked', **kwargs)
        out = stft(xt)
        g = rng.randn(*out.shape).astype(dtype)
        (grad_x,) = torch.autograd.grad((out * torch.from_numpy(g)).sum(), xt)
        store[f'stft/{name}/grad_out_stacked'] = g
        store[f'stft/{name}/grad_x'] = grad_x.numpy()
        spec = torch.from_numpy(rng.randn(*out.shape).astype(dtype)).requires_grad_(True)
        sig = stft.inverse(spec)
        gs = rng.randn(*sig.shape).astype(dtype)
        (grad_spec,) = torch.autograd.grad((sig * torch.from_numpy(gs)).sum(), spec)
        store[f'stft/{name}/inv_in_stacked'] = spec.detach().numpy()
        store[f'stft/{name}/inv_out'] = sig.detach().numpy()
        store[f'stft/{name}/inv_grad_out'] = gs
        store[f'stft/{name}/inv_grad_in'] = grad_spec.numpy()
        store[f'stft/{name}/x'] = x
        for samples in (1, 7, 255, 256, 1023, 1024, 1025, 16000, 64000, 128000):
            entry['frames'][str(samples)] = int(stft.samples_to_frames(samples))
        entry['frames_to_samples'] = {str(m): int(stft.frames_to_samples(m))
                                      for m in (1, 2, 7, 253, 503)}
        index['stft'][name] = entry


REGRESSION = {
    'mse_loss': (R.mse_loss, [dict(), dict(reduction='mean'), dict(reduction=None)]),
    'log_mse_loss': (R.log_mse_loss, [dict(), dict(reduction=None),
                                      dict(soft_sdr_max=20), dict(reduction='mean', soft_sdr_max=30)]),
    'log1p_mse_loss': (R.log1p_mse_loss, [dict(), dict(reduction=None), dict(reduction='mean')]),
    'sdr_loss': (R.sdr_loss, [dict(), dict(reduction=None), dict(reduction='sum', soft_sdr_max=20)]),
    'si_sdr_loss': (R.si_sdr_loss, [dict(), dict(reduction=None), dict(offset_invariant=True),
                                    dict(grad_stop=True), dict(soft_sdr_max=20),
                                    dict(reduction='sum', offset_invariant=True, grad_stop=True,
                                         soft_sdr_max=30)]),
    'source_aggregated_sdr_loss': (R.source_aggregated_sdr_loss, [dict(), dict(soft_sdr_max=20)]),
}


def record_regression(store, index):
    rng = np.random.RandomState(1)
    shapes = {'k2_t4000': (2, 4000), 'k3_t1000': (3, 1000), 'b4_k2_t501': (4, 2, 501),
              'vec_t100': (100,)}
    for sname, shape in shapes.items():
        target = rng.randn(*shape).astype(np.float32)
        estimate = (target + 0.3 * rng.randn(*shape) + 0.05).astype(np.float32)
        store[f'regression/{sname}/estimate'] = estimate
        store[f'regression/{sname}/target'] = target
        for fname, (fn, variants) in REGRESSION.items():
            for v, kwargs in enumerate(variants):
                e = torch.from_numpy(estimate).requires_grad_(True)
                t = torch.from_numpy(target)
                out = fn(e, t, **kwargs)
                key = f'regression/{sname}/{fname}/{v}'
                store[key] = out.detach().numpy()
                (grad,) = torch.autograd.grad(out.sum(), e)
                store[key + '/grad'] = grad.numpy()
                index['regression'].setdefault(fname, {})[str(v)] = kwargs


PIT_LOSS_FNS = {
    'mse': torch.nn.functional.mse_loss,
    'pt_mse': R.mse_loss,
    'log_mse': R.log_mse_loss,
    'log1p_mse': R.log1p_mse_loss,
    'sdr': R.sdr_loss,
    'si_sdr': R.si_sdr_loss,
}


def record_pit(store, index):
    rng = np.random.RandomState(2)
    cases = {
        # name: (shape, axis)
        'tkf_k2': ((37, 2, 65), -2),
        'tkf_k3': ((29, 3, 33), 1),
        'kt_k2': ((2, 3000), 0),
        'kt_k3': ((3, 1777), 0),
        'kft_k2': ((2, 17, 50), 0),
        'abkcf_k3': ((2, 3, 3, 5, 16), -3),
        'k4_vec': ((4, 64), 0),
    }
    for name, (shape, axis) in cases.items():
        k = shape[axis]
        target = np.abs(rng.randn(*shape)).astype(np.float32)
        order = rng.permutation(k)
        estimate = np.take(target, order, axis=axis) + 0.2 * rng.randn(*shape).astype(np.float32)
        estimate = estimate.astype(np.float32)
        store[f'pit/{name}/estimate'] = estimate
        store[f'pit/{name}/target'] = target
        index['pit'][name] = dict(axis=axis, shape=list(shape), loss_fns=[])
        for lname, fn in PIT_LOSS_FNS.items():
            if lname != 'mse' and (axis % len(shape)) != 0:
                continue  # regression losses reduce the last axis; speakers must lead
            if lname in ('si_sdr',) and len(shape) >= 2 and shape[-2] >= 10:
                continue
            e = torch.from_numpy(estimate).requires_grad_(True)
            t = torch.from_numpy(target)
            loss, perm = S.pit_loss(e, t, axis=axis, loss_fn=fn, return_permutation=True)
            (grad,) = torch.autograd.grad(loss, e)
            store[f'pit/{name}/{lname}/loss'] = loss.detach().numpy()
            store[f'pit/{name}/{lname}/perm'] = np.asarray(perm, dtype=np.int64)
            store[f'pit/{name}/{lname}/grad'] = grad.numpy()
            index['pit'][name]['loss_fns'].append(lname)
            if len(shape) == 2 or lname == 'mse':
                matrix = S.compute_pairwise_losses(e.detach(), t, axis=axis, loss_fn=fn)
                store[f'pit/{name}/{lname}/pairwise'] = matrix.numpy()
                for red in ('mean', 'sum'):
                    val, cols = S.pit_loss_from_loss_matrix(matrix, reduction=red,
                                                            return_permutation=True)
                    store[f'pit/{name}/{lname}/matrix_{red}'] = val.numpy()
                store[f'pit/{name}/{lname}/matrix_cols'] = np.asarray(cols, dtype=np.int64)


def record_dc(store, index):
    rng = np.random.RandomState(3)
    for name, (n, e_dim, k) in {'n100_e20_k3': (100, 20, 3), 'n4104_e20_k2': (4104, 20, 2),
                                'n999_e7_k4': (999, 7, 4)}.items():
        x = rng.randn(n, e_dim).astype(np.float32)
        x /= np.linalg.norm(x, axis=-1, keepdims=True)
        t = np.eye(k, dtype=np.float32)[rng.randint(0, k, size=n)]
        xt = torch.from_numpy(x).requires_grad_(True)
        loss = S.deep_clustering_loss(xt, torch.from_numpy(t))
        (grad,) = torch.autograd.grad(loss, xt)
        store[f'dc/{name}/x'] = x
        store[f'dc/{name}/t'] = t
        store[f'dc/{name}/loss'] = loss.detach().numpy()
        store[f'dc/{name}/grad'] = grad.numpy()
        index['dc'][name] = dict(N=n, E=e_dim, K=k)


def record_models(store, index):
    """review()/loss() of the three hot-path models on random ragged batches
    (shapes follow tests/test_models/test_bss.py:21-41)."""
    from padertorch.contrib.examples.source_separation.pit.model import (
        PermutationInvariantTrainingModel)
    from padertorch.contrib.tcl.dc import DeepClusteringModel
    from padertorch.contrib.examples.source_separation.tasnet.model import TasNet

    rng = np.random.RandomState(4)
    lengths, F_, K, E = [40, 36, 31, 17], 33, 2, 20
    masks = [rng.rand(t, K, F_).astype(np.float32) for t in lengths]
    y_abs = [np.abs(rng.randn(t, F_)).astype(np.float32) for t in lengths]
    x_abs = [np.abs(rng.randn(t, K, F_)).astype(np.float32) for t in lengths]
    cpd = [np.cos(rng.uniform(-np.pi, np.pi, size=(t, K, F_))).astype(np.float32)
           for t in lengths]
    batch = dict(Y_abs=[torch.from_numpy(a) for a in y_abs],
                 X_abs=[torch.from_numpy(a) for a in x_abs],
                 cos_phase_difference=[torch.from_numpy(a) for a in cpd])
    model_out = [torch.from_numpy(m).requires_grad_(True) for m in masks]
    review = PermutationInvariantTrainingModel.review(None, batch, model_out)
    total = review['losses']['pit_mse_loss'] + review['losses']['pit_ips_loss']
    grads = torch.autograd.grad(total, model_out)
    for b, t in enumerate(lengths):
        store[f'models/pit/mask_{b}'] = masks[b]
        store[f'models/pit/y_abs_{b}'] = y_abs[b]
        store[f'models/pit/x_abs_{b}'] = x_abs[b]
        store[f'models/pit/cpd_{b}'] = cpd[b]
        store[f'models/pit/grad_mask_{b}'] = grads[b].numpy()
    store['models/pit/pit_mse_loss'] = review['losses']['pit_mse_loss'].detach().numpy()
    store['models/pit/pit_ips_loss'] = review['losses']['pit_ips_loss'].detach().numpy()
    index['models']['pit'] = dict(lengths=lengths, F=F_, K=K)

    emb = [rng.randn(t, E, F_).astype(np.float32) for t in lengths]
    emb = [e / np.linalg.norm(e, axis=1, keepdims=True) for e in emb]
    tmask = [np.moveaxis(np.eye(K, dtype=np.float32)[rng.randint(0, K, size=(t, F_))], -1, 1)
             .copy() for t in lengths]
    emb_t = [torch.from_numpy(e).requires_grad_(True) for e in emb]
    review = DeepClusteringModel.review(None, dict(target_mask=[torch.from_numpy(m) for m in tmask]),
                                        emb_t)
    grads = torch.autograd.grad(review['losses']['dc_loss'], emb_t)
    for b in range(len(lengths)):
        store[f'models/dc/embedding_{b}'] = emb[b]
        store[f'models/dc/target_mask_{b}'] = tmask[b]
        store[f'models/dc/grad_embedding_{b}'] = grads[b].numpy()
    store['models/dc/dc_loss'] = review['losses']['dc_loss'].detach().numpy()
    index['models']['dc'] = dict(lengths=lengths, F=F_, K=K, E=E)

    num_samples = [4000, 3600, 3001, 1777]
    s = rng.randn(len(num_samples), K, max(num_samples)).astype(np.float32)
    est = (s[:, ::-1] + 0.4 * rng.randn(*s.shape)).astype(np.float32)
    est[2] = (s[2] + 0.4 * rng.randn(*s[2].shape)).astype(np.float32)
    est_t = torch.from_numpy(est.copy()).requires_grad_(True)
    out = TasNet.loss(None, dict(s=torch.from_numpy(s), num_samples=num_samples),
                      dict(out=est_t))
    for name, value in out.items():
        store[f'models/tasnet/{name}'] = value.detach().numpy()
        (grad,) = torch.autograd.grad(value, est_t, retain_graph=True)
        store[f'models/tasnet/{name}/grad'] = grad.numpy()
    store['models/tasnet/s'] = s
    store['models/tasnet/estimate'] = est
    index['models']['tasnet'] = dict(num_samples=num_samples, K=K)


def record_step(store, index):
    """The bench step on a small batch, produced by reference code only:
    ops.STFT -> abs -> mask * Y_abs -> pit_loss (pit/model.py:117-128, pit/data.py:49-77)."""
    rng = np.random.RandomState(5)
    B, K, T = 3, 2, 6000
    s = (0.1 * rng.randn(B, K, T)).astype(np.float32)
    y = s.sum(1)
    stft = STFT(1024, 256)
    Y = stft(torch.from_numpy(y))
    X = stft(torch.from_numpy(s)).transpose(1, 2)       # [B, M, K, F]
    M, Fb = Y.shape[-2:]
    masks = rng.rand(B, M, K, Fb).astype(np.float32)
    losses, perms = [], []
    for b in range(B):
        if b == 1:   # make the swapped assignment the winner for one example
            masks[b] = masks[b][:, ::-1] * 0 + (X[b].abs().numpy()[:, ::-1]
                                                / (Y[b].abs().numpy()[:, None] + 1e-3))
        loss, perm = S.pit_loss(torch.from_numpy(masks[b]) * Y[b].abs()[:, None, :], X[b].abs(),
                                axis=-2, return_permutation=True)
        losses.append(loss.numpy())
        perms.append(perm)
    store['step/y'] = y
    store['step/s'] = s
    store['step/masks'] = masks
    store['step/Y_abs'] = Y.abs().numpy()
    store['step/X_abs'] = X.abs().numpy()
    store['step/loss'] = np.stack(losses)
    store['step/perm'] = np.asarray(perms, dtype=np.int64)
    index['step'] = dict(B=B, K=K, T=T, M=int(M), F=int(Fb))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)          # reproducible reduction order
    index = dict(stft={}, regression={}, pit={}, dc={}, models={}, step={},
                 reference_commit='ca62dbc', torch=torch.__version__, numpy=np.__version__)
    groups = dict(stft=record_stft, regression=record_regression, pit=record_pit,
                  dc=record_dc, models=record_models, step=record_step)
    for group, fn in groups.items():
        store = {}
        fn(store, index)
        path = os.path.join(OUT, f'{group}.npz')
        np.savez_compressed(path, **store)
        print(f'{path}: {len(store)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')
    with open(os.path.join(OUT, 'index.json'), 'w') as fd:
        json.dump(index, fd, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
