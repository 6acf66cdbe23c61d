"""Drop-in switch: route an importable ``padertorch`` (its models and its unmodified Trainer) through
the kernels of this package.

The reference has no plugin registry on this path; its boundary is attribute lookup at call time
(``pt.ops.losses.pit_loss(...)`` in pit/model.py:124, ``pt.ops.losses.deep_clustering_loss`` in
tcl/dc.py:79, ``pt.pit_loss`` / ``pt.ops.STFT`` in tasnet/model.py:169 and tas_coders.py:170), so
the integration is a set of module-attribute replacements.  CUDA float32 tensors take
the kernels; anything else (CPU tensors, float64) is handed to the original reference function, which
is the reference's behaviour, not a fallback of ours.
"""
import functools

import torch

_ORIGINALS = {}

_LOSS_NAMES = ['mse_loss', 'log_mse_loss', 'sdr_loss', 'si_sdr_loss', 'log1p_mse_loss',
               'source_aggregated_sdr_loss', 'deep_clustering_loss', 'pit_loss',
               'compute_pairwise_losses', 'pit_loss_from_loss_matrix']


def _on_device(*tensors):
    return all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in tensors)


def _route(ours, theirs, handles=None):
    """`ours` for float32 CUDA tensors (and, if given, when ``handles(*args, **kwargs)`` says the kernels cover
    the case); everything else is the reference's own function object, untouched."""
    @functools.wraps(theirs)
    def routed(estimate, target, *args, **kwargs):
        if _on_device(estimate, target) and (handles is None or handles(*args, **kwargs)):
            return ours(estimate, target, *args, **kwargs)
        return theirs(estimate, target, *args, **kwargs)
    routed.__wrapped_reference__ = theirs
    return routed


def _route_matrix(ours, theirs):
    """pit_loss_from_loss_matrix: one positional argument (the K x K matrix)."""
    @functools.wraps(theirs)
    def routed(pair_wise_loss_matrix, **kwargs):
        from . import _lib
        if (_on_device(pair_wise_loss_matrix) and pair_wise_loss_matrix.dim() == 2
                and pair_wise_loss_matrix.shape[-1] <= _lib.MAX_SOURCES
                and kwargs.get('algorithm', 'optimal') in ('optimal', 'hungarian', 'brute_force')):
            return ours(pair_wise_loss_matrix, **kwargs)
        return theirs(pair_wise_loss_matrix, **kwargs)
    routed.__wrapped_reference__ = theirs
    return routed


def _kernel_loss_fn(axis=None, loss_fn=torch.nn.functional.mse_loss, *args, **kwargs):
    """pit_loss / compute_pairwise_losses: the kernels cover the loss functions registered in _FAST_PIT; an
    opaque callable or cross entropy is the reference's business (its permutation loop, its einsum)."""
    from .ops.losses import source_separation as our_ss
    return our_ss._fast_spec(loss_fn) is not None


def _all_on_device(items):
    return len(items) > 0 and all(_on_device(t) for t in items)


def _patch_models(pt, replace):
    """Model-level drop-in: the per-example Python loops of the three
    hot-path models become one or two launches for the whole minibatch.  Every method keeps its name,
    arguments and return structure; CPU batches run the original method."""
    import importlib
    from . import review as ours

    def module(name):
        try:
            return importlib.import_module(name)
        except Exception:      # an example package whose own dependencies are missing stays untouched
            return None

    pit = module('padertorch.contrib.examples.source_separation.pit.model')
    if pit is not None:
        original_review = pit.PermutationInvariantTrainingModel.review

        def pit_review(self, batch, model_out):
            """pit/model.py:112-151 with the loop :117-135 replaced by review.pit_review_losses (the MSE and
            the ideal-phase-sensitive loss of the whole list in one pass); `images` as in :142-147."""
            if not (_all_on_device(model_out) and _all_on_device(batch['Y_abs'])):
                return original_review(self, batch, model_out)
            losses = ours.pit_review_losses(list(model_out), list(batch['Y_abs']), list(batch['X_abs']),
                                            list(batch['cos_phase_difference']))
            b = 0   # only the first example of a batch is rendered (:142)
            images = {'observation': pit.stft_to_image(batch['Y_abs'][b])}
            for i in range(model_out[b].shape[1]):
                images[f'mask_{i}'] = pit.mask_to_image(model_out[b][:, i, :])
                images[f'estimation_{i}'] = pit.stft_to_image(batch['X_abs'][b][:, 0, :])
            return dict(losses=losses, images=images)

        pit_review.__wrapped_reference__ = original_review
        replace(pit.PermutationInvariantTrainingModel, 'review', pit_review)

    dc = module('padertorch.contrib.tcl.dc')
    if dc is not None:
        original_dc = dc.DeepClusteringModel.review

        def dc_review(self, batch, model_out):
            """tcl/dc.py:76-84: the loop over (embedding, target_mask) pairs -> review.dc_review_loss on the
            model's native 't e f' layout (no rearrange copies)."""
            if not (_all_on_device(model_out) and _all_on_device(batch['target_mask'])):
                return original_dc(self, batch, model_out)
            return {'losses': {'dc_loss': ours.dc_review_loss(list(model_out), list(batch['target_mask']))}}

        dc_review.__wrapped_reference__ = original_dc
        replace(dc.DeepClusteringModel, 'review', dc_review)

    tasnet = module('padertorch.contrib.examples.source_separation.tasnet.model')
    if tasnet is not None:
        original_loss = tasnet.TasNet.loss

        def tasnet_loss(self, inputs, outputs):
            """tasnet/model.py:154-176: 3 loss functions x B examples of pit_loss -> one statistics pass and
            one loss-set launch (review.tasnet_losses)."""
            s, x = inputs['s'], outputs['out']
            if not (_on_device(x) and _on_device(s) and x.shape == s.shape):
                return original_loss(self, inputs, outputs)
            return ours.tasnet_losses(x, s, [int(n) for n in inputs['num_samples']])

        tasnet_loss.__wrapped_reference__ = original_loss
        replace(tasnet.TasNet, 'loss', tasnet_loss)


def patch_padertorch(pt=None, models=True):
    """Replace the hot-path ops of the importable ``padertorch`` package by this package's and, with
    ``models=True``, the per-example loss loops of the three hot-path models (PIT ``review``, deep-clustering
    ``review``, ``TasNet.loss``) by their batched forms.  Returns the list of patched attribute paths.
    Idempotent."""
    if pt is None:
        import padertorch as pt
    from . import ops as ours
    from .ops.losses import source_separation as our_ss
    if _ORIGINALS:
        return sorted(_ORIGINALS)
    ref_reg = pt.ops.losses.regression
    ref_ss = pt.ops.losses.source_separation
    patched = {}

    def replace(module, name, value):
        key = f'{getattr(module, "__module__", None) and module.__module__ + "." + module.__name__ or module.__name__}.{name}'
        if hasattr(module, name):
            _ORIGINALS[key] = (module, name, getattr(module, name))
            setattr(module, name, value)
            patched[key] = value

    for name in _LOSS_NAMES:
        ours_fn = getattr(ours, name, None) or getattr(our_ss, name)
        home = ref_reg if hasattr(ref_reg, name) else ref_ss
        theirs_fn = getattr(home, name)
        if name == 'pit_loss_from_loss_matrix':
            routed = _route_matrix(ours_fn, theirs_fn)
        else:
            routed = _route(ours_fn, theirs_fn,
                            _kernel_loss_fn if name in ('pit_loss', 'compute_pairwise_losses') else None)
        # the reference's own function objects (and our routed wrappers) take the kernel path when
        # they are passed to pit_loss as loss_fn
        if ours_fn in our_ss._FAST_PIT:
            our_ss.register_fast_loss(theirs_fn, ours_fn)
            our_ss.register_fast_loss(routed, ours_fn)
        for module in (home, pt.ops.losses, pt.ops, pt):
            if getattr(module, name, None) is theirs_fn:
                replace(module, name, routed)

    ref_stft_module = pt.ops._stft
    RefSTFT = ref_stft_module.STFT

    class STFT(ours.STFT):
        """padertorch_b200.ops.STFT that hands CPU / float64 inputs to the reference STFT."""

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self._args, self._kwargs, self._reference = args, kwargs, None

        def _ref(self):
            if self._reference is None:
                self._reference = RefSTFT(*self._args, **self._kwargs)
            ref = self._reference
            ref.fading, ref.pad = self.fading, self.pad
            ref.complex_representation = self.complex_representation
            return ref

        def __call__(self, inputs):
            if _on_device(inputs):
                return super().__call__(inputs)
            return self._ref()(inputs)

        def inverse(self, stft_signal):
            real = torch.view_as_real(stft_signal) if torch.is_complex(stft_signal) else stft_signal
            if _on_device(real):
                return super().inverse(stft_signal)
            return self._ref().inverse(stft_signal)

    STFT.__name__ = 'STFT'
    STFT.__qualname__ = 'STFT'
    for module in (ref_stft_module, pt.ops):
        if getattr(module, 'STFT', None) is RefSTFT:
            replace(module, 'STFT', STFT)
    if models:
        _patch_models(pt, replace)
    # modules that bound the names at import time (``from padertorch.ops import STFT`` in tasnet/tas_coders.py:5,
    # ``from padertorch.ops.losses.regression import si_sdr_loss`` ...): rebind them as well
    import sys
    originals = {id(old): new for (module, name, old), new in
                 ((entry, getattr(entry[0], entry[1])) for entry in list(_ORIGINALS.values()))
                 if isinstance(module, type(sys)) }
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not mod_name.startswith('padertorch.') or mod_name.startswith('padertorch_b200'):
            continue
        for attr, value in list(vars(mod).items()):
            new = originals.get(id(value))
            if new is not None and value is not new and f'{mod.__name__}.{attr}' not in _ORIGINALS:
                replace(mod, attr, new)
    return sorted(patched)


def unpatch_padertorch():
    """Undo patch_padertorch()."""
    for module, name, value in _ORIGINALS.values():
        setattr(module, name, value)
    _ORIGINALS.clear()
