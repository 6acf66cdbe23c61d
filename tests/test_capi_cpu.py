"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/b200sep.h declares (no compute calls -- there is no GPU here), the ctypes table matches the
header, the integer frame arithmetic is bit exact against the reference-recorded fixtures, the product
never imports the oracle, and CPU tensors are refused loudly (no fallback)."""
import ast
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200sep.h')


def header_declarations():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = {}
    for match in re.finditer(r'B2S_API\s+([\w\s\*]+?)\s*\b(b2s_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        ret, name, args = match.groups()
        args = args.strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        decls[name] = (ret.strip(), n)
    return decls


@pytest.fixture(scope='module')
def lib():
    from padertorch_b200 import build
    build.build()
    from padertorch_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from padertorch_b200 import _lib
    decls = header_declarations()
    assert len(decls) >= 25, sorted(decls)
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, (_, nargs) in decls.items():
        assert hasattr(lib, name), name
        assert len(_lib.SIGNATURES[name][1]) == nargs, (name, nargs, len(_lib.SIGNATURES[name][1]))
    assert lib.b2s_version() == 100
    assert lib.b2s_last_error() is not None


def test_library_has_sm100a_code():
    import subprocess
    from padertorch_b200 import build
    out = subprocess.run(['cuobjdump', '--list-elf', build.LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in out.stdout, out.stdout


def test_constants_match_header():
    from padertorch_b200 import _lib
    text = open(HEADER).read()
    defines = dict(re.findall(r'#define\s+(B2S_\w+)\s+(-?\d+)\b', text))
    expect = {'B2S_SPEC_INTERLEAVED': _lib.SPEC_INTERLEAVED, 'B2S_SPEC_CONCAT': _lib.SPEC_CONCAT,
              'B2S_SPEC_ABS': _lib.SPEC_ABS, 'B2S_SPEC_LOG1P_ABS': _lib.SPEC_LOG1P_ABS,
              'B2S_PIT_META': _lib.PIT_META, 'B2S_PAIR_META': _lib.PAIR_META, 'B2S_DC_META': _lib.DC_META,
              'B2S_MAX_SOURCES': _lib.MAX_SOURCES, 'B2S_DC_MAX_CHANNELS': _lib.DC_MAX_CHANNELS,
              'B2S_LOSS_MSE': _lib.LOSS_MSE, 'B2S_LOSS_LOG_MSE': _lib.LOSS_LOG_MSE,
              'B2S_LOSS_LOG1P_MSE': _lib.LOSS_LOG1P_MSE, 'B2S_LOSS_SDR': _lib.LOSS_SDR,
              'B2S_LOSS_SI_SDR': _lib.LOSS_SI_SDR, 'B2S_LOSS_SA_SDR': _lib.LOSS_SA_SDR,
              'B2S_FLAG_OFFSET_INVARIANT': _lib.FLAG_OFFSET_INVARIANT, 'B2S_FLAG_GRAD_STOP': _lib.FLAG_GRAD_STOP,
              'B2S_REDUCE_NONE': _lib.REDUCE_NONE, 'B2S_REDUCE_SUM': _lib.REDUCE_SUM,
              'B2S_REDUCE_MEAN': _lib.REDUCE_MEAN, 'B2S_VERSION': 100}
    for name, value in expect.items():
        assert int(defines[name]) == value, name


def test_frame_arithmetic_bit_exact(lib, golden):
    """Integer path against values recorded from the reference."""
    import padertorch_b200 as b2s
    from oracle import stft as OS
    for name, entry in golden.index['stft'].items():
        stft = b2s.ops.STFT(**entry['kwargs'])
        for samples, frames in entry['frames'].items():
            assert stft.samples_to_frames(int(samples)) == frames, (name, samples)
        for frames, samples in entry['frames_to_samples'].items():
            assert stft.frames_to_samples(int(frames)) == samples, (name, frames)
    # reference tests' known answers: tests/test_ops/test_stft.py:44-70, 139-165
    stft = b2s.ops.STFT(1024, 256)
    for fading, expect in ((False, (1, 1, 2)), (True, (7, 7, 8))):
        stft.fading = fading
        assert tuple(stft.samples_to_frames(n) for n in (1023, 1024, 1025)) == expect
    stft = b2s.ops.STFT(512, 20, window_length=40)
    for fading, expect in ((False, (50, 50, 51)), (True, (52, 52, 53))):
        stft.fading = fading
        assert tuple(stft.samples_to_frames(n) for n in (1019, 1020, 1021)) == expect
    # sweep against the oracle, including arrays, 'half' fading and pad=False
    rng = np.random.RandomState(0)
    for _ in range(200):
        size = int(rng.choice([64, 256, 400, 512, 1024]))
        shift = int(rng.randint(1, size))
        wl = int(rng.randint(max(shift, 2), size + 1))
        fading = [None, 'full', 'half', True, False][rng.randint(5)]
        pad = bool(rng.randint(2))
        stft = b2s.ops.STFT(size, shift, window_length=wl, fading=fading, pad=pad, window='hann')
        n = rng.randint(wl, 200000, size=5)
        want = OS.samples_to_frames(n, wl, shift, pad, fading)
        np.testing.assert_array_equal(stft.samples_to_frames(n), want)
        assert [stft.samples_to_frames(int(v)) for v in n] == list(want)
        m = rng.randint(1, 500, size=5)
        np.testing.assert_array_equal(stft.frames_to_samples(m), OS.frames_to_samples(m, wl, shift, fading))
        assert [stft.frames_to_samples(int(v)) for v in m] == list(OS.frames_to_samples(m, wl, shift, fading))
        np.testing.assert_array_equal(stft.sample_index_to_frame_index(n),
                                      OS.sample_index_to_frame_index(n, wl, shift, fading))
        # frames __call__ would produce == conv1d output length of the reference port
        for samples in (int(n[0]), wl, wl + 1):
            frames, _ = stft._frames_of_call(samples)
            ref = OS.ReferenceSTFT(size, shift, window_length=wl, fading=fading, pad=pad, window='hann')
            assert ref(torch.zeros(samples)).shape[0] == frames, (size, shift, wl, fading, pad, samples)


def test_windows_match_oracle():
    import padertorch_b200 as b2s
    from oracle import stft as OS
    from padertorch_b200.ops._stft import _biorthogonal_window
    for kwargs in (dict(size=1024, shift=256), dict(size=512, shift=20, window_length=40, window='hamming'),
                   dict(size=256, shift=64, symmetric_window=True, window='hann'),
                   dict(size=100, shift=25, window='hann'), dict(size=512, shift=48, window_length=96)):
        stft = b2s.ops.STFT(**kwargs)
        ref = OS.ReferenceSTFT(**kwargs)
        np.testing.assert_array_equal(stft.window, ref.window)
        np.testing.assert_array_equal(_biorthogonal_window(stft.window, stft.shift),
                                      OS.biorthogonal_window(ref.window, ref.shift))


def test_no_cpu_fallback():
    import padertorch_b200 as b2s
    x = torch.zeros(2, 3000)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        b2s.ops.STFT(1024, 256)(x)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        b2s.ops.pit_loss(torch.zeros(5, 2, 7), torch.zeros(5, 2, 7), axis=1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        b2s.ops.si_sdr_loss(torch.zeros(2, 70), torch.zeros(2, 70))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        b2s.ops.deep_clustering_loss(torch.zeros(50, 4), torch.zeros(50, 2))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        b2s.review.tasnet_losses(torch.zeros(2, 2, 100), torch.zeros(2, 2, 100), [100, 90])


def test_argument_validation_mirrors_reference():
    import padertorch_b200 as b2s
    with pytest.raises(AssertionError, match='even FFT sizes'):
        b2s.ops.STFT(1023, 256)
    with pytest.raises(AssertionError, match='predefined output_types'):
        b2s.ops.STFT(1024, 256, complex_representation='polar')
    with pytest.raises(AssertionError):
        b2s.ops.STFT(1024, 256, fading='quarter')
    with pytest.raises(AssertionError, match='Are you sure'):
        b2s.ops.pit_loss(torch.zeros(3, 30, 4), torch.zeros(3, 30, 4), axis=1)
    with pytest.raises(AssertionError):
        b2s.ops.pit_loss(torch.zeros(3, 2, 4), torch.zeros(3, 2, 5), axis=1)
    with pytest.raises(AssertionError, match='Uncommon value'):
        b2s.ops.sdr_loss(torch.zeros(2, 8), torch.zeros(2, 8), soft_sdr_max=80)
    with pytest.raises(AssertionError, match='Number of speakers'):
        b2s.ops.si_sdr_loss(torch.zeros(12, 8), torch.zeros(12, 8))


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    offenders = []
    package = os.path.join(ROOT, 'padertorch_b200')
    for folder, _, files in os.walk(package):
        for name in files:
            if not name.endswith('.py'):
                continue
            path = os.path.join(folder, name)
            tree = ast.parse(open(path).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module and node.level == 0:
                    names = [node.module]
                if any(n == 'oracle' or n.startswith('oracle.') for n in names):
                    offenders.append(path)
    assert not offenders, offenders
    for name in os.listdir(os.path.join(package, 'csrc')):
        if name.endswith(('.cu', '.cuh')):
            assert 'oracle' not in open(os.path.join(package, 'csrc', name)).read(), name
