"""The packed warp FFT (padertorch_b200/csrc/rfft_packed.cuh) is __host__ __device__: compile its
host emulation (tests/host/rfft_emulate.cpp runs the 32 lanes of a warp pass by pass on the CPU, shared
memory = a plain array) with g++ and compare all 513 bins with a double-precision DFT.  This pins the
lane <-> butterfly maps, exchange layouts, twiddles, lane-0 re-pairing and real split without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_include():
    for cand in (os.environ.get('CUDA_HOME'), '/usr/local/cuda'):
        if cand and os.path.exists(os.path.join(cand, 'include', 'cuda_runtime.h')):
            return os.path.join(cand, 'include')
    return None


@pytest.mark.skipif(shutil.which('g++') is None or _cuda_include() is None, reason='needs g++ and the CUDA headers')
def test_packed_rfft_host_emulation(tmp_path):
    exe = tmp_path / 'rfft_emulate'
    subprocess.run(['g++', '-O1', '-std=c++17', '-I', _cuda_include(), '-o', str(exe),
                    os.path.join(ROOT, 'tests', 'host', 'rfft_emulate.cpp')], check=True)
    out = subprocess.run([str(exe), '8'], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    err = float(out.stdout.split()[-1])
    assert err < 1e-6, out.stdout


@pytest.mark.skipif(shutil.which('g++') is None or _cuda_include() is None, reason='needs g++ and the CUDA headers')
def test_pair_transform_prototype_host_emulation(tmp_path):
    """csrc/cfft_pair.cuh (two real frames per 1024-point complex FFT, one exchange): both spectra
    against a double-precision DFT, lane by lane on the CPU."""
    exe = tmp_path / 'cfft_pair_emulate'
    subprocess.run(['g++', '-O1', '-std=c++17', '-I', _cuda_include(), '-o', str(exe),
                    os.path.join(ROOT, 'tests', 'host', 'cfft_pair_emulate.cpp')], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert float(out.stdout.split()[-1]) < 1e-6, out.stdout
