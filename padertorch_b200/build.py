"""Build libb200sep.so (the C-ABI CUDA library, include/b200sep.h) in-tree with nvcc for sm_100a.

    python -m padertorch_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libb200sep.so')
SOURCES = ['capi.cu', 'stft.cu', 'pit.cu', 'pairstats.cu', 'dc.cu', 'fused.cu', 'fused_bwd.cu', 'targets.cu', 'targets_fused.cu', 'gemm_umma.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']
NVCC_FLAGS += os.environ.get('B2S_NVCC_EXTRA', '').split()   # A/B builds of compile-time alternatives (tools/)
if os.environ.get('B2S_TUNING', '0') not in ('', '0'):      # per-warp trace stamps / ablation bits of the fused kernel
    NVCC_FLAGS.append('-DB2S_TUNING=1')


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


STAMP = os.path.join(OBJ, 'source.sha256')


def source_hash():
    """sha256 over every input of the build (sources, headers, the C ABI header, this file, the flags)."""
    import hashlib
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh')))
    paths.append(os.path.join(HERE, '..', 'include', 'b200sep.h'))
    paths.append(os.path.abspath(__file__))
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def is_current():
    """True when the library on disk was built from exactly the sources on disk (content hash, not mtime: a
    checkout or a copy to another box changes every mtime but not the contents)."""
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_hash()


def build(force=False, verbose=False):
    """Compile every kernel source and link the shared library.  Returns the library path."""
    if not force and is_current():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(name):
        obj = os.path.join(OBJ, name.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, name), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f'nvcc failed for {name}:\n{proc.stdout}\n{proc.stderr}')
        if verbose:
            sys.stderr.write(proc.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objects = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, '-shared', '-o', LIB, *objects, '-gencode', 'arch=compute_100a,code=sm_100a']
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'link failed:\n{proc.stdout}\n{proc.stderr}')
    with open(STAMP, 'w') as f:
        f.write(source_hash() + '\n')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
