"""In-tree build of the C-ABI library (`cwn_b200/csrc/libcwn_b200.so`) with nvcc for sm_100a.

The `.so` is git-ignored (history stays source-only) but travels to the GPU box with the gpurun snapshot.
`python -m cwn_b200.build` or `__graft_entry__.build()` runs it; it is skipped when the library is newer than
every source file."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, 'libcwn_b200.so')
SOURCES = ['plan.cu', 'gsa.cu', 'dense.cu', 'optim.cu', 'collate.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('cwn_b200: nvcc not found; the CUDA library cannot be built')
    return nvcc


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(ROOT, 'include', 'cwn_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Profiling aid: an alternative build of the library (`libcwn_b200_<name>.so`) with extra -D flags; select it
    at run time with CWN_B200_LIB=<path>."""
    out = os.path.join(CSRC, f'libcwn_b200_{name}.so')
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [_nvcc()] + NVCC_FLAGS + [f'-D{d}' for d in defines] + ['-o', out] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('cwn_b200: nvcc failed\n' + proc.stdout + proc.stderr)
    return out


def build_library(force=False, verbose=False):
    if os.environ.get('CWN_B200_LIB'):
        return os.environ['CWN_B200_LIB']
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB + '.tmp'] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('cwn_b200: nvcc failed\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    os.replace(LIB + '.tmp', LIB)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
