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
SOURCES = ['plan.cu', 'gsa.cu', 'gsa_ws.cu', 'gsa_f64.cu', 'dense.cu', 'optim.cu', 'collate.cu', 'head.cu', 'cin_msg.cu']
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


def _compile(srcs, out, defines=(), verbose=False, force=False):
    """Each source to its own object (in parallel, skipped when the object is newer than the source, every header and
    the flags), then one link step. Objects live in csrc/_obj/ (git-ignored, gpurun-ignored)."""
    from concurrent.futures import ThreadPoolExecutor
    tag = '_'.join(sorted(d.replace('=', '-') for d in defines)) or 'default'
    objdir = os.path.join(CSRC, '_obj', tag)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers += [os.path.join(ROOT, 'include', 'cwn_b200.h'), os.path.abspath(__file__)]
    newest_header = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != '-shared'] + [f'-D{d}' for d in defines]

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > newest_header):
            return obj, ''
        cmd = [_nvcc()] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('cwn_b200: nvcc failed\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
        return obj, proc.stderr

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(one, srcs))
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC', '-o', out + '.tmp'] + \
          [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('cwn_b200: link failed\n' + proc.stdout + proc.stderr)
    os.replace(out + '.tmp', out)
    if verbose:
        print(''.join(log for _, log in results))
    return out


def build_variant(name, defines):
    """Profiling aid: an alternative build of the library (`libcwn_b200_<name>.so`) with extra -D flags; select it
    at run time with CWN_B200_LIB=<path>."""
    out = os.path.join(CSRC, f'libcwn_b200_{name}.so')
    return _compile([os.path.join(CSRC, s) for s in SOURCES], out, defines=tuple(defines))


def build_library(force=False, verbose=False):
    if os.environ.get('CWN_B200_LIB'):
        return os.environ['CWN_B200_LIB']
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return _compile(srcs, LIB, verbose=verbose, force=force)


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
