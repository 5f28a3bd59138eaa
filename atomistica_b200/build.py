"""In-tree build of libatomistica_b200.so (hand-written CUDA, sm_100a only).

    python -m atomistica_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so lives next to this file so that it
travels with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libatomistica_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '-I' + os.path.join(HERE, '..', 'include'), '-I/usr/include']
VISIBLE = ['-Xcompiler', '-fvisibility=default']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu') or f.endswith('.cpp'))


def _newest_dep():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, '..', 'include', 'atomistica_b200.h'))
    deps.append(os.path.abspath(__file__))
    return max(os.path.getmtime(d) for d in deps)


def _newest_header():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh') or f.endswith('.h')]
    deps.append(os.path.join(HERE, '..', 'include', 'atomistica_b200.h'))
    deps.append(os.path.abspath(__file__))
    return max(os.path.getmtime(d) for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + '.o')
    # a translation unit depends on itself and on every header, not on the other .cu files
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(_newest_header(),
                                                            os.path.getmtime(os.path.join(CSRC, src))):
        return obj
    cmd = [NVCC] + NVCC_FLAGS + VISIBLE + ['-x', 'cu', '-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_dep():
        return LIB
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, sources()))
    cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + \
        ['-lcudart_static', '-ldl', '-lpthread', '-lrt']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
