"""In-tree nvcc build of libsessrec_b200.so (sm_100a only).  No GPU is needed to build."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
INCLUDE = PKG.parent / 'include'
LIB = PKG / 'libsessrec_b200.so'
OBJ = PKG / 'build'

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default', '--expt-relaxed-constexpr',
              '-I', str(INCLUDE)] + os.environ.get('SESSREC_NVCC_EXTRA', '').split()


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and Path(c).exists():
            return c
    raise RuntimeError('nvcc not found; cannot build libsessrec_b200.so')


def sources():
    return sorted(CSRC.glob('*.cu'))


def _stale(target, deps):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link the shared library; returns its path."""
    srcs = sources()
    hdrs = list(CSRC.glob('*.cuh')) + list(INCLUDE.glob('*.h'))
    if not force and not _stale(LIB, srcs + hdrs + [Path(__file__)]):
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)

    def one(src):
        obj = OBJ / (src.stem + '.o')
        if force or _stale(obj, [src] + hdrs + [Path(__file__)]):
            cmd = [nvcc, *NVCC_FLAGS, '-Xptxas', '-v' if verbose else '-warn-spills', '-c', str(src), '-o', str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f'nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}')
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(one, srcs))
    tmp = LIB.with_suffix('.so.tmp')
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
