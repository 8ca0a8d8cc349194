"""In-tree build of libnekstab_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / 'csrc'
LIB = PKG / 'libnekstab_b200.so'
SOURCES = ['nsb_core.cu', 'nsb_orth.cu', 'nsb_sem.cu', 'nsb_conv.cu', 'nsb_comm.cu', 'nsb_krylov.cu']
HEADERS = [CSRC / 'nsb_internal.h', CSRC / 'nsb_device.cuh', ROOT / 'include' / 'nekstab_b200.h']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '-diag-suppress', '128']


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libnekstab_b200.so cannot be built')
    return exe


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source into one shared library next to the package."""
    if not force and not stale():
        return LIB
    objdir = PKG / 'build'
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    inc = ['-I', str(ROOT / 'include'), '-I', str(CSRC)]
    procs = []
    for s in SOURCES:
        obj = objdir / (s + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *inc, '-c', str(CSRC / s), '-o', str(obj)]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {s}:\n{out}')
        if verbose and out:
            print(out)
        objs.append(str(obj))
    tmp = str(LIB) + '.tmp'
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', tmp, *objs, '-ldl']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
