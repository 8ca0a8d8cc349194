"""Reader for Nek5000 ``.f%05d`` field files (the reference's fixtures).  TEST INFRASTRUCTURE ONLY.

Format ([UPSTREAM-RECALL] prepost.f mfo_*; layout confirmed on
``examples/cylinder/BF_1cyl0.f00001``): 132-byte ASCII header
``#std wdsize nx ny nz nelo nelg time istep fid nfiles rdcode``, a float32 endian tag
6.54321, ``nelo`` int32 global element ids, then per field group (X, U, P, T, S..)
element-by-element blocks of ``ndim`` (X, U) or 1 (P, T) components.
"""
from __future__ import annotations

import numpy as np


def read_fld(path):
    with open(path, 'rb') as f:
        raw = f.read()
    hdr = raw[:132].decode('ascii')
    tok = hdr.split()
    if tok[0] != '#std':
        raise ValueError(f'not a Nek field file: {hdr[:16]!r}')
    wd, nx, ny, nz, nelo, nelg = (int(t) for t in tok[1:7])
    time, istep = float(tok[7]), int(tok[8])
    rdcode = tok[11] if len(tok) > 11 else ''
    tag_le = np.frombuffer(raw, dtype='<f4', count=1, offset=132)[0]
    end = '<' if abs(tag_le - 6.54321) < 1e-5 else '>'
    ft = np.dtype(end + ('f8' if wd == 8 else 'f4'))
    ndim = 3 if nz > 1 else 2
    npt = nx * ny * nz
    off = 136
    elmap = np.frombuffer(raw, dtype=end + 'i4', count=nelo, offset=off).astype(np.int64)
    off += 4 * nelo
    shape = (nelo, nz, ny, nx) if ndim == 3 else (nelo, ny, nx)
    out = dict(wdsize=wd, nx=nx, ny=ny, nz=nz, nel=nelo, nelg=nelg, time=time, istep=istep,
               rdcode=rdcode, elmap=elmap, ndim=ndim)

    def take(ncomp):
        nonlocal off
        a = np.frombuffer(raw, dtype=ft, count=nelo * ncomp * npt, offset=off)
        off += a.nbytes
        a = a.reshape(nelo, ncomp, npt).astype(np.float64)
        return [a[:, c, :].reshape(shape) for c in range(ncomp)]

    i = 0
    while i < len(rdcode):
        c = rdcode[i]
        if c == 'X':
            out['x'] = take(ndim)
        elif c == 'U':
            out['u'] = take(ndim)
        elif c == 'P':
            out['p'] = take(1)[0]
        elif c == 'T':
            out['t'] = take(1)[0]
        elif c == 'S':
            ns = int(rdcode[i + 1:i + 3])
            out['s'] = [take(1)[0] for _ in range(ns)]
            i += 2
        i += 1
    return out
