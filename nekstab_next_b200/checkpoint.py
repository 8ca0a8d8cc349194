"""Krylov checkpoint / restart wire formats of the reference (host side).

* ``HES<session>%04d``  -- Hessenberg matrix as list-directed text, row by row
  ``write(67,*) ((H(i,j), j=1,k), i=1,k+1)`` (core/eigensolvers.f90:837-843); the reader accepts any
  whitespace-separated layout like Fortran's ``read(67,*)`` (core/eigensolvers.f90:246-266).
* ``Spectre_H*.dat`` / ``Spectrum_*.dat`` -- three columns ``Re Im residual`` in ``3E15.7``
  (core/eigensolvers.f90:518-525, 815-829; core/linear_stab.f90:295-312).
* ``KRY<session>0.f%05d`` -- Krylov vectors as Nek5000 field files (``outpost2`` in
  core/eigensolvers.f90:803-809, read back by ``load_files``, core/IO.f90:11-72):
  132-byte ASCII header ``#std wdsize nx ny nz nelo nelg time istep fid nfiles rdcode``, float32
  endian tag 6.54321, int32 global element ids, then per field group element-by-element blocks.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence

import numpy as np


# ---- Fortran E15.7 ------------------------------------------------------------------------------
def fortran_e(v: float, width: int = 15, digits: int = 7) -> str:
    """``E15.7`` as gfortran prints it: 0.dddddddE+ee, right-justified."""
    if not np.isfinite(v):
        return ('NaN' if np.isnan(v) else ('Infinity' if v > 0 else '-Infinity')).rjust(width)
    if v == 0.0:
        s = '0.' + '0' * digits + 'E+00'
    else:
        e = int(np.floor(np.log10(abs(v)))) + 1
        m = abs(v) / 10.0 ** e
        ms = f'{m:.{digits}f}'
        if ms.startswith('1'):          # rounding pushed the mantissa to 1.0000000
            e += 1
            ms = f'{abs(v) / 10.0 ** e:.{digits}f}'
        s = ('-' if v < 0 else '') + ms + f'E{e:+03d}'
    return s.rjust(width)


def write_spectrum(path, vals: np.ndarray, residual: np.ndarray):
    """Three columns Re, Im, residual in (3E15.7), one eigenvalue per line."""
    with open(path, 'w') as f:
        for v, r in zip(np.asarray(vals), np.asarray(residual)):
            f.write(fortran_e(v.real) + fortran_e(v.imag) + fortran_e(float(r)) + '\n')


def log_transform(x):
    """core/eigensolvers.f90:860-869: log(x); a real argument gives a real result (the imaginary part of
    log of a negative real, pi, is dropped exactly as the reference does)."""
    x = np.asarray(x, dtype=np.complex128)
    y = np.log(x)
    return np.where(x.imag == 0, y.real + 0j, y)


def write_ns_spectrum(path, vals: np.ndarray, residual: np.ndarray, speriod: float):
    """The log-transformed spectrum of the linearised operator (unit 20 of outpost_ks,
    core/eigensolvers.f90:547-548; core/linear_stab.f90:72): log(lambda) / T per Ritz value."""
    write_spectrum(path, log_transform(vals) / speriod, residual)


def write_singvals(path, sigma: np.ndarray, residual: np.ndarray):
    """outpost_singvals (core/linear_stab.f90:313-329): two columns sigma, residual in (2E15.7)."""
    with open(path, 'w') as f:
        for v, r in zip(np.asarray(sigma, dtype=float), np.asarray(residual, dtype=float)):
            f.write(fortran_e(float(v)) + fortran_e(float(r)) + '\n')


def read_singvals(path):
    a = np.loadtxt(path, ndmin=2)
    return a[:, 0], a[:, 1]


def read_spectrum(path):
    a = np.loadtxt(path, ndmin=2)
    return a[:, 0] + 1j * a[:, 1], a[:, 2]


# ---- Hessenberg matrix ------------------------------------------------------------------------------
def hessenberg_name(session: str, k: int) -> str:
    return f'HES{session}{k:04d}'


def write_hessenberg(path, H: np.ndarray, k: int):
    """H(1:k+1, 1:k) row by row, list-directed (what ``write(67,*)`` produces is whitespace-separated
    reals; full double precision is kept so a restart reproduces the factorisation)."""
    with open(path, 'w') as f:
        vals = [repr(float(H[i, j])) for i in range(k + 1) for j in range(k)]
        for a in range(0, len(vals), 3):
            f.write('  ' + '  '.join(vals[a:a + 3]) + '\n')


def read_hessenberg(path, k_dim: int, mstart: int) -> np.ndarray:
    """Restart read of core/eigensolvers.f90:246-266: the file holds (mstart+1) x mstart values row by
    row; returns the (k_dim+1) x k_dim array with the leading block filled (subsampled to k_dim
    columns when k_dim < mstart, like the reference)."""
    vals = np.array(Path(path).read_text().replace('D', 'E').split(), dtype=np.float64)
    if vals.size != (mstart + 1) * mstart:
        raise ValueError(f'{path}: expected {(mstart + 1) * mstart} values for mstart={mstart}, found {vals.size}')
    blk = vals.reshape(mstart + 1, mstart)
    H = np.zeros((k_dim + 1, k_dim), order='F')
    r, c = min(mstart + 1, k_dim + 1), min(mstart, k_dim)
    H[:r, :c] = blk[:r, :c]
    return H


# ---- Nek field files ----------------------------------------------------------------------------------
def field_name(prefix: str, session: str, index: int) -> str:
    return f'{prefix}{session}0.f{index:05d}'


def write_fld(path, fields: dict, nx: int, ny: int, nz: int, time: float = 0.0, istep: int = 0,
              elmap: Optional[Sequence[int]] = None, wdsize: int = 8):
    """fields: ordered dict with optional keys 'x' (list of ndim arrays), 'u' (list of ndim arrays),
    'p' (array), 't' (array); every array has shape (nel, nz, ny, nx) / (nel, ny, nx)."""
    ndim = 3 if nz > 1 else 2
    first = (fields.get('u') or fields.get('x') or [fields.get('p')])[0]
    nel = first.shape[0]
    rd = ''.join(c for c, k in (('X', 'x'), ('U', 'u'), ('P', 'p'), ('T', 't')) if fields.get(k) is not None)
    # [UPSTREAM-RECALL] prepost.f mfo_write_hdr: '#std',1x,i1,1x,i2,1x,i2,1x,i2,1x,i10,1x,i10,1x,e20.13,
    # 1x,i9,1x,i6,1x,i6,1x,10a,1x,1pe15.7 (p0th),1x,l1 (if_press_mesh); checked on the reference's fixtures
    hdr = '#std %1d %2d %2d %2d %10d %10d %s %9d %6d %6d %-10s %14.7E %s' % (
        wdsize, nx, ny, nz, nel, nel, fortran_e(time, 20, 13), istep, 0, 1, rd, 1.0, 'F')
    dt = np.dtype('<f8' if wdsize == 8 else '<f4')
    with open(path, 'wb') as f:
        f.write(hdr.ljust(132).encode('ascii'))
        f.write(np.array([6.54321], dtype='<f4').tobytes())
        em = np.arange(1, nel + 1, dtype='<i4') if elmap is None else np.asarray(elmap, dtype='<i4')
        f.write(em.tobytes())
        npt = nx * ny * nz
        for key in ('x', 'u'):
            if fields.get(key) is not None:
                comps = [np.asarray(c, dtype=np.float64).reshape(nel, npt) for c in fields[key]]
                assert len(comps) == ndim
                f.write(np.stack(comps, axis=1).astype(dt).tobytes())       # (nel, ndim, npt)
        for key in ('p', 't'):
            if fields.get(key) is not None:
                f.write(np.asarray(fields[key], dtype=np.float64).reshape(nel, npt).astype(dt).tobytes())


def read_fld(path) -> dict:
    raw = Path(path).read_bytes()
    tok = raw[:132].decode('ascii').split()
    if tok[0] != '#std':
        raise ValueError(f'{path}: not a Nek field file')
    wd, nx, ny, nz, nelo, nelg = (int(t) for t in tok[1:7])
    rd = tok[11] if len(tok) > 11 else ''
    end = '<' if abs(np.frombuffer(raw, '<f4', 1, 132)[0] - 6.54321) < 1e-5 else '>'
    ft = np.dtype(end + ('f8' if wd == 8 else 'f4'))
    ndim = 3 if nz > 1 else 2
    npt = nx * ny * nz
    shape = (nelo, nz, ny, nx) if ndim == 3 else (nelo, ny, nx)
    off = 136
    out = dict(wdsize=wd, nx=nx, ny=ny, nz=nz, nel=nelo, nelg=nelg, time=float(tok[7]), istep=int(tok[8]),
               rdcode=rd, ndim=ndim, elmap=np.frombuffer(raw, end + 'i4', nelo, off).astype(np.int64))
    off += 4 * nelo

    def take(nc):
        nonlocal off
        a = np.frombuffer(raw, ft, nelo * nc * npt, off).reshape(nelo, nc, npt).astype(np.float64)
        off += nelo * nc * npt * ft.itemsize
        return [a[:, c].reshape(shape) for c in range(nc)]

    for c in rd:
        if c == 'X':
            out['x'] = take(ndim)
        elif c == 'U':
            out['u'] = take(ndim)
        elif c == 'P':
            out['p'] = take(1)[0]
        elif c == 'T':
            out['t'] = take(1)[0]
    return out
