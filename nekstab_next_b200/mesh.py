"""Synthetic box meshes for the benchmark configurations.

Host-side setup only (numpy); the GLL nodes come from the library's own ``nsb_gll``.  Elements are
ordered x fastest, z slowest, so a contiguous range of elements is a slab of z-layers -- the
element partition Nek would hand to consecutive MPI ranks.
"""
from __future__ import annotations

import numpy as np

from .api import gll


def partition_range(nel_total: int, rank: int, nranks: int, granule: int = 1):
    """Contiguous element range of ``rank`` (whole ``granule``s, e.g. z-layers)."""
    ngran = nel_total // granule
    base, rem = divmod(ngran, nranks)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo * granule, hi * granule


def box_mesh(nelx: int, nely: int, nelz: int, N: int, deform: float = 0.0, rank: int = 0,
             nranks: int = 1, lengths=(1.0, 1.0, 1.0)):
    """Coordinates, lexicographic global node ids and Dirichlet mask of this rank's slab.

    Returns dict(x, y, z, glo, mask, nel, nel_total, e0) with arrays of shape (nel, lx, lx, lx).
    """
    lx = N + 1
    zg, _, _ = gll(N)
    r = 0.5 * (zg + 1.0)
    nel_total = nelx * nely * nelz
    e0, e1 = partition_range(nel_total, rank, nranks, granule=nelx * nely)
    k0, k1 = e0 // (nelx * nely), e1 // (nelx * nely)
    nzl = k1 - k0

    def line(i0, i1, nel, length):
        return (np.arange(i0, i1)[:, None] + r[None, :]) * (length / nel)

    xl, yl, zl = line(0, nelx, nelx, lengths[0]), line(0, nely, nely, lengths[1]), line(k0, k1, nelz, lengths[2])
    shp = (nzl, nely, nelx, lx, lx, lx)
    x = np.empty(shp)
    y = np.empty(shp)
    z = np.empty(shp)
    x[:] = xl[None, None, :, None, None, :]
    y[:] = yl[None, :, None, None, :, None]
    z[:] = zl[:, None, None, :, None, None]
    gi = np.arange(nelx)[:, None] * N + np.arange(lx)[None, :]
    gj = np.arange(nely)[:, None] * N + np.arange(lx)[None, :]
    gk = np.arange(k0, k1)[:, None] * N + np.arange(lx)[None, :]
    nxg, nyg, nzg = nelx * N + 1, nely * N + 1, nelz * N + 1
    glo = np.empty(shp, dtype=np.int64)
    glo[:] = (gk[:, None, None, :, None, None] * nyg + gj[None, :, None, None, :, None]) * nxg \
        + gi[None, None, :, None, None, :]
    mask = np.ones(shp)
    bi = (gi == 0) | (gi == nxg - 1)
    bj = (gj == 0) | (gj == nyg - 1)
    bk = (gk == 0) | (gk == nzg - 1)
    mask[np.broadcast_to(bi[None, None, :, None, None, :], shp)] = 0.0
    mask[np.broadcast_to(bj[None, :, None, None, :, None], shp)] = 0.0
    mask[np.broadcast_to(bk[:, None, None, :, None, None], shp)] = 0.0
    nel = nzl * nely * nelx
    x, y, z = (a.reshape(nel, lx, lx, lx) for a in (x, y, z))
    glo, mask = glo.reshape(nel, lx, lx, lx), mask.reshape(nel, lx, lx, lx)
    if deform != 0.0:
        bump = deform * np.sin(np.pi * x / lengths[0]) * np.sin(np.pi * y / lengths[1]) \
            * np.sin(np.pi * z / lengths[2])
        x, y, z = x + bump, y + bump, z + bump
    return dict(x=x, y=y, z=z, glo=glo, mask=mask, nel=nel, nel_total=nel_total, e0=e0)


def taylor_green(x, y, z):
    """Analytic base flow of SURVEY.md section 8d."""
    tp = 2.0 * np.pi
    return (np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z),
            -np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z),
            np.zeros_like(x))
