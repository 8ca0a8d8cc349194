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


# ------------------------------------------------------------------------------------------------
# The reference's mesh files: <case>.re2 (element corners, curved sides, boundary conditions) -- where Nek takes
# the velocity mask v1mask of nsb_sem_create from.  [UPSTREAM-RECALL: Nek5000 reader_re2.f (bin_rd1_mesh /
# bin_rd1_curve / bin_rd1_bc); the layout below is pinned byte for byte by the reference's own files
# examples/cylinder/1cyl.re2 (#v002) and examples/back_fstep/baseflow/bfs.re2 (#v003), tests/test_mesh_files.py.]
# ------------------------------------------------------------------------------------------------
def read_re2(path) -> dict:
    """80-byte ASCII header ``#v00x nelgt ndim nelgv``, a float32 endian tag 6.54321, then 8-byte reals throughout
    (#v002 / #v003; #v001 holds 4-byte reals): per element (group, x(2^d), y(2^d)[, z(2^d)]) with the corners in
    preprocessor order (2-D: counter-clockwise from (r, s) = (-1, -1); 3-D: the bottom face then the top face);
    ncurve and 8 reals per curved side (element, side, 5 parameters, type); then, for every field with boundary
    conditions, nbc and 8 reals per face (element, side, 5 parameters, 3-character type).  #v003 stores the type
    'MSH' and the boundary id in the fifth parameter (the .par / .usr file maps ids to types).
    Returns dict(version, ndim, nel, nelv, group, xc, yc[, zc], curves=[(e, side, params, type)],
    bcs=[one list per field of (e, side, params, type)]); elements and sides are 1-based as in the file."""
    raw = open(path, 'rb').read()
    hdr = raw[:80].decode('ascii')
    version = hdr[:5]
    if version not in ('#v001', '#v002', '#v003'):
        raise ValueError(f'{path}: not a .re2 file ({hdr[:16]!r})')
    nel, ndim, nelv = int(hdr[5:14]), int(hdr[14:17]), int(hdr[17:27])
    tag = np.frombuffer(raw, dtype='<f4', count=1, offset=80)[0]
    end = '<' if abs(tag - 6.54321) < 1e-5 else '>'
    if abs(np.frombuffer(raw, dtype=end + 'f4', count=1, offset=80)[0] - 6.54321) > 1e-5:
        raise ValueError(f'{path}: endian tag not found')
    wd = 4 if version == '#v001' else 8
    ft = np.dtype(end + ('f4' if wd == 4 else 'f8'))
    off = 84
    nc = 2 ** ndim
    nw = 1 + ndim * nc

    def take(count):
        nonlocal off
        if off + count * wd > len(raw):
            raise ValueError(f'{path}: truncated')
        a = np.frombuffer(raw, dtype=ft, count=count, offset=off).astype(np.float64)
        off += count * wd
        return a

    def text(rec):                      # the character field sits in the bytes of the last real of a record
        return rec.astype(ft).tobytes()[7 * wd:7 * wd + 3].decode('ascii')

    el = take(nel * nw).reshape(nel, nw)
    out = dict(version=version, ndim=ndim, nel=nel, nelv=nelv, group=el[:, 0].astype(np.int64))
    if ndim == 2:
        out['xc'], out['yc'] = el[:, 1:5].copy(), el[:, 5:9].copy()
    else:                               # x(1:4) y(1:4) z(1:4) of the bottom face, then the same for the top face
        out['xc'] = np.concatenate([el[:, 1:5], el[:, 13:17]], axis=1)
        out['yc'] = np.concatenate([el[:, 5:9], el[:, 17:21]], axis=1)
        out['zc'] = np.concatenate([el[:, 9:13], el[:, 21:25]], axis=1)

    def records():
        n = int(take(1)[0])
        recs = take(n * 8).reshape(n, 8)
        return [(int(r[0]), int(r[1]), r[2:7].copy(), text(r)) for r in recs]

    out['curves'] = records()
    out['bcs'] = []
    while off < len(raw):
        out['bcs'].append(records())
    return out


def face_nodes(lx: int, ndim: int, side: int):
    """Index tuple (without the element axis) of the GLL points of preprocessor side ``side`` (1-based) of an
    element stored (.., k, j, i): 1: s = -1, 2: r = +1, 3: s = +1, 4: r = -1, 5: t = -1, 6: t = +1."""
    full = slice(None)
    last = lx - 1
    where = {1: ('j', 0), 2: ('i', last), 3: ('j', last), 4: ('i', 0), 5: ('k', 0), 6: ('k', last)}
    if side not in where or (ndim == 2 and side > 4):
        raise ValueError(f'side {side} of a {ndim}-D element')
    axis, pos = where[side]
    idx = {'i': full, 'j': full, 'k': full}
    idx[axis] = pos
    return (idx['j'], idx['i']) if ndim == 2 else (idx['k'], idx['j'], idx['i'])


def dirichlet_mask(re2: dict, lx: int, types=('W  ', 'v  ', 'V  '), ids=(), field: int = 0, elements=None,
                   glo=None):
    """v1mask of the perturbation velocity from the boundary conditions of a .re2 file: 0 on every point of a face
    whose type is in ``types`` (walls and prescribed-velocity faces: the perturbation vanishes there; outflow 'O',
    periodic 'P' and interior faces keep 1) or, for #v003 files, whose boundary id is in ``ids``.  ``elements``:
    1-based global ids of the local elements in storage order (Nek's LGLEL; default: all, in file order).
    ``glo``: global node numbering of the local points; the zeros then reach every copy of a node (an element that
    touches a wall with a corner only), as Nek's multiplicative dsop over the mask does -- on one rank; across ranks
    the caller combines the masks the same way.
    Shape (nel_local, lx, lx[, lx]).  A mask that differs between velocity components (SYM) is not expressed."""
    ndim = re2['ndim']
    elements = np.arange(1, re2['nel'] + 1) if elements is None else np.asarray(elements, dtype=np.int64)
    local = {int(g): l for l, g in enumerate(elements)}
    mask = np.ones((len(elements),) + (lx,) * ndim)
    for e, side, params, typ in re2['bcs'][field]:
        hit = typ in types or (typ == 'MSH' and int(params[4]) in ids)
        if hit and e in local:
            mask[(local[e],) + face_nodes(lx, ndim, side)] = 0.0
    if glo is not None:
        g = np.asarray(glo).reshape(-1)
        lowest = np.ones(int(g.max()) + 1)
        np.minimum.at(lowest, g, mask.reshape(-1))
        mask = lowest[g].reshape(mask.shape)
    return mask


def periodic_glo_num(glo, re2: dict, coords, lx: int, field: int = 0, elements=None):
    """Global numbering with the partners of periodic faces ('P  ': parameters 1 and 2 = partner element and side)
    identified, as Nek's numbering has them: the points of a face are matched with those of its partner by their
    position along the face (the coordinates tangential to it), the two ids are merged, and the merged classes are
    renumbered 0 .. n-1.  coords = (x, y[, z]) of the local points; elements as in dirichlet_mask.  Partners that are
    not local are left alone (a multi-rank caller merges on the global numbering before partitioning)."""
    ndim = re2['ndim']
    glo = np.asarray(glo)
    elements = np.arange(1, re2['nel'] + 1) if elements is None else np.asarray(elements, dtype=np.int64)
    local = {int(g): l for l, g in enumerate(elements)}
    parent = np.arange(int(glo.max()) + 1)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for e, side, params, typ in re2['bcs'][field]:
        pe, ps = int(params[0]), int(params[1])
        if typ != 'P  ' or e not in local or pe not in local:
            continue
        mine = (local[e],) + face_nodes(lx, ndim, side)
        theirs = (local[pe],) + face_nodes(lx, ndim, ps)
        # tangential coordinates: those that agree between the two faces (the normal one differs by the period)
        keys = []
        for c in coords:
            a, b = np.asarray(c)[mine].ravel(), np.asarray(c)[theirs].ravel()
            if abs(np.sort(a) - np.sort(b)).max() <= 1e-6 * max(1.0, abs(a).max()):
                keys.append((a, b))
        if not keys:
            raise ValueError(f'periodic faces ({e}, {side}) and ({pe}, {ps}) share no tangential coordinate')
        oa = np.lexsort([np.round(k[0], 6) for k in keys])
        ob = np.lexsort([np.round(k[1], 6) for k in keys])
        for a, b in zip(glo[mine].ravel()[oa], glo[theirs].ravel()[ob]):
            ra, rb = find(int(a)), find(int(b))
            if ra != rb:
                parent[rb] = ra
    roots = np.array([find(a) for a in range(parent.size)])
    _, new = np.unique(roots, return_inverse=True)
    return new[glo].astype(np.int64)
