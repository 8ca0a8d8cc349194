"""Shared builders for the parity tests: the same seeded inputs go to the oracle and the GPU."""
from __future__ import annotations

import numpy as np

from oracle import sem as osem
from oracle import krylov as okr


class BoxProblem:
    """3-D (or 2-D) box mesh with the oracle's geometry and the synthetic operator
    M = alpha I + beta B^-1 mask QQ^T (h1 A + h2 B [+ B c.grad])  (SURVEY.md section 8d)."""

    def __init__(self, nel=(3, 3, 3), N=4, deform=0.05, nfields=1, alpha=1.0, beta=None, h1=1.0,
                 h2=0.1, conv=False, pressure=False, time_in_dot=False, seed=0):
        self.N, self.nfields = N, nfields
        self.dim = len(nel)
        if self.dim == 3:
            x, y, z, glo = osem.box_mesh(*nel, N, deform=deform)
            x0, y0, z0, _ = osem.box_mesh(*nel, N)
            self.coords = (x, y, z)
            self.mask = osem.boundary_mask_box(None, x0, y0, z0)
        else:
            x, y, glo = osem.box_mesh_2d(*nel, N, deform=deform)
            x0, y0, _ = osem.box_mesh_2d(*nel, N)
            self.coords = (x, y)
            self.mask = osem.boundary_mask_box(None, x0, y0, None, lengths=(1.0, 1.0))
        self.glo = glo
        self.geo = osem.geometry(N, *self.coords)
        self.d = osem.dgll(N)
        self.bm1 = self.geo['bm1']
        self.binv = 1.0 / osem.dssum(self.bm1, glo)
        self.vmult = 1.0 / osem.multiplicity(glo)
        self.h1, self.h2, self.alpha = h1, h2, alpha
        self.conv = None
        if conv:
            if self.dim == 3:
                tp = 2 * np.pi
                self.conv = (np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z),
                             -np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z), 0.3 + 0 * x)
            else:
                self.conv = (1.0 + 0 * x, 0.5 + y * 0)
        self.shape = x.shape
        self.npts = x.size
        self.pressure = pressure
        self.np_pr = self.npts // 2 if pressure else 0
        self.time_in_dot = time_in_dot
        self.rng = np.random.default_rng(seed)
        if beta is None:
            # scale so the spectrum of M sits inside the unit disc: beta = -1/lambda_max estimate
            u = self.random_field()
            lam = 0.0
            for _ in range(30):
                v = self.l_apply(u)
                lam = np.sqrt(osem.glsc3(v, v, self.bm1) / osem.glsc3(u, u, self.bm1))
                u = v / lam
            beta = -1.0 / (1.05 * lam)
        self.beta = beta

    # ---- oracle-side operator -----------------------------------------------------------------
    def l_apply(self, u):
        w = osem.axhelm(u, self.geo['g'], self.d, self.h1, self.h2, self.bm1)
        if self.conv is not None:
            _, wq = osem.gll(self.N)
            w3 = wq[:, None, None] * wq[None, :, None] * wq[None, None, :] if self.dim == 3 \
                else wq[:, None] * wq[None, :]
            w = w + w3[None] * osem.convect(u, self.conv, self.geo['rst'], self.d)
        return osem.dssum(w, self.glo) * self.mask * self.binv

    def m_apply_field(self, u):
        return self.alpha * u + self.beta * self.l_apply(u)

    def random_field(self):
        u = self.rng.standard_normal(self.shape)
        return osem.dssum(u, self.glo) * self.vmult * self.mask

    def random_kvec(self):
        f = [self.random_field() for _ in range(self.nfields)]
        if self.pressure:
            f.append(self.rng.standard_normal(self.np_pr))
        return okr.KVec(f, float(self.rng.standard_normal()) if self.time_in_dot else 0.0)

    def octx(self):
        in_dot = [True] * self.nfields + ([False] if self.pressure else [])
        return okr.Ctx(bm1s=self.bm1, in_dot=in_dot, time_in_dot=self.time_in_dot)

    def omatvec(self, q):
        f = [self.m_apply_field(q.f[c]) for c in range(self.nfields)]
        if self.pressure:
            f.append(q.f[self.nfields].copy())
        return okr.KVec(f, q.time)

    # ---- GPU-side objects ---------------------------------------------------------------------
    def gpu(self, ctx, ncols):
        import nekstab_next_b200 as nb
        lens = [self.npts] * self.nfields + ([self.np_pr] if self.pressure else [])
        in_dot = [True] * self.nfields + ([False] if self.pressure else [])
        lay = nb.Layout(ctx, lens, in_dot, time_in_dot=self.time_in_dot)
        lay.set_weight([self.bm1] * self.nfields)
        basis = nb.Basis(lay, ncols)
        z = self.coords[2] if self.dim == 3 else None
        semg = nb.Sem(ctx, self.N, self.coords[0], self.coords[1], z, mask=self.mask, glo_num=self.glo)
        op = nb.sem_operator(semg, self.nfields, self.alpha, self.beta, self.h1, self.h2, conv=self.conv)
        return lay, basis, semg, op


def upload(vec, kv):
    vec.upload(kv.f, kv.time)


def download(vec, shape=None):
    """Device vector -> oracle KVec; fields of npts entries are reshaped to ``shape`` if given."""
    f, t = vec.download()
    if shape is not None:
        n = int(np.prod(shape))
        f = [a.reshape(shape) if a.size == n else a for a in f]
    return okr.KVec(f, t)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / den)
