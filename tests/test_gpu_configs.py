"""BASELINE.json configs[0..2] as parity cases: the reference's own curved 2-D meshes
(examples/cylinder, examples/back_fstep; committed as tests/golden/*.npz) with the synthetic SEM
operator standing in for the Navier-Stokes time-stepper (out of scope).

  configs[0]  cylinder, direct Arnoldi eigensolve, k_dim = 100          -> test_cylinder_arnoldi_k100
  configs[1]  cylinder, Newton-Krylov fixed point: the GMRES inner solve -> test_cylinder_ts_gmres
  configs[2]  back_fstep, transient growth (direct-adjoint Arnoldi)      -> test_bfs_direct_adjoint
Tolerances: Ritz values 1e-6 relative, ||V^T B V - I|| < 1e-10 (BASELINE.json north_star).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from helpers import upload, download
from oracle import sem as osem, krylov as okr

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / 'golden'


class MeshCase:
    def __init__(self, name, ctx, ncols, conv=True):
        import nekstab_next_b200 as nb
        g = np.load(GOLD / f'{name}_mesh.npz')
        self.N = json.loads((GOLD / 'known_answers.json').read_text())[name]['N']
        self.x, self.y, self.U, self.V = g['x'], g['y'], g['u'], g['v']
        self.glo = g['glo'].astype(np.int64)
        self.geo = osem.geometry(self.N, self.x, self.y)
        self.d = osem.dgll(self.N)
        self.bm1 = self.geo['bm1']
        self.binv = 1.0 / osem.dssum(self.bm1, self.glo)
        self.vmult = 1.0 / osem.multiplicity(self.glo)
        self.mask = np.ones_like(self.x)
        self.conv = (self.U, self.V) if conv else None      # the reference's base flow convects the perturbation
        self.h1, self.h2, self.alpha = 0.05, 0.1, 1.0
        self.shape, self.npts = self.x.shape, self.x.size
        rng = np.random.default_rng(3)
        u = osem.dssum(rng.standard_normal(self.shape), self.glo) * self.vmult
        lam = 1.0
        for _ in range(25):
            v = self.l_apply(u)
            lam = np.sqrt(osem.glsc3(v, v, self.bm1) / osem.glsc3(u, u, self.bm1))
            u = v / lam
        self.beta = -1.0 / (1.05 * lam)
        self.rng = rng
        self.octx = okr.Ctx(bm1s=self.bm1, in_dot=[True, True])
        self.lay = nb.Layout(ctx, [self.npts, self.npts], [True, True])
        self.lay.set_weight([self.bm1, self.bm1])
        self.B = nb.Basis(self.lay, ncols)
        self.sem = nb.Sem(ctx, self.N, self.x, self.y, None, mask=None, glo_num=self.glo)
        self.op = nb.sem_operator(self.sem, 2, self.alpha, self.beta, self.h1, self.h2, conv=self.conv)

    def l_apply(self, u):
        w = osem.axhelm(u, self.geo['g'], self.d, self.h1, self.h2, self.bm1)
        if self.conv is not None:
            _, wq = osem.gll(self.N)
            w = w + (wq[:, None] * wq[None, :])[None] * osem.convect(u, self.conv, self.geo['rst'], self.d)
        return osem.dssum(w, self.glo) * self.mask * self.binv

    def omatvec(self, q):
        return okr.KVec([self.alpha * f + self.beta * self.l_apply(f) for f in q.f], q.time)

    def seed(self):
        f = [osem.dssum(self.rng.standard_normal(self.shape), self.glo) * self.vmult for _ in range(2)]
        q = okr.KVec(f, 0.0)
        okr.k_normalize(self.octx, q)
        return q


def test_cylinder_arnoldi_k100(ctx):
    import nekstab_next_b200 as nb
    K = 100
    C = MeshCase('cyl', ctx, K + 1)
    q0 = C.seed()
    Qo = [okr.k_zero_like(q0) for _ in range(K + 1)]
    okr.k_copy(Qo[0], q0)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(C.octx, C.omatvec, Qo, Ho, 1, K, K)
    upload(C.B[0], q0)
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(C.B, H, 1, K, K, C.op)
    assert np.max(np.abs(H - Ho)) <= 1e-9 * np.max(np.abs(Ho))
    G = C.B.gram(K + 1)
    assert np.max(np.abs(G - np.eye(K + 1))) < 1e-10
    _, vals = nb.eig(H[:K, :K])
    _, vals_o = okr.eig(Ho[:K, :K])
    assert np.max(np.abs(vals[:10] - vals_o[:10]) / np.abs(vals_o[:10])) < 1e-6
    assert np.max(np.abs(vals.imag)) > 0          # convection by the base flow: complex Ritz pairs


def test_cylinder_ts_gmres(ctx):
    import nekstab_next_b200 as nb
    ks = 20
    C = MeshCase('cyl', ctx, ks + 2)
    W = nb.Basis(C.lay, 2)
    rhs = C.seed()
    upload(W[0], rhs)
    sol_ref, hist_ref, calls_ref = okr.ts_gmres(C.octx, C.omatvec, rhs, maxiter=3, ksize=ks, tol=1e-20)
    hist, calls = nb.ts_gmres(C.B, C.op, W[0], W[1], maxiter=3, ksize=ks, tol=1e-20)
    assert calls == calls_ref and np.allclose(hist, hist_ref, rtol=1e-6)
    got = download(W[1])
    for a, b in zip(got.f, sol_ref.f):
        assert np.max(np.abs(a - b.ravel())) <= 1e-8 * np.max(np.abs(b))


def test_bfs_direct_adjoint(ctx):
    """Transient-growth map = adjoint(forward(q)) (core/matvec.f90:478-495).  With the symmetric
    operator (no convection) the B-adjoint is the operator itself, so the gains are the squares of
    its eigenvalues; checked against the oracle applying the map twice."""
    import nekstab_next_b200 as nb
    K = 40
    C = MeshCase('bfs', ctx, K + 1, conv=False)
    tg = nb.compose_operators(C.lay, C.op, C.op)
    q0 = C.seed()
    Qo = [okr.k_zero_like(q0) for _ in range(K + 1)]
    okr.k_copy(Qo[0], q0)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(C.octx, lambda q: C.omatvec(C.omatvec(q)), Qo, Ho, 1, K, K)
    upload(C.B[0], q0)
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(C.B, H, 1, K, K, tg)
    assert np.max(np.abs(H - Ho)) <= 1e-9 * np.max(np.abs(Ho))
    _, gains = nb.eig(H[:K, :K])
    _, gains_o = okr.eig(Ho[:K, :K])
    assert np.max(np.abs(gains[:6] - gains_o[:6]) / np.abs(gains_o[:6])) < 1e-6
    assert np.max(np.abs(gains.imag)) < 1e-8 and gains.real.min() > -1e-10     # A^dagger A: real, non-negative
    assert np.max(np.abs(C.B.gram(K + 1) - np.eye(K + 1))) < 1e-10
    assert tg.count() == K and C.op.count() == 2 * K
