"""The C restatement (CPU baseline / checker) agrees with the numpy oracle."""
import numpy as np

from helpers import BoxProblem
from oracle import cref, krylov as okr, sem as osem


def test_c_axhelm_dssum_match_numpy():
    P = BoxProblem(nel=(2, 3, 2), N=7, deform=0.05, nfields=1, seed=2)
    u = P.rng.standard_normal(P.shape)
    g = np.ascontiguousarray(P.geo['g'])
    for h1, h2 in ((1.0, 0.0), (0.6, 0.2)):
        a = cref.axhelm3d(u, g, P.bm1, P.d, h1, h2)
        b = osem.axhelm(u, P.geo['g'], P.d, h1, h2, P.bm1)
        assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b))
    off, idx = cref.gs_lists(P.glo)
    assert np.max(np.abs(cref.dssum(u, off, idx) - osem.dssum(u, P.glo))) < 1e-14


def test_c_arnoldi_matches_numpy():
    P = BoxProblem(nel=(2, 2, 2), N=5, deform=0.04, nfields=3, seed=3)
    c = P.octx()
    K = 10
    q0 = P.random_kvec()
    okr.k_normalize(c, q0)
    Qo = [okr.k_zero_like(q0) for _ in range(K + 1)]
    okr.k_copy(Qo[0], q0)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(c, P.omatvec, Qo, Ho, 1, K, K)
    n = 3 * P.npts
    Q = np.zeros((K + 1, n))
    Q[0] = np.concatenate([f.ravel() for f in q0.f])
    H = np.zeros((K + 1, K), order='F')
    off, idx = cref.gs_lists(P.glo)
    cref.arnoldi(Q, H, 0, K - 1, np.ascontiguousarray(P.bm1), np.ascontiguousarray(P.geo['g']),
                 np.ascontiguousarray(P.bm1), np.ascontiguousarray(P.binv), np.ascontiguousarray(P.mask),
                 P.d, P.N + 1, P.shape[0], 3, off, idx, P.h1, P.h2, P.alpha, P.beta)
    assert np.max(np.abs(H - Ho)) <= 1e-11 * np.max(np.abs(Ho))
    for j in range(K + 1):
        ref = np.concatenate([f.ravel() for f in Qo[j].f])
        assert np.max(np.abs(Q[j] - ref)) <= 1e-10 * np.max(np.abs(ref))
