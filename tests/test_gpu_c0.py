"""GPU parity of the C0 (unique-node) storage layout: the same reference semantics -- BM1-weighted inner products of
continuous fields, update_hessenberg_matrix, the SEM operator, arnoldi_factorization -- on vectors stored once per
distinct node (nsb_layout_create_c0), against the oracle and against the element-local layout."""
import numpy as np
import pytest

from helpers import BoxProblem, upload, download, relerr
from oracle import krylov as okr, sem as osem

pytestmark = pytest.mark.gpu


def _c0(P, ctx, ncols, conv=True):
    import nekstab_next_b200 as nb
    z = P.coords[2] if P.dim == 3 else None
    sem = nb.Sem(ctx, P.N, P.coords[0], P.coords[1], z, mask=P.mask, glo_num=P.glo)
    lens = [P.npts] * P.nfields + ([P.np_pr] if P.pressure else [])
    in_dot = [True] * P.nfields + ([False] if P.pressure else [])
    lay = nb.Layout(ctx, lens, in_dot, time_in_dot=P.time_in_dot, c0_sem=sem, n_c0=P.nfields)
    lay.set_weight([P.bm1] * P.nfields)
    B = nb.Basis(lay, ncols)
    op = nb.sem_operator(sem, P.nfields, P.alpha, P.beta, P.h1, P.h2, conv=P.conv if conv else None)
    return lay, B, sem, op


@pytest.mark.parametrize('nel,N', [((3, 2, 2), 7), ((2, 3, 2), 4), ((4, 3), 5)])
def test_c0_layout_vectors_and_dots(ctx, nel, N):
    import nekstab_next_b200 as nb
    P = BoxProblem(nel=nel, N=N, deform=0.05, nfields=len(nel), pressure=True, time_in_dot=True, beta=-1e-3, seed=3)
    c = P.octx()
    lay, B, sem, op = _c0(P, ctx, 4)
    nuni = int(np.unique(P.glo).size)
    assert lay.n_c0 == P.nfields and lay.c0_rows == nuni          # one row per distinct node
    a, b = P.random_kvec(), P.random_kvec()
    upload(B[0], a); upload(B[1], b)
    ga = download(B[0], P.shape)
    for x, y in zip(ga.f, a.f):
        assert np.array_equal(x.ravel(), np.asarray(y).ravel())     # continuous field: exact round trip
    assert ga.time == a.time
    ref = okr.k_dot(c, a, b)
    assert abs(B[0].dot(B[1]) - ref) <= 1e-12 * abs(ref) + 1e-13 * okr.k_norm(c, a) * okr.k_norm(c, b)
    assert abs(B[0].norm() - okr.k_norm(c, a)) <= 1e-13 * okr.k_norm(c, a)
    nb.k_copy(B[2], B[0])
    B[2].axpby(0.5, B[1], -2.0, skip_time=False)
    r = a.copy()
    okr.axpby(r, 0.5, b, -2.0, skip_time=False)
    g = download(B[2], P.shape)
    for x, y in zip(g.f, r.f):
        assert relerr(x.ravel(), np.asarray(y).ravel()) <= 1e-14
    for o in (op, B, lay, sem):
        o.close()


@pytest.mark.parametrize('conv', [False, True])
@pytest.mark.parametrize('nel,N,nf', [((3, 2, 2), 7, 3), ((2, 2, 3), 7, 1), ((3, 3, 2), 4, 2), ((4, 3), 5, 2)])
def test_c0_operator_matches_oracle(ctx, conv, nel, N, nf):
    P = BoxProblem(nel=nel, N=N, deform=0.05, nfields=nf, pressure=True, time_in_dot=True, conv=conv, beta=-1e-3, seed=5)
    lay, B, sem, op = _c0(P, ctx, 2)
    q = P.random_kvec()
    upload(B[0], q)
    op.matvec(B[0], B[1])
    ref = P.omatvec(q)
    got = download(B[1], P.shape)
    for x, y in zip(got.f, ref.f):
        assert relerr(x.ravel(), np.asarray(y).ravel()) <= 1e-12
    assert got.time == ref.time
    for o in (op, B, lay, sem):
        o.close()


@pytest.mark.parametrize('k', [3, 40, 61, 130])
@pytest.mark.parametrize('mode', ['cgs2', 'mgs2', 'dgks'])
def test_c0_orthonormalize_matches_reference_mgs2(ctx, k, mode):
    import nekstab_next_b200 as nb
    from test_gpu_orth import build_basis
    if mode == 'mgs2' and k > 61:
        pytest.skip('literal mode covered up to k = 61')
    P = BoxProblem(nel=(2, 2, 2), N=4, nfields=2, pressure=True, time_in_dot=True, beta=-1e-3, seed=20 + k)
    c = P.octx()
    lay, B, sem, op = _c0(P, ctx, k + 1, conv=False)
    Q = build_basis(P, c, k)
    f = P.random_kvec()
    for q in Q[: max(1, k // 2)]:
        okr.axpby(f, 1.0, q, 50.0, skip_time=False)
    for i, q in enumerate(Q):
        upload(B[i], q)
    upload(B[k], f)
    H = np.zeros((k + 1, k))
    fref = f.copy()
    okr.update_hessenberg_matrix(c, H, fref, Q, k)
    m = dict(cgs2=nb.ORTH_CGS2, mgs2=nb.ORTH_MGS2_REF, dgks=nb.ORTH_DGKS)[mode]
    h, _ = nb.orthonormalize(B, k, k, m)
    assert np.max(np.abs(h - H[:, k - 1])) <= 1e-12 * np.linalg.norm(H[:, k - 1])
    got = download(B[k], P.shape)
    for x, y in zip(got.f, fref.f):
        assert np.max(np.abs(x.ravel() - np.asarray(y).ravel())) <= 1e-11 * max(np.max(np.abs(y)), 1e-300)
    G = B.gram(k + 1)
    assert np.max(np.abs(G - np.eye(k + 1))) < 1e-10
    for o in (op, B, lay, sem):
        o.close()


@pytest.mark.parametrize('conv', [False, True])
def test_c0_arnoldi_matches_oracle_and_element_local_layout(ctx, conv):
    import nekstab_next_b200 as nb
    P = BoxProblem(nel=(3, 2, 2), N=7, nfields=3, conv=conv, seed=7)
    c = P.octx()
    K = 24
    q0 = P.random_kvec()
    okr.k_normalize(c, q0)
    Qo = [okr.k_zero_like(q0) for _ in range(K + 1)]
    okr.k_copy(Qo[0], q0)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(c, P.omatvec, Qo, Ho, 1, K, K)
    lay, B, sem, op = _c0(P, ctx, K + 1, conv=conv)
    upload(B[0], q0)
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(B, H, 1, 10, K, op)
    nb.arnoldi_factorization(B, H, 11, K, K, op)
    assert np.max(np.abs(H - Ho)) <= 1e-10 * np.max(np.abs(Ho))
    G = B.gram(K + 1)
    assert np.max(np.abs(G - np.eye(K + 1))) < 1e-10
    got = download(B[K], P.shape)
    assert relerr(got.f[0].ravel(), np.asarray(Qo[K].f[0]).ravel()) <= 1e-8
    # replay (captured graphs) and the element-local layout give the same H
    upload(B[0], q0)
    Hr = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(B, Hr, 1, K, K, op)
    assert np.array_equal(Hr, H)
    layL, BL, SL, opL = P.gpu(ctx, K + 1)
    upload(BL[0], q0)
    HL = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(BL, HL, 1, K, K, opL)
    assert np.max(np.abs(HL - H)) <= 1e-11 * np.max(np.abs(H))
    # Krylov-Schur restart machinery (rotation) on the C0 basis
    Z = np.linalg.qr(np.random.default_rng(0).standard_normal((K, K)))[0]
    B.rotate(K, Z)
    G = B.gram(K)
    assert np.max(np.abs(G - np.eye(K))) < 1e-10
    for o in (op, B, lay, sem, opL, BL, layL, SL):
        o.close()


def test_c0_host_operator_and_errors(ctx):
    """Host operator on a C0 basis (element-local arrays cross the boundary as before); the time-stepper pieces
    refuse the layout instead of mis-indexing it."""
    import nekstab_next_b200 as nb
    P = BoxProblem(nel=(2, 2, 2), N=4, nfields=2, seed=8)
    c = P.octx()
    K = 6
    lay, B, sem, op = _c0(P, ctx, K + 1, conv=False)

    def host_mv(fields, t):
        return [P.m_apply_field(f.reshape(P.shape)).ravel() for f in fields], t

    hop = nb.host_operator(lay, host_mv, linear=True)
    q0 = P.random_kvec()
    okr.k_normalize(c, q0)
    upload(B[0], q0)
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(B, H, 1, K, K, hop)
    Qo = [okr.k_zero_like(q0) for _ in range(K + 1)]
    okr.k_copy(Qo[0], q0)
    Ho = np.zeros((K + 1, K))
    okr.arnoldi_factorization(c, P.omatvec, Qo, Ho, 1, K, K)
    assert np.max(np.abs(H - Ho)) <= 1e-11 * np.max(np.abs(Ho))
    with pytest.raises(nb.NsbError):
        sem.dssum(B[0], 0)                      # element-local kernels do not accept unique-node fields
    with pytest.raises(nb.NsbError):
        sem.hmholtz(B[0], B[1], 0, 1.0, 1.0)
    for o in (hop, op, B, lay, sem):
        o.close()
