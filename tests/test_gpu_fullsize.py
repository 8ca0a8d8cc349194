"""BASELINE.json's headline configuration at FULL size (box 32^3 elements, N = 7, three components,
50.3 M dof): the oracle cannot run it in seconds, so parity is checked through size-independent
properties of the path, all evaluated on the device through the C ABI:

  * the Arnoldi relation  M V_k = V_k+1 H  column by column,
  * BM1-orthonormality of the basis  ||V^T B V - I|| < 1e-10  (north-star bound),
  * linearity and BM1-self-adjointness of the Helmholtz matvec,
  * dssum: a direct-stiffness-summed field is continuous, so dssum(vmult * dssum(u)) = dssum(u),
  * H of the fused CGS2 path = H of the literal reference ordering (MGS2_REF) to 1e-12,
  * K = 100 like the benchmark, so the kernel that dominates it (the 2-D TMA + register-retention fused sweep,
    k >= 54) is checked with thousands of 64-row blocks per CTA, i.e. through many mbarrier phase flips,
  * replaying the factorisation (CUDA graphs of whole Arnoldi steps) reproduces H bit for bit,
  * DGKS with the device-side decision gives the same H to 1e-11 and an orthonormal basis.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NELX, N, NC, K = 32, 7, 3, 100      # k_dim = 100: the benchmark's own Krylov dimension


@pytest.fixture(scope='module')
def full(ctx):
    import nekstab_next_b200 as nb
    m = nb.mesh.box_mesh(NELX, NELX, NELX, N, deform=0.05)
    sem = nb.Sem(ctx, N, m['x'], m['y'], m['z'], mask=m['mask'], glo_num=m['glo'])
    npts = sem.npts
    lay = nb.Layout(ctx, [npts] * NC, [True] * NC)
    lay.set_weight([sem.get('bm1')] * NC)
    Q = nb.Basis(lay, K + 4)
    # M = I - L / (1.05 rho): spectrum inside the unit disc like the benchmark's operator (rho from a short
    # power iteration on L = B^-1 mask QQ^T (A + 0.1 B))
    Lop = nb.sem_operator(sem, NC, 0.0, 1.0, 1.0, 0.1)
    rng0 = np.random.default_rng(5)
    Q[0].upload([rng0.standard_normal(npts) for _ in range(NC)])
    for f in range(NC):
        sem.dssum(Q[0], f)
        sem.col2(Q[0], f, 'vmult')
        sem.col2(Q[0], f, 'mask')
    rho = 1.0
    for _ in range(12):
        Lop.matvec(Q[0], Q[1])
        rho = nb.k_normalize(Q[1])
        nb.k_copy(Q[0], Q[1])
    Lop.close()
    op = nb.sem_operator(sem, NC, 1.0, -1.0 / (1.05 * rho), 1.0, 0.1)
    rng = np.random.default_rng(0)
    seed = [rng.standard_normal(npts) for _ in range(NC)]
    yield dict(nb=nb, sem=sem, lay=lay, Q=Q, op=op, npts=npts, seed=seed, mask=m['mask'].ravel())
    op.close()
    Q.close()
    sem.close()


def _continuous_seed(full, col):
    """Random field made continuous and masked the way prepare_seed does (dssum, vmult, mask)."""
    nb, sem, Q = full['nb'], full['sem'], full['Q']
    Q[col].upload(full['seed'])
    for f in range(NC):
        sem.dssum(Q[col], f)
        sem.col2(Q[col], f, 'vmult')
        sem.col2(Q[col], f, 'mask')


def test_full_size_arnoldi_relation_and_orthonormality(full):
    nb, Q, op = full['nb'], full['Q'], full['op']
    assert full['npts'] * NC == 50331648
    _continuous_seed(full, 0)
    nb.k_normalize(Q[0])
    H = np.zeros((K + 1, K), order='F')
    nb.arnoldi_factorization(Q, H, 1, K, K, op)
    G = Q.gram(K + 1)
    assert np.max(np.abs(G - np.eye(K + 1))) < 1e-10
    wrk, acc = K + 1, K + 2
    for j in (0, 30, 53, 54, K // 2 + 20, K - 1):
        op.matvec(Q[j], Q[wrk])
        nb.k_matmul(Q[acc], Q, H[:j + 2, j], j + 2)
        nb.k_sub2(Q[wrk], Q[acc])
        assert nb.k_norm(Q[wrk]) < 1e-10 * max(1.0, np.linalg.norm(H[:j + 2, j]))
    # replay (every step of the second run is a captured CUDA graph): identical bits
    Hr = np.zeros((K + 1, K), order='F')
    _continuous_seed(full, 0)
    nb.k_normalize(Q[0])
    nb.arnoldi_factorization(Q, Hr, 1, K, K, op)
    assert np.array_equal(Hr, H)
    # DGKS, decision on the device: same Krylov space, H equal to rounding, basis orthonormal
    Hd = np.zeros((K + 1, K), order='F')
    _continuous_seed(full, 0)
    nb.k_normalize(Q[0])
    nb.arnoldi_factorization(Q, Hd, 1, K, K, op, nb.ORTH_DGKS)
    passes = nb.arnoldi_passes(Q, 1, K, nb.ORTH_DGKS)
    assert set(np.unique(passes)) <= {1, 2}
    Gd = Q.gram(K + 1)
    assert np.max(np.abs(Gd - np.eye(K + 1))) < 1e-10
    for j in (0, 54, K - 1):
        op.matvec(Q[j], Q[wrk])
        nb.k_matmul(Q[acc], Q, Hd[:j + 2, j], j + 2)
        nb.k_sub2(Q[wrk], Q[acc])
        assert nb.k_norm(Q[wrk]) < 1e-10 * max(1.0, np.linalg.norm(Hd[:j + 2, j]))
    # same Krylov space: early columns of H agree to rounding, leading Ritz values to the north-star 1e-6
    assert np.max(np.abs(Hd[:12, :10] - H[:12, :10])) <= 1e-11 * np.max(np.abs(H))
    ev, evd = np.linalg.eigvals(H[:K, :K]), np.linalg.eigvals(Hd[:K, :K])
    lead = np.argsort(-np.abs(ev))[:4]
    for lam in ev[lead]:
        assert np.min(np.abs(evd - lam)) <= 1e-6 * abs(lam)
    # the literal reference ordering (column-by-column MGS, unconditional second pass) gives the same H
    H2 = np.zeros((K + 1, K), order='F')
    _continuous_seed(full, 0)
    nb.k_normalize(Q[0])
    nb.arnoldi_factorization(Q, H2, 1, 6, K, op, nb.ORTH_MGS2_REF)
    assert np.max(np.abs(H2[:7, :6] - H[:7, :6])) <= 1e-12 * np.max(np.abs(H))


def test_full_size_matvec_linearity_and_symmetry(full):
    nb, Q, op = full['nb'], full['Q'], full['op']
    rng = np.random.default_rng(1)
    u, v, w, Mu, Mv = 0, 1, 2, 3, 4
    _continuous_seed(full, u)
    Q[v].upload([rng.standard_normal(full['npts']) for _ in range(NC)])
    for f in range(NC):
        full['sem'].dssum(Q[v], f)
        full['sem'].col2(Q[v], f, 'vmult')
        full['sem'].col2(Q[v], f, 'mask')
    a, b = 0.7, -1.3
    nb.k_copy(Q[w], Q[u])
    Q[w].axpby(a, Q[v], b, skip_time=False)            # w = a u + b v
    op.matvec(Q[u], Q[Mu])
    op.matvec(Q[v], Q[Mv])
    uMv, Muv = nb.k_dot(Q[u], Q[Mv]), nb.k_dot(Q[Mu], Q[v])
    # self-adjoint in the BM1 inner product (relative to |u| |M v|: <u, v> itself is a cancelling sum)
    assert abs(uMv - Muv) <= 1e-12 * nb.k_norm(Q[u]) * nb.k_norm(Q[Mv])
    op.matvec(Q[w], Q[5])
    Q[Mu].axpby(a, Q[Mv], b, skip_time=False)          # a M u + b M v
    nb.k_sub2(Q[5], Q[Mu])
    assert nb.k_norm(Q[5]) <= 1e-12 * nb.k_norm(Q[Mu])


def test_full_size_dssum_projection(full):
    nb, Q, sem = full['nb'], full['Q'], full['sem']
    Q[0].upload(full['seed'])
    for f in range(NC):
        sem.dssum(Q[0], f)
    nb.k_copy(Q[1], Q[0])
    for f in range(NC):
        sem.col2(Q[1], f, 'vmult')
        sem.dssum(Q[1], f)
    nb.k_sub2(Q[1], Q[0])
    assert nb.k_norm(Q[1]) <= 1e-13 * nb.k_norm(Q[0])
