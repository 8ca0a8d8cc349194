"""GPU parity: SEM geometry, axhelm (sum factorisation with G1..G6), dssum, mask, and the fused
operator vs the oracle; 1e-12 relative in fp64 (BASELINE.json north_star)."""
import json
from pathlib import Path

import numpy as np
import pytest

from helpers import BoxProblem, upload, download, relerr
from oracle import sem as osem

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / 'golden'
TOL = 1e-12


@pytest.mark.parametrize('nel,N,deform', [((3, 2, 2), 7, 0.05), ((2, 2, 3), 7, 0.0), ((2, 3, 2), 4, 0.05),
                                          ((2, 2, 2), 1, 0.02), ((2, 2, 2), 10, 0.03), ((5, 4), 5, 0.04),
                                          ((3, 3), 7, 0.0)])
def test_geometry_and_kernels(ctx, nel, N, deform):
    P = BoxProblem(nel=nel, N=N, deform=deform, nfields=1, seed=N)
    lay, B, S, op = P.gpu(ctx, 4)
    # geometry
    assert relerr(S.get('bm1'), P.bm1) <= TOL
    assert relerr(S.get('jac'), P.geo['jac']) <= TOL
    names = ['g1', 'g2', 'g3', 'g4', 'g5', 'g6'] if P.dim == 3 else ['g1', 'g2', 'g4']
    gscale = np.max(np.abs(P.geo['g']))
    for i, nm in enumerate(names):
        assert np.max(np.abs(S.get(nm) - P.geo['g'][i])) <= TOL * gscale
    assert relerr(S.get('binvm1'), P.binv) <= TOL
    assert np.array_equal(S.get('vmult'), P.vmult)
    assert np.array_equal(S.get('mask'), P.mask)
    # axhelm on a non-continuous random field (element-local operator)
    u = P.rng.standard_normal(P.shape)
    B[0].upload([u])
    for h1, h2 in ((1.0, 0.0), (0.7, 0.3)):
        S.axhelm(B[0], B[1], 0, h1, h2)
        ref = osem.axhelm(u, P.geo['g'], P.d, h1, h2, P.bm1)
        assert relerr(B[1].download()[0][0], ref.ravel()) <= TOL
    # dssum, col2, ax
    S.dssum(B[0], 0)
    assert relerr(B[0].download()[0][0], osem.dssum(u, P.glo).ravel()) <= 1e-14
    B[0].upload([u])
    S.ax(B[0], B[2], 0, 1.0, 0.1)
    ref = osem.ax(u, P.geo['g'], P.d, P.glo, P.mask, 1.0, 0.1, P.bm1)
    assert relerr(B[2].download()[0][0], ref.ravel()) <= TOL
    S.col2(B[2], 0, 'binvm1')
    assert relerr(B[2].download()[0][0], (ref * P.binv).ravel()) <= TOL


@pytest.mark.parametrize('conv', [False, True])
@pytest.mark.parametrize('nel,N', [((3, 2, 2), 7), ((4, 3), 5)])
def test_fused_operator(ctx, conv, nel, N):
    P = BoxProblem(nel=nel, N=N, deform=0.05, nfields=len(nel), pressure=True, time_in_dot=True, conv=conv, seed=3)
    lay, B, S, op = P.gpu(ctx, 3)
    q = P.random_kvec()
    upload(B[0], q)
    op.matvec(B[0], B[1])
    ref = P.omatvec(q)
    got = download(B[1])
    for a, b in zip(got.f, ref.f):
        assert relerr(a, b.ravel()) <= TOL
    assert got.time == ref.time
    assert op.count() == 1


@pytest.mark.parametrize('nf', [1, 2])
def test_fused_operator_one_and_two_components_with_convection(ctx, nf):
    """The N = 7 tensor-core kernel has one instantiation per component count; these two (with the convective
    term staged in the ring) are not reached by the three-component cases above."""
    P = BoxProblem(nel=(3, 2, 2), N=7, deform=0.05, nfields=nf, conv=True, seed=4)
    lay, B, S, op = P.gpu(ctx, 3)
    q = P.random_kvec()
    upload(B[0], q)
    op.matvec(B[0], B[1])
    ref = P.omatvec(q)
    got = download(B[1])
    for a, b in zip(got.f, ref.f):
        assert relerr(a, b.ravel()) <= TOL


@pytest.mark.parametrize('nf', [1, 2])
def test_fused_convective_operator_large_mesh(ctx, nf):
    """axhelm3d_dmma8_kernel<1, conv> and <2, conv> on a 24^3-element mesh (7.1 M points): the ring protocol of
    these instantiations (4 / 4 warp groups over 4 buffers) runs ~93 elements per CTA, i.e. dozens of mbarrier
    phase flips per buffer -- the size at which the first ring protocol failed to launch."""
    P = BoxProblem(nel=(24, 24, 24), N=7, deform=0.05, nfields=nf, conv=True, beta=-1e-4, seed=9)
    lay, B, S, op = P.gpu(ctx, 2)
    q = P.random_kvec()
    upload(B[0], q)
    op.matvec(B[0], B[1])
    ref = P.omatvec(q)
    got = download(B[1])
    for a, b in zip(got.f, ref.f):
        assert relerr(a, b.ravel()) <= TOL
    for o in (op, S, B, lay):
        o.close()


def test_seed_noise_device_composition(ctx):
    """op_add_noise (core/utils.f90:297-359): mth_rand noise -> opdssum -> opcolv VMULT -> dsavg -> bcdirVC,
    the averaging and masking on the device, against the same composition of oracle pieces (independent dssum),
    plus one value of mth_rand worked out by hand from the formula at :412-414."""
    import nekstab_next_b200 as nb
    from oracle import krylov as okr
    P = BoxProblem(nel=(3, 2, 2), N=5, deform=0.03, nfields=3, seed=12)
    lay, B, S, op = P.gpu(ctx, 1)
    nrm = nb.seed.seed_noise(S, B[0], P.coords, e0=0)
    raw = nb.seed.noise_fields(P.coords, 0)
    ref = []
    for c in range(3):
        v = osem.dssum(raw[c], P.glo) * P.vmult
        v = osem.dssum(v, P.glo) * P.vmult
        ref.append(v * P.mask)
    kv = okr.KVec(ref, 0.0)
    nref = okr.k_normalize(P.octx(), kv)
    assert abs(nrm - nref) <= 1e-12 * nref
    got = download(B[0])
    for a, b in zip(got.f, kv.f):
        assert relerr(a, b.ravel()) <= 1e-12
    # hand-evaluated: element 1, local point (2,1,1) -> ix=2, iy=1, iz=1, ieg=1
    import math
    x, y, z = (a[0, 0, 0, 1] for a in P.coords)
    fc = (3.0e4, -1.5e3, 0.5e5)
    r = fc[0] * (1 + x * math.sin(y)) + fc[1] * 2 * 1 + fc[2] * 2
    r = fc[0] * (1 + z * math.sin(r)) + fc[1] * 1 * 2 + fc[2] * 1
    r = math.cos(1.0e3 * math.sin(1.0e3 * math.sin(r)))
    assert abs(raw[0][0, 0, 0, 1] - r) <= 1e-9       # cos/sin of arguments ~1e5: last digits depend on libm
    for o in (op, S, B, lay):
        o.close()


@pytest.mark.parametrize('name', ['cyl', 'bfs'])
def test_reference_meshes(ctx, name):
    """Config 1: the reference's own curved 2-D meshes (examples/cylinder, examples/back_fstep)."""
    import nekstab_next_b200 as nb
    known = json.loads((GOLD / 'known_answers.json').read_text())[name]
    g = np.load(GOLD / f'{name}_mesh.npz')
    x, y, u, v, glo = g['x'], g['y'], g['u'], g['v'], g['glo'].astype(np.int64)
    N = known['N']
    S = nb.Sem(ctx, N, x, y, None, mask=None, glo_num=glo)
    bm1 = S.get('bm1')
    assert abs(bm1.sum() - known['sum_bm1']) <= 1e-12 * known['sum_bm1']
    geo = osem.geometry(N, x, y)
    assert relerr(bm1, geo['bm1']) <= TOL
    lay = nb.Layout(ctx, [x.size, x.size], [True, True])
    lay.set_weight([bm1, bm1])
    B = nb.Basis(lay, 3)
    B[0].upload([u, v])
    uu = B[0].dot(B[0])
    assert abs(uu - known['uu']) <= 1e-12 * known['uu']
    S.ax(B[0], B[1], 0, 1.0, 0.0)
    ref = osem.ax(u, geo['g'], osem.dgll(N), glo, np.ones_like(u), 1.0, 0.0, geo['bm1'])
    assert relerr(B[1].download()[0][0], ref.ravel()) <= TOL
    assert int(glo.max()) + 1 == known['nunique']


def test_mask_inconsistency_is_rejected(ctx):
    import nekstab_next_b200 as nb
    P = BoxProblem(nel=(2, 2, 2), N=3, nfields=1)
    bad = P.mask.copy()
    bad[0, -1, 1, 1] = 0.0   # a face node shared with the element above, zeroed on one side only
    with pytest.raises(nb.NsbError):
        nb.Sem(ctx, P.N, *P.coords, mask=bad, glo_num=P.glo)


@pytest.mark.parametrize('nel,N,deform', [((3, 2, 2), 7, 0.05), ((2, 2, 2), 4, 0.0), ((4, 3), 5, 0.04)])
def test_helmholtz_pcg(ctx, nel, N, deform):
    """nsb_sem_hmholtz vs the oracle restatement of Nek's cggo: same iteration count, same solution,
    and the solution satisfies the assembled Helmholtz problem."""
    P = BoxProblem(nel=nel, N=N, deform=deform, nfields=1, seed=5)
    lay, B, S, op = P.gpu(ctx, 4)
    h1, h2 = 0.7, 2.0
    # right-hand side in the range of the assembled operator: B f, summed over copies, masked
    f = P.random_field()
    rhs = osem.dssum(P.bm1 * f, P.glo) * P.mask
    xo, ito, reso = osem.cggo(rhs, P.geo['g'], P.d, P.glo, P.mask, P.bm1, h1, h2, tol=1e-11, maxit=400)
    B[0].upload([rhs])
    it, res = S.hmholtz(B[0], B[1], 0, h1, h2, tol=1e-11, maxit=400)
    x = B[1].download()[0][0].reshape(P.shape)
    assert abs(it - ito) <= 1 and res <= 1e-11
    assert relerr(x, xo) <= 1e-8
    lhs = osem.dssum(osem.axhelm(x, P.geo['g'], P.d, h1, h2, P.bm1), P.glo) * P.mask
    assert relerr(lhs, rhs) <= 1e-8


def test_helmholtz_pcg_three_systems_side_by_side(ctx):
    """nsb_sem_hmholtz_vec: each of the three systems follows the iteration sequence of its own single solve."""
    P = BoxProblem(nel=(3, 2, 2), N=7, deform=0.05, nfields=3, seed=9)
    lay, B, S, op = P.gpu(ctx, 3)
    h1, h2 = 0.3, 5.0
    rhs = [osem.dssum(P.bm1 * P.random_field(), P.glo) * P.mask * s for s in (1.0, 1e-3, 40.0)]
    rhs[1] = rhs[1] * (np.sin(7 * P.coords[0]) + 1.2)          # a different, rougher right-hand side
    rhs[1] = osem.dssum(rhs[1] * P.vmult, P.glo) * P.mask
    B[0].upload(rhs)
    its, ress = S.hmholtz_vec(B[0], B[1], 0, 3, h1, h2, tol=1e-11, maxit=400)
    xs = B[1].download()[0]
    single_its = []
    for f in range(3):
        xo, ito, _ = osem.cggo(rhs[f], P.geo['g'], P.d, P.glo, P.mask, P.bm1, h1, h2, tol=1e-11, maxit=400)
        assert abs(its[f] - ito) <= 1 and ress[f] <= 1e-11
        assert relerr(xs[f].reshape(P.shape), xo) <= 1e-8
        it1, _ = S.hmholtz(B[0], B[2], f, h1, h2, tol=1e-11, maxit=400)
        single_its.append(it1)
        # same iterates as the stand-alone solve up to the summation order of (w, p): that inner product is
        # accumulated in the axhelm epilogue, whose element-to-warp assignment depends on the number of systems
        assert relerr(B[2].download()[0][f], xs[f]) <= 1e-10
    assert all(abs(a - b) <= 1 for a, b in zip(its, single_its)) and len(set(its)) > 1   # they stop at different iterations
