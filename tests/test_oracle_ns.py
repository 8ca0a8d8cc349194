"""The pressure-coupled perturbation step of oracle/ns.py ([UPSTREAM-RECALL] Nek5000 P_N - P_N-2 path, parity unpinned)
against mathematics that does not depend on the restatement: transposition, exactness on polynomials, the null space
and symmetry of the consistent Poisson operator, discrete incompressibility, linearity and the temporal order."""
import numpy as np
import pytest

from oracle import ns as ons
from oracle import sem as osem


def setup(dim, nel, N, deform=0.05, lxd=None):
    if dim == 3:
        x, y, z, glo = osem.box_mesh(*nel, N, deform=deform)
        x0, y0, z0, _ = osem.box_mesh(*nel, N)
        mask = osem.boundary_mask_box(None, x0, y0, z0)
        coords = (x, y, z)
    else:
        x, y, glo = osem.box_mesh_2d(*nel, N, deform=deform)
        x0, y0, _ = osem.box_mesh_2d(*nel, N)
        mask = osem.boundary_mask_box(None, x0, y0, None, lengths=(1.0, 1.0))
        coords = (x, y)
    geo = osem.geometry(N, *coords)
    ps = ons.pressure_setup(N, geo)
    dl = osem.dealias_setup(N, lxd or (3 * (N + 1) + 1) // 2, geo['rst'])
    return dict(coords=coords, glo=glo, mask=mask, geo=geo, ps=ps, dl=dl, N=N, dim=dim,
                binv=1.0 / osem.dssum(geo['bm1'], glo), vmult=1.0 / osem.multiplicity(glo))


def c0_field(m, rng):
    u = rng.standard_normal(m['coords'][0].shape)
    return osem.dssum(u, m['glo']) * m['vmult'] * m['mask']


@pytest.mark.parametrize('dim,nel,N', [(2, (3, 2), 5), (3, (2, 2, 2), 4), (3, (2, 1, 2), 7)])
def test_gradt_is_the_transpose_of_div(dim, nel, N):
    m = setup(dim, nel, N)
    rng = np.random.default_rng(1)
    vel = [rng.standard_normal(m['coords'][0].shape) for _ in range(dim)]
    p = rng.standard_normal(m['ps']['bm2'].shape)
    lhs = np.sum(ons.opdiv(vel, m['ps']) * p)
    rhs = sum(np.sum(v * g) for v, g in zip(vel, ons.opgradt(p, m['ps'])))
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)


@pytest.mark.parametrize('dim,nel,N', [(2, (2, 2), 5), (3, (2, 2, 1), 5)])
def test_div_is_exact_on_polynomials(dim, nel, N):
    """Undeformed mesh: sum_q (D u)_q = int div u exactly, and (D u)_q / bm2_q = div u pointwise for polynomial u."""
    m = setup(dim, nel, N, deform=0.0)
    c = m['coords']
    if dim == 2:
        x, y = c
        vel = [x ** 3 * y, -1.5 * x ** 2 * y ** 2 + x]          # div = 3 x^2 y - 3 x^2 y = 0
        div_exact = 0 * x
        vel2 = [x ** 2 * y, x * y ** 3]
        div2 = lambda X, Y: 2 * X * Y + 3 * X * Y ** 2
    else:
        x, y, z = c
        vel = [x ** 2 * y * z, -x * y ** 2 * z + z ** 3, x * y]   # div = 2xyz - 2xyz + 0 = 0
        vel2 = [x ** 2 * z, y ** 3, x * z ** 2]
        div2 = lambda X, Y, Z: 2 * X * Z + 3 * Y ** 2 + 2 * X * Z
    ps = m['ps']
    assert np.max(np.abs(ons.opdiv(vel, ps))) <= 1e-13
    cp = [osem.interp_fine(a, ps['I12']) for a in c]              # coordinates of the pressure points
    got = ons.opdiv(vel2, ps) / ps['bm2']
    assert np.max(np.abs(got - div2(*cp))) <= 1e-11


@pytest.mark.parametrize('dim,nel,N,deform', [(2, (3, 3), 4, 0.0), (3, (2, 2, 2), 4, 0.0), (2, (3, 3), 4, 0.05)])
def test_consistent_poisson_operator(dim, nel, N, deform):
    """E = D B^-1 D^T: symmetric, positive semi-definite; on affine elements (where the Gauss rule integrates
    div w exactly) the constants are its null space for an all-Dirichlet velocity, on deformed ones they are only
    nearly so -- as in Nek, where `ortho` removes the mean regardless."""
    m = setup(dim, nel, N, deform=deform)
    ps = m['ps']
    n2 = ps['bm2'].size
    E = np.zeros((n2, n2))
    for j in range(n2):
        e = np.zeros(n2)
        e[j] = 1.0
        E[:, j] = ons.cdabdtp(e.reshape(ps['bm2'].shape), ps, m['glo'], m['mask'], m['binv']).ravel()
    assert np.max(np.abs(E - E.T)) <= 1e-12 * np.max(np.abs(E))
    ev = np.linalg.eigvalsh(0.5 * (E + E.T))
    assert ev[0] >= -1e-12 * ev[-1]
    if deform == 0.0:
        assert np.max(np.abs(E @ np.ones(n2))) <= 1e-12 * np.max(np.abs(E))
        assert ev[1] > 1e-8 * ev[-1], 'only the constant mode is singular'
    else:
        assert np.max(np.abs(E @ np.ones(n2))) <= 1e-2 * np.max(np.abs(E))
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n2)
    rhs = E @ (x - x.mean())
    rhs -= rhs.mean()
    dp, it, drop = ons.esolve(rhs.reshape(ps['bm2'].shape), ps, m['glo'], m['mask'], m['binv'], tol=1e-12, maxit=5000)
    res = E @ dp.ravel() - rhs
    # deformed elements: the iteration works on mean-free vectors while E 1 is only nearly zero -- the solve is exact
    # up to that defect, like Nek's
    assert np.max(np.abs(res - res.mean())) <= (1e-8 if deform == 0.0 else 1e-4) * np.max(np.abs(rhs)), (it, drop)


@pytest.mark.parametrize('dim,nel,N,conv,deform', [(2, (3, 3), 5, True, 0.0), (3, (2, 2, 2), 4, True, 0.0),
                                                   (2, (3, 3), 5, False, 0.05), (3, (2, 2, 2), 4, True, 0.05)])
def test_step_is_discretely_incompressible_and_linear(dim, nel, N, conv, deform):
    m = setup(dim, nel, N, deform=deform)
    rng = np.random.default_rng(5)
    c = m['coords']
    base = None
    if conv:
        tp = 2 * np.pi
        base = [np.sin(tp * c[0]) * np.cos(tp * c[1]), -np.cos(tp * c[0]) * np.sin(tp * c[1])]
        if dim == 3:
            base.append(0.2 + 0 * c[0])
    va = [c0_field(m, rng) for _ in range(dim)]
    vb = [c0_field(m, rng) for _ in range(dim)]
    pa = rng.standard_normal(m['ps']['bm2'].shape)
    pb = rng.standard_normal(m['ps']['bm2'].shape)
    # deformed elements: E is regular (its smallest eigenvalue is the quadrature defect of int div w), so the pressure
    # system is solved as it stands; removing the mean there (Nek's ortho) leaves an O(1e-6) inconsistency that also
    # breaks the linearity of the step at that level
    run = lambda v, p: ons.ns_steps(m['glo'], m['mask'], m['geo'], N, m['ps'], m['dl'], base, v, p, 0.05, 2e-3, 4,
                                    mean_free=(deform == 0.0))
    ua, qa = run(va, pa)
    ub, qb = run(vb, pb)
    us, qs = run([a + 2.0 * b for a, b in zip(va, vb)], pa + 2.0 * pb)
    scale = max(np.max(np.abs(u)) for u in ua)
    assert np.max(np.abs(ons.opdiv(ua, m['ps']))) <= 1e-9 * scale          # D v = 0 after every step
    for a, b, s in zip(ua, ub, us):
        assert np.max(np.abs(s - (a + 2.0 * b))) <= 1e-8 * scale
    for u in ua:                                                           # continuous, zero on the walls
        assert np.max(np.abs(osem.dssum(u, m['glo']) * m['vmult'] - u)) <= 1e-12 * scale
        assert np.max(np.abs(u * (1 - m['mask']))) == 0.0


def test_stokes_mode_decay_and_temporal_order():
    """Stokes flow in the unit square (no base flow): the kinetic energy of a solenoidal start decays, and halving dt
    reduces the difference to a fine-step reference by about the order of the scheme (BDF ramp 1-2-3: at least 2)."""
    N = 6
    m = setup(2, (2, 2), N, deform=0.0)
    x, y = m['coords']
    # stream function psi = sin^2(pi x) sin^2(pi y): u = psi_y, v = -psi_x vanish on the walls, div = 0
    s, c_ = np.sin, np.cos
    u0 = 2 * np.pi * s(np.pi * x) ** 2 * s(np.pi * y) * c_(np.pi * y)
    v0 = -2 * np.pi * s(np.pi * x) * c_(np.pi * x) * s(np.pi * y) ** 2
    p0 = 0 * m['ps']['bm2']
    T, nu = 0.02, 0.1
    run = lambda nst: ons.ns_steps(m['glo'], m['mask'], m['geo'], N, m['ps'], m['dl'], None, [u0, v0], p0, nu, T / nst, nst)
    ref, _ = run(64)
    e = []
    for nst in (4, 8, 16):
        v, _ = run(nst)
        e.append(np.sqrt(sum(osem.glsc3(a - b, a - b, m['geo']['bm1']) for a, b in zip(v, ref))))
    ke0 = sum(osem.glsc3(a, a, m['geo']['bm1']) for a in (u0, v0))
    ke1 = sum(osem.glsc3(a, a, m['geo']['bm1']) for a in ref)
    assert 0.0 < ke1 < ke0
    assert e[0] / e[1] > 2.5 and e[1] / e[2] > 2.5, e


@pytest.mark.parametrize('dim,nel,N,deform', [(2, (6, 6), 5, 0.05), (3, (3, 3, 3), 4, 0.04)])
def test_two_level_preconditioner(dim, nel, N, deform):
    """Element-wise fast diagonalisation + coarse level: symmetric positive definite, the sparse coarse operator equals
    R E R^T obtained by probing, and the preconditioned iteration converges in far fewer steps to the same solution."""
    m = setup(dim, nel, N, deform=deform)
    ps = m['ps']
    shp = ps['bm2'].shape
    n2, ne = ps['bm2'].size, shp[0]
    A = lambda p: ons.cdabdtp(p, ps, m['glo'], m['mask'], m['binv'])
    fd = ons.coarse_setup(ons.fdm_setup(N, m['geo'], ps), ps, m['glo'], m['mask'], m['binv'])
    for e in (0, ne // 2, ne - 1):
        v = np.zeros(shp)
        v[e] = 1.0
        col = A(v).reshape(ne, -1).sum(axis=1)
        assert np.max(np.abs(col - fd['coarse']['Ec'][:, e])) <= 1e-13 * np.max(np.abs(col))
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal(shp), rng.standard_normal(shp)
    Ma, Mb = ons.fdm_apply(a, fd), ons.fdm_apply(b, fd)
    assert abs(np.sum(Ma * b) - np.sum(a * Mb)) <= 1e-12 * abs(np.sum(Ma * b))
    assert np.sum(Ma * a) > 0 and np.sum(Mb * b) > 0
    x = rng.standard_normal(shp)
    x -= x.mean()
    rhs = A(x)
    x0, it0, _ = ons.esolve(rhs, ps, m['glo'], m['mask'], m['binv'], tol=1e-10, maxit=4000, mean_free=False)
    x1, it1, _ = ons.esolve(rhs, ps, m['glo'], m['mask'], m['binv'], tol=1e-10, maxit=4000, mean_free=False, fdm=fd)
    assert it1 < 0.5 * it0
    assert np.max(np.abs((x1 - x1.mean()) - (x0 - x0.mean()))) <= 1e-6 * np.max(np.abs(x0))


def test_adjoint_stepper_is_dual_to_the_forward_one():
    """The adjoint stepper integrates the CONTINUOUS adjoint equations with the same scheme (what Nek does for
    exponential_prop%rmatvec), so <A v, w>_B = <v, A+ w>_B holds up to discretisation errors, not to rounding:
    1.7e-6 relative on this resolved problem at dt = 2.5e-3 (4.5e-7 at half the step: first order in dt) -- and 8.6e-3,
    five thousand times more, with the forward operator in place of A+."""
    N = 7
    m = setup(2, (3, 3), N, deform=0.0)
    x, y = m['coords']
    base = [np.sin(np.pi * x) ** 2 * np.sin(2 * np.pi * y), -np.sin(2 * np.pi * x) * np.sin(np.pi * y) ** 2]   # solenoidal, no-slip
    rng = np.random.default_rng(4)

    def smooth_solenoidal(kx, ky):
        psi_y = np.sin(kx * np.pi * x) ** 2 * ky * np.pi * np.sin(2 * ky * np.pi * y) / 1.0
        psi_x = kx * np.pi * np.sin(2 * kx * np.pi * x) * np.sin(ky * np.pi * y) ** 2
        return [psi_y, -psi_x]

    v0, w0 = smooth_solenoidal(1, 1), smooth_solenoidal(1, 2)
    p0 = 0 * m['ps']['bm2']
    nu, dt, nst = 0.02, 2.5e-3, 40
    run = lambda q, adj: ons.ns_steps(m['glo'], m['mask'], m['geo'], N, m['ps'], m['dl'], base, q, p0, nu, dt, nst,
                                      tol_v=1e-12, tol_p=1e-12, adjoint=adj)[0]
    Av, Atw, Aw = run(v0, False), run(w0, True), run(w0, False)
    dot = lambda a, b: sum(osem.glsc3(p, q, m['geo']['bm1']) for p, q in zip(a, b))
    lhs, rhs, wrong = dot(Av, w0), dot(v0, Atw), dot(v0, Aw)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs), (lhs, rhs)
    assert abs(lhs - wrong) >= 1000 * abs(lhs - rhs), (lhs, rhs, wrong)


@pytest.mark.parametrize('adjoint', [False, True])
def test_stored_orbit_of_a_time_periodic_base_flow(adjoint):
    """ns_steps(orbit=...): every step linearises about its own base flow (core/linear_operators.f90:254-275).  An
    orbit that stands still is the steady stepper bit for bit; the map stays linear in the perturbation; and the
    result depends on the orbit only through the steps actually taken (changing the column of a later step than the
    horizon's last does nothing, changing the first one does)."""
    N = 5
    m = setup(2, (2, 2), N, deform=0.03)
    x, y = m['coords']
    base = [1.0 + 0.3 * np.sin(np.pi * y), 0.4 * np.cos(np.pi * x)]
    rng = np.random.default_rng(11)
    vmult = 1.0 / osem.multiplicity(m['glo'])
    field = lambda: osem.dssum(rng.standard_normal(x.shape), m['glo']) * vmult * m['mask']
    v0, w0, p0 = [field(), field()], [field(), field()], 0 * m['ps']['bm2']
    nu, dt, nst = 0.05, 2e-3, 3
    run = lambda q, orb: ons.ns_steps(m['glo'], m['mask'], m['geo'], N, m['ps'], m['dl'], base, q, p0, nu, dt, nst,
                                      mean_free=False, adjoint=adjoint, orbit=orb)[0]
    still = run(v0, [base] * nst)
    steady = run(v0, None)
    assert all(np.array_equal(a, b) for a, b in zip(still, steady))
    orbit = [[(1.0 + 0.2 * s) * base[0], base[1] + 0.1 * s * x] for s in range(nst)]
    a, b = run(v0, orbit), run(w0, orbit)
    c = run([2.0 * p - 0.5 * q for p, q in zip(v0, w0)], orbit)
    scale = max(np.max(np.abs(f)) for f in a)
    assert max(np.max(np.abs(f - (2.0 * p - 0.5 * q))) for f, p, q in zip(c, a, b)) <= 1e-9 * scale
    assert max(np.max(np.abs(p - q)) for p, q in zip(a, steady)) > 1e-5 * scale
    longer = run(v0, orbit + [[0 * x, 0 * x]])
    assert all(np.array_equal(p, q) for p, q in zip(a, longer))
    other = run(v0, [[0 * x, 0 * x]] + orbit[1:])
    assert max(np.max(np.abs(p - q)) for p, q in zip(a, other)) > 1e-5 * scale
