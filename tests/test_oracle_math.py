"""The oracle restatements that have no reference fixture to pin them ([UPSTREAM-RECALL] pieces) checked
against independent mathematics on the CPU: Gauss-Legendre dealiasing operators, EXT/BDF coefficients,
the LightKrylov-style eigs / svds drivers against numpy's dense solvers."""
import numpy as np
import pytest

from oracle import krylov as okr
from oracle import sem as osem


def test_gauss_legendre_and_interpolation_operators():
    for lxd in (6, 9, 12):
        z, w = osem.gl(lxd)
        for p in range(2 * lxd):                       # exact to degree 2 lxd - 1
            exact = 0.0 if p % 2 else 2.0 / (p + 1)
            assert abs(np.sum(w * z ** p) - exact) < 1e-13
        D = osem.deriv_matrix(z)
        for p in range(lxd):                           # derivative matrix exact on P_{lxd-1}
            assert np.max(np.abs(D @ z ** p - (p * z ** (p - 1) if p else 0 * z))) < 1e-10
    zg, _ = osem.gll(7)
    zd, _ = osem.gl(12)
    J = osem.interp_matrix(zg, zd)
    for p in range(8):                                 # interpolation reproduces P_7
        assert np.max(np.abs(J @ zg ** p - zd ** p)) < 1e-13
    assert np.allclose(osem.interp_matrix(zg, zg), np.eye(8))


def test_dealiased_convection_equals_the_weak_form_integral():
    """sum_p [J^T (c . grad)(J u)]_p = integral of c . grad u, here for polynomials the 3/2 rule integrates
    exactly (3-D and 2-D, affine elements), and v^T [..] = integral of v c . grad u for a test function v."""
    N = 7
    x, y, z, _ = osem.box_mesh(2, 1, 2, N)
    geo = osem.geometry(N, x, y, z)
    dl = osem.dealias_setup(N, 12, geo['rst'])
    cf = osem.set_convect([x ** 3, y * x, 1.0 + z ** 2], dl)
    r = osem.convect_dealiased(x ** 4 * y, cf, dl)
    assert abs(r.sum() - (4 / 7 / 2 + 1 / 12)) < 1e-13
    v = x * z + y ** 2                                  # integral of v * (4 x^6 y + x^5 y) over the unit cube
    exact = (4 / 8) * (1 / 2) * (1 / 2) + (1 / 7) * (1 / 2) * (1 / 2) + (4 / 7) * (1 / 4) + (1 / 6) * (1 / 4)
    assert abs(np.sum(v * r) - exact) < 1e-13
    x2, y2, _ = osem.box_mesh_2d(2, 3, 5)
    g2 = osem.geometry(5, x2, y2)
    dl2 = osem.dealias_setup(5, 9, g2['rst'])
    r2 = osem.convect_dealiased(x2 ** 3 * y2, osem.set_convect([y2 ** 2, x2 + 1.0], dl2), dl2)
    # y^2 * 3 x^2 y + (x + 1) x^3 -> 3 * (1/3) * (1/4) + 1/5 + 1/4
    assert abs(r2.sum() - (0.25 + 0.2 + 0.25)) < 1e-13


def test_bdf_ext_coefficients_and_update():
    # Nek's convention: bd[0] u^{n+1} = sum_i bd[i] u^{n+1-i} + dt f ; consistency and order conditions
    for o, bd in ((1, [1.0, 1.0]), (2, [1.5, 2.0, -0.5]), (3, [11 / 6, 3.0, -1.5, 1 / 3])):
        i = np.arange(1, o + 1)
        b = np.array(bd[1:])
        assert abs(bd[0] - b.sum()) < 1e-15                       # constants are reproduced
        # u(t) = t^p sampled at t = 0 (new), -1, -2, ..: bd0 * 0 - sum_i b_i (-i)^p = derivative at 0 * dt
        for p in range(1, o + 1):
            lhs = bd[0] * 0.0 - np.sum(b * (-i.astype(float)) ** p)
            assert abs(lhs - (1.0 if p == 1 else 0.0)) < 1e-14
    for o, ab in ((1, [1.0, 0, 0]), (2, [2.0, -1.0, 0]), (3, [3.0, -3.0, 1.0])):
        i = np.arange(1, 4)
        for p in range(o):                                        # extrapolation to t = 0 exact on P_{o-1}
            assert abs(np.sum(np.array(ab) * (-i.astype(float)) ** p) - (1.0 if p == 0 else 0.0)) < 1e-14
    rng = np.random.default_rng(0)
    bf, e1, e2, v0, v1, v2, bm1 = (rng.standard_normal(50) for _ in range(7))
    bf0, e10, e20 = bf.copy(), e1.copy(), e2.copy()
    ab, bd = [3.0, -3.0, 1.0], [11 / 6, 3.0, -1.5, 1 / 3]
    osem.bdf_ext(bf, e1, e2, [v0, v1, v2], bm1, ab, bd, 7.0)
    assert np.array_equal(e2, e10) and np.array_equal(e1, bf0)
    assert np.allclose(bf, 3 * bf0 - 3 * e10 + e20 + 7.0 * bm1 * (3 * v0 - 1.5 * v1 + v2 / 3), rtol=1e-14)


def _dense_problem(n, seed, sym=False):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) / np.sqrt(n)
    if sym:
        A = 0.5 * (A + A.T)
    w = rng.random(n) + 0.5                                      # inner-product weight
    c = okr.Ctx(bm1s=w, in_dot=[True], time_in_dot=False)
    return A, w, c, okr.KVec([rng.standard_normal(n)], 0.0)


def test_oracle_eigs_against_dense_eigenvalues():
    n = 60
    A, w, c, q0 = _dense_problem(n, 4)
    okr.k_normalize(c, q0)
    vals, vecs, res, k, H = okr.eigs(c, lambda q: okr.KVec([A @ q.f[0]], q.time), q0, n, nev=3, tol=1e-10)
    ev = np.linalg.eigvals(A)
    for i in np.where(res < 1e-10)[0]:
        assert np.min(np.abs(ev - vals[i])) < 1e-8
    assert np.count_nonzero(res < 1e-10) >= 3


def test_oracle_svds_against_dense_svd_in_the_weighted_inner_product():
    """svds with rmatvec = the adjoint in the W inner product: singular values of W^1/2 A W^-1/2."""
    n = 50
    A, w, c, u0 = _dense_problem(n, 7)
    okr.k_normalize(c, u0)
    adj = (A.T * w[None, :]) / w[:, None]                        # W^-1 A^T W
    sig, uv, vv, res, k, B = okr.svds(c, lambda q: okr.KVec([A @ q.f[0]], q.time),
                                      lambda q: okr.KVec([adj @ q.f[0]], q.time), u0, n, nev=3, tol=1e-10)
    ref = np.linalg.svd(np.sqrt(w)[:, None] * A / np.sqrt(w)[None, :], compute_uv=False)
    conv = np.sort(sig[res < 1e-10])[::-1]
    assert len(conv) >= 3 and np.allclose(conv[:3], ref[:3], rtol=1e-9)


def test_adjoint_stepper_oracle_is_the_discrete_adjoint():
    """Pins the oracle's adjoint pieces independently of the GPU: convect_dealiased_t is the exact transpose of
    convect_dealiased, and scalar_steps_adjoint satisfies <A u, v>_B = <u, A^+ v>_B (continuous, masked u, v) for
    the BDF/EXT stepper with its order ramp, in 3-D and 2-D."""
    from oracle import sem as osem
    rng = np.random.default_rng(4)
    for dim in (3, 2):
        N = 3
        if dim == 3:
            x, y, z, glo = osem.box_mesh(2, 2, 2, N, deform=0.04)
            x0, y0, z0, _ = osem.box_mesh(2, 2, 2, N)
            mask = osem.boundary_mask_box(None, x0, y0, z0)
            geo = osem.geometry(N, x, y, z)
            vel = [np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), 0.3 * np.sin(np.pi * z)]
        else:
            x, y, glo = osem.box_mesh_2d(3, 2, N, deform=0.04)
            x0, y0, _ = osem.box_mesh_2d(3, 2, N)
            mask = osem.boundary_mask_box(None, x0, y0, None, lengths=(1.0, 1.0))
            geo = osem.geometry(N, x, y)
            vel = [1.0 + 0.3 * np.sin(np.pi * y), 0.4 * np.cos(np.pi * x)]
        dl = osem.dealias_setup(N, 3 * (N + 1) // 2, geo['rst'])
        cf = osem.set_convect(vel, dl)
        a, b = rng.standard_normal(x.shape), rng.standard_normal(x.shape)
        lhs, rhs = np.sum(b * osem.convect_dealiased(a, cf, dl)), np.sum(a * osem.convect_dealiased_t(b, cf, dl))
        assert abs(lhs - rhs) <= 1e-12 * np.sqrt(np.sum(a * a) * np.sum(b * b)) * np.max(np.abs(cf[0]))
        vm = 1.0 / osem.multiplicity(glo)
        u = osem.dssum(rng.standard_normal(x.shape), glo) * vm * mask
        v = osem.dssum(rng.standard_normal(x.shape), glo) * vm * mask
        for nsteps in (1, 2, 4):
            Au = osem.scalar_steps(glo, mask, geo, N, cf, dl, u, 0.05, 5e-3, nsteps)
            Atv = osem.scalar_steps_adjoint(glo, mask, geo, N, cf, dl, v, 0.05, 5e-3, nsteps)
            l, r = np.sum(geo['bm1'] * Au * v), np.sum(geo['bm1'] * u * Atv)
            assert abs(l - r) <= 1e-10 * np.sqrt(np.sum(geo['bm1'] * u * u) * np.sum(geo['bm1'] * v * v)), (dim, nsteps, l, r)


@pytest.mark.parametrize('order', [2, 4])
def test_finite_difference_frechet_map(order):
    """forward_finite_difference_map (core/matvec.f90:246-379): for F(x) = A x + 0.3 x^3 the differences reproduce
    the Jacobian A q + 0.9 X^2 q up to the truncation error (eps0^2 and eps0^4, both below rounding here) and the
    rounding noise 1e-16 |F| / eps0; a linear map is reproduced whatever the base state."""
    from oracle import krylov as okr
    rng = np.random.default_rng(5 + order)
    n = 200
    A = rng.standard_normal((n, n)) / np.sqrt(n)
    c = okr.Ctx(bm1s=rng.uniform(0.5, 1.5, n), in_dot=[True], time_in_dot=True)
    X = okr.KVec([rng.standard_normal(n)], 0.7)
    q = okr.KVec([rng.standard_normal(n)], -0.2)
    F = lambda x: okr.KVec([A @ x.f[0] + 0.3 * x.f[0] ** 3], 2.0 * x.time)
    f = okr.forward_finite_difference_map(c, F, X, q, order)
    exact = A @ q.f[0] + 0.9 * X.f[0] ** 2 * q.f[0]
    assert np.max(np.abs(f.f[0] - exact)) <= 1e-8 * np.max(np.abs(exact))
    assert abs(f.time - 2.0 * q.time) <= 1e-8
    L = lambda x: okr.KVec([A @ x.f[0]], x.time)
    g = okr.forward_finite_difference_map(c, L, X, q, order)
    assert np.max(np.abs(g.f[0] - A @ q.f[0])) <= 1e-8 * np.max(np.abs(A @ q.f[0]))
    with pytest.raises(ValueError):
        okr.forward_finite_difference_map(c, F, X, q, 3)


@pytest.mark.parametrize('dim', [2, 3])
def test_norm_grad_integrates_polynomial_gradients_exactly(dim):
    """norm_grad (core/utils.f90:446-486) on a deformed mesh: for polynomial fields the collocation derivatives are
    the exact ones wherever the mapping is affine, and on the unit box
    int (d(x^2 y)/dx)^2 + (d(x^2 y)/dy)^2 = 4/9 + 1/5; with the smooth deformation the GLL rule converges to it."""
    N = 7
    if dim == 2:
        x, y, _ = osem.box_mesh_2d(2, 3, N, deform=0.0)
        coords, u, v = (x, y), x * x * y, 0 * x
        exact = 4.0 / 9.0 + 1.0 / 5.0
    else:
        x, y, z, _ = osem.box_mesh(2, 2, 2, N, deform=0.0)
        coords, u, v = (x, y, z), x * x * y * z, z ** 3
        # |grad(x^2 y z)|^2 + |grad z^3|^2 over the unit cube: 4/27 + 1/15 + 1/15 + 9/5
        exact = 4.0 / 27.0 + 2.0 / 15.0 + 9.0 / 5.0
    geo = osem.geometry(N, *coords)
    vel = [u, v] + ([0 * u] if dim == 3 else [])
    assert abs(osem.norm_grad(vel, geo, N, geo['bm1']) - exact) <= 1e-12
    # a sponge-like weight (bm1s zeroed where x > 0.5) only counts the rest of the domain
    half = osem.norm_grad([u] + [0 * u] * (dim - 1), geo, N, geo['bm1'] * (coords[0] <= 0.5 + 1e-12))
    full = osem.norm_grad([u] + [0 * u] * (dim - 1), geo, N, geo['bm1'])
    assert 0.0 < half < full
    # deformed elements: the same integral up to the quadrature error of the mapped integrand
    if dim == 2:
        xd, yd, _ = osem.box_mesh_2d(2, 3, N, deform=0.03)
        geod = osem.geometry(N, xd, yd)
        assert abs(osem.norm_grad([xd * xd * yd, 0 * xd], geod, N, geod['bm1']) - exact) <= 1e-8


@pytest.mark.parametrize('dim', [2, 3])
def test_compute_cfl_of_uniform_flows_on_affine_elements(dim):
    """compute_cfl on a box of equal elements of size h: u.grad r = 2 u / h, so a uniform flow U gives
    dt sum_a (2 |U_a| / h_a) max_i(1 / dr_i) with the largest 1 / dr at the end points, 1 / (z_2 - z_1); linear in dt and
    in the velocity; a flow along one axis only sees that axis' spacing."""
    N = 6
    z, _ = osem.gll(N)
    end = 1.0 / (z[1] - z[0])
    if dim == 2:
        x, y, _ = osem.box_mesh_2d(4, 2, N, deform=0.0)
        coords, h, U = (x, y), (0.25, 0.5), (1.5, -0.7)
    else:
        x, y, z3, _ = osem.box_mesh(2, 4, 1, N, deform=0.0)
        coords, h, U = (x, y, z3), (0.5, 0.25, 1.0), (0.3, -1.1, 2.0)
    geo = osem.geometry(N, *coords)
    vel = [u + 0 * coords[0] for u in U]
    dt = 0.01
    exact = dt * sum(2.0 * abs(u) / ha for u, ha in zip(U, h)) * end
    assert abs(osem.compute_cfl(vel, geo, N, dt) - exact) <= 1e-12 * exact
    assert abs(osem.compute_cfl(vel, geo, N, 3 * dt) - 3 * exact) <= 1e-12 * exact
    one = [vel[0]] + [0 * v for v in vel[1:]]
    assert abs(osem.compute_cfl(one, geo, N, dt) - dt * 2.0 * abs(U[0]) / h[0] * end) <= 1e-12 * exact
    # a flow that vanishes at the element ends and peaks inside meets the centred spacing there
    bump = [np.where(np.isclose(np.abs(np.mod(coords[0] / h[0], 1.0) - 0.5), 0.5), 0.0, 1.0)] + [0 * v for v in vel[1:]]
    inner = 1.0 / (0.5 * (z[2] - z[0]))
    assert abs(osem.compute_cfl(bump, geo, N, dt) - dt * 2.0 / h[0] * inner) <= 1e-12
