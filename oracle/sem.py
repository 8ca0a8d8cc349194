"""Spectral-element pieces of the oracle (numpy, fp64).  TEST INFRASTRUCTURE ONLY.

Restates the Nek5000 routines the reference reaches through ``nek_advance``
(reference call site ``core/linear_operators.f90:247``).
Nek5000 itself is not vendored in /root/reference, so the formulas follow the
published algorithm ([UPSTREAM-RECALL]: ``speclib.f`` zwgll/dgll, ``coef.f``
geom1/glmapm1/geodat1, ``hmholtz.f`` axhelm, ``navier5.f`` local_grad3,
``math.f`` glsc3, gslib gs_op(add) behind ``dssum``).

Array layout is Nek's element-local one: ``u[e, k, j, i]`` with ``i`` fastest
(C order of shape (nel, lz, ly, lx)), shared nodes duplicated.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# GLL quadrature and derivative matrix  ([UPSTREAM-RECALL] speclib.f zwgll/dgll)
# ----------------------------------------------------------------------------
def _legendre(n: int, x: np.ndarray):
    """P_n(x) and P_{n-1}(x) by the three-term recurrence."""
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for m in range(1, n):
        p0, p1 = p1, ((2 * m + 1) * x * p1 - m * p0) / (m + 1)
    return p1, p0


def gll(n: int):
    """Nodes z[0..n] (ascending, in [-1,1]) and weights w of the (n+1)-point GLL rule.

    Interior nodes are the roots of P_n'(x); w_i = 2 / (n (n+1) P_n(z_i)^2).
    """
    if n < 1:
        raise ValueError("polynomial order must be >= 1")
    x = -np.cos(np.pi * np.arange(n + 1) / n)  # Chebyshev-Lobatto start
    for _ in range(100):
        pn, pnm1 = _legendre(n, x)
        # q(x) = (1-x^2) P_n'(x) = n (P_{n-1} - x P_n);  q'(x) = -n (n+1) P_n
        q = n * (pnm1 - x * pn)
        dq = -n * (n + 1) * pn
        dx = q / dq
        dx[0] = dx[-1] = 0.0
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # enforce antisymmetry
    pn, _ = _legendre(n, x)
    w = 2.0 / (n * (n + 1) * pn * pn)
    return x, w


def dgll(n: int):
    """GLL derivative matrix D[i, j] = dl_j/dx (z_i):  (du/dr)_i = sum_j D[i,j] u_j."""
    z, _ = gll(n)
    pn, _ = _legendre(n, z)
    d = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        for j in range(n + 1):
            if i != j:
                d[i, j] = pn[i] / (pn[j] * (z[i] - z[j]))
    d[0, 0] = -n * (n + 1) / 4.0
    d[n, n] = n * (n + 1) / 4.0
    return d


# ----------------------------------------------------------------------------
# Meshes
# ----------------------------------------------------------------------------
def box_mesh(nelx, nely, nelz, n, deform=0.0, lengths=(1.0, 1.0, 1.0)):
    """Structured box of nelx*nely*nelz hexahedra of order n on [0,L]^3.

    Returns x, y, z of shape (nel, lx, lx, lx) (element order: ex fastest) and the
    0-based lexicographic unique-node numbering ``glo`` (int64, same shape).
    ``deform`` adds the smooth map of SURVEY.md section 8d (variant B).
    """
    lx = n + 1
    zg, _ = gll(n)
    r = 0.5 * (zg + 1.0)

    def line(nel, length):
        h = length / nel
        return (np.arange(nel)[:, None] + r[None, :]) * h  # (nel, lx)

    xl, yl, zl = line(nelx, lengths[0]), line(nely, lengths[1]), line(nelz, lengths[2])
    nel = nelx * nely * nelz
    x = np.empty((nelz, nely, nelx, lx, lx, lx))
    y = np.empty_like(x)
    z = np.empty_like(x)
    x[:] = xl[None, None, :, None, None, :]
    y[:] = yl[None, :, None, None, :, None]
    z[:] = zl[:, None, None, :, None, None]
    gi = (np.arange(nelx)[:, None] * n + np.arange(lx)[None, :]).astype(np.int64)
    gj = (np.arange(nely)[:, None] * n + np.arange(lx)[None, :]).astype(np.int64)
    gk = (np.arange(nelz)[:, None] * n + np.arange(lx)[None, :]).astype(np.int64)
    nxg, nyg = nelx * n + 1, nely * n + 1
    glo = np.empty(x.shape, dtype=np.int64)
    glo[:] = (gk[:, None, None, :, None, None] * nyg + gj[None, :, None, None, :, None]) * nxg \
        + gi[None, None, :, None, None, :]
    x = x.reshape(nel, lx, lx, lx)
    y = y.reshape(nel, lx, lx, lx)
    z = z.reshape(nel, lx, lx, lx)
    glo = glo.reshape(nel, lx, lx, lx)
    if deform != 0.0:
        bump = deform * np.sin(np.pi * x / lengths[0]) * np.sin(np.pi * y / lengths[1]) \
            * np.sin(np.pi * z / lengths[2])
        x, y, z = x + bump, y + bump, z + bump
    return x, y, z, glo


def box_mesh_2d(nelx, nely, n, deform=0.0, lengths=(1.0, 1.0)):
    lx = n + 1
    zg, _ = gll(n)
    r = 0.5 * (zg + 1.0)
    xl = (np.arange(nelx)[:, None] + r[None, :]) * (lengths[0] / nelx)
    yl = (np.arange(nely)[:, None] + r[None, :]) * (lengths[1] / nely)
    x = np.empty((nely, nelx, lx, lx))
    y = np.empty_like(x)
    x[:] = xl[None, :, None, :]
    y[:] = yl[:, None, :, None]
    gi = (np.arange(nelx)[:, None] * n + np.arange(lx)[None, :]).astype(np.int64)
    gj = (np.arange(nely)[:, None] * n + np.arange(lx)[None, :]).astype(np.int64)
    glo = np.empty(x.shape, dtype=np.int64)
    glo[:] = gj[:, None, :, None] * (nelx * n + 1) + gi[None, :, None, :]
    nel = nelx * nely
    x, y, glo = x.reshape(nel, lx, lx), y.reshape(nel, lx, lx), glo.reshape(nel, lx, lx)
    if deform != 0.0:
        bump = deform * np.sin(np.pi * x / lengths[0]) * np.sin(np.pi * y / lengths[1])
        x, y = x + bump, y + bump
    return x, y, glo


def glo_num_from_coords(coords, tol=1e-8):
    """Global numbering by coordinate matching (for the Nek field-file fixtures).

    ``coords`` is a tuple of arrays of identical shape; points closer than ``tol``
    (relative to the domain extent) share an id.  Returns int64 ids, 0-based.
    """
    shape = coords[0].shape
    pts = np.stack([c.ravel() for c in coords], axis=1)
    span = np.maximum(pts.max(axis=0) - pts.min(axis=0), 1e-300)
    q = np.round((pts - pts.min(axis=0)) / (span * tol)).astype(np.int64)
    _, inv = np.unique(q, axis=0, return_inverse=True)
    return inv.reshape(shape).astype(np.int64)


def boundary_mask_box(glo_shape_mesh, x, y, z=None, lengths=(1.0, 1.0, 1.0), tol=1e-12):
    """Homogeneous-Dirichlet mask: 0 on the box boundary, 1 inside (undeformed coords)."""
    m = np.ones_like(x)
    for c, length in zip((x, y, z), lengths):
        if c is None:
            continue
        m[(np.abs(c) < tol) | (np.abs(c - length) < tol)] = 0.0
    return m


# ----------------------------------------------------------------------------
# Geometry  ([UPSTREAM-RECALL] coef.f glmapm1 / geodat1)
# ----------------------------------------------------------------------------
def grad_rst(u, d):
    """local_grad3 / local_grad2: reference-space derivatives of an element-local field."""
    if u.ndim == 4:  # (e, k, j, i)
        ur = np.einsum('il,ekjl->ekji', d, u)
        us = np.einsum('jl,ekli->ekji', d, u)
        ut = np.einsum('kl,elji->ekji', d, u)
        return ur, us, ut
    ur = np.einsum('il,ejl->eji', d, u)
    us = np.einsum('jl,eli->eji', d, u)
    return ur, us


def geometry(n, x, y, z=None):
    """Geometric factors from nodal coordinates.

    3-D returns dict with jac, bm1, g[6] = (G1..G6), and rx..tz (each already times jac
    as in Nek).  2-D returns g[3] = (G1, G2, G4).
    """
    d = dgll(n)
    _, w = gll(n)
    if z is not None:
        xr, xs, xt = grad_rst(x, d)
        yr, ys, yt = grad_rst(y, d)
        zr, zs, zt = grad_rst(z, d)
        jac = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr)
        rx = ys * zt - yt * zs
        ry = xt * zs - xs * zt
        rz = xs * yt - xt * ys
        sx = yt * zr - yr * zt
        sy = xr * zt - xt * zr
        sz = xt * yr - xr * yt
        tx = yr * zs - ys * zr
        ty = xs * zr - xr * zs
        tz = xr * ys - xs * yr
        w3 = w[:, None, None] * w[None, :, None] * w[None, None, :]
        sc = w3[None] / jac
        g = np.stack([
            (rx * rx + ry * ry + rz * rz) * sc,
            (sx * sx + sy * sy + sz * sz) * sc,
            (tx * tx + ty * ty + tz * tz) * sc,
            (rx * sx + ry * sy + rz * sz) * sc,
            (rx * tx + ry * ty + rz * tz) * sc,
            (sx * tx + sy * ty + sz * tz) * sc,
        ])
        return dict(jac=jac, bm1=jac * w3[None], g=g,
                    rst=(rx, ry, rz, sx, sy, sz, tx, ty, tz))
    xr, xs = grad_rst(x, d)
    yr, ys = grad_rst(y, d)
    jac = xr * ys - xs * yr
    rx, ry, sx, sy = ys, -xs, -yr, xr
    w2 = w[:, None] * w[None, :]
    sc = w2[None] / jac
    g = np.stack([(rx * rx + ry * ry) * sc, (sx * sx + sy * sy) * sc, (rx * sx + ry * sy) * sc])
    return dict(jac=jac, bm1=jac * w2[None], g=g, rst=(rx, ry, sx, sy))


# ----------------------------------------------------------------------------
# Operator kernels
# ----------------------------------------------------------------------------
def axhelm(u, g, d, h1=1.0, h2=0.0, bm1=None):
    """Element-local Helmholtz operator w = h1 * D^T (G (D u)) + h2 * bm1 * u.

    ([UPSTREAM-RECALL] hmholtz.f axhelm, general/deformed branch.)  No dssum, no mask.
    """
    if u.ndim == 4:
        ur, us, ut = grad_rst(u, d)
        wr = h1 * (g[0] * ur + g[3] * us + g[4] * ut)
        ws = h1 * (g[1] * us + g[3] * ur + g[5] * ut)
        wt = h1 * (g[2] * ut + g[4] * ur + g[5] * us)
        w = np.einsum('li,ekjl->ekji', d, wr) + np.einsum('lj,ekli->ekji', d, ws) \
            + np.einsum('lk,elji->ekji', d, wt)
    else:
        ur, us = grad_rst(u, d)
        wr = h1 * (g[0] * ur + g[2] * us)
        ws = h1 * (g[1] * us + g[2] * ur)
        w = np.einsum('li,ejl->eji', d, wr) + np.einsum('lj,eli->eji', d, ws)
    if h2 != 0.0:
        w = w + h2 * bm1 * u
    return w


def dssum(u, glo):
    """Direct-stiffness summation: every copy of a global node gets the sum of all copies."""
    flat = glo.ravel()
    nglob = int(flat.max()) + 1
    acc = np.bincount(flat, weights=u.ravel(), minlength=nglob)
    return acc[flat].reshape(u.shape)


def multiplicity(glo):
    flat = glo.ravel()
    cnt = np.bincount(flat)
    return cnt[flat].reshape(glo.shape).astype(np.float64)


def ax(u, g, d, glo, mask, h1=1.0, h2=0.0, bm1=None):
    """Nek's ax(w,x,h1,h2,n): axhelm + dssum + col2(mask)."""
    return dssum(axhelm(u, g, d, h1, h2, bm1), glo) * mask


def glsc3(a, b, mult):
    """sum_i a_i b_i mult_i  ([UPSTREAM-RECALL] math.f glsc3, single rank: no gop)."""
    return float(np.sum(a.ravel() * b.ravel() * mult.ravel()))


def gradm1(u, geo, d):
    """[UPSTREAM-RECALL] navier5.f gradm1: physical-space collocation derivatives on the GLL points,
    du/dx_b = sum_a (rst[a][b] / jac) du/dr_a  (rst already holds one factor jac, `jacmi` removes it)."""
    g = grad_rst(u, d)
    dim = len(g)
    return [sum(geo['rst'][a * dim + b] * g[a] for a in range(dim)) / geo['jac'] for b in range(dim)]


def norm_grad(vel, geo, n, bm1s):
    """core/utils.f90:446-486: sum_c sum_b glsc3(du_c/dx_b, bm1s, du_c/dx_b) -- no square root; the number outpost_ks
    compares with 1.1 to drop spurious Ritz vectors (core/eigensolvers.f90:587-594).  2-D: the four terms of :470-471,
    3-D: all nine (:473-479)."""
    d = dgll(n)
    norma = 0.0
    for u in vel:
        for du in gradm1(u, geo, d):
            norma += glsc3(du, du, bm1s)
    return norma


def compute_cfl(vel, geo, n, dt):
    """[UPSTREAM-RECALL] Nek5000 compute_cfl + getdr (call sites core/linear_stab.f90:222,231): the maximum over the
    GLL points of dt (|u.grad r| dri_i + |u.grad s| dsi_j [+ |u.grad t| dti_k]), u.grad r = (u rx + v ry + w rz) / jac
    (rst holds the metrics times jac), dr_i = z_2 - z_1 at the ends and (z_i+1 - z_i-1) / 2 inside."""
    z, _ = gll(n)
    dr = np.empty(n + 1)
    dr[0], dr[n] = z[1] - z[0], z[n] - z[n - 1]
    dr[1:n] = 0.5 * (z[2:] - z[:-2])
    dri = 1.0 / dr
    dim = len(vel)
    total = 0.0
    for a in range(dim):
        ua = sum(vel[b] * geo['rst'][a * dim + b] for b in range(dim)) / geo['jac']
        shape = [1] * ua.ndim
        shape[ua.ndim - 1 - a] = n + 1                             # axis of direction a: i fastest (last axis)
        total = total + np.abs(dt * ua * dri.reshape(shape))
    return float(np.max(total))


def convect(u, vel, rst, d):
    """Pointwise (times jac) convective derivative  jac * (U . grad) u  via local_grad3."""
    if u.ndim == 4:
        ur, us, ut = grad_rst(u, d)
        rx, ry, rz, sx, sy, sz, tx, ty, tz = rst
        return (vel[0] * (rx * ur + sx * us + tx * ut) + vel[1] * (ry * ur + sy * us + ty * ut)
                + vel[2] * (rz * ur + sz * us + tz * ut))
    ur, us = grad_rst(u, d)
    rx, ry, sx, sy = rst
    return vel[0] * (rx * ur + sx * us) + vel[1] * (ry * ur + sy * us)


# ----------------------------------------------------------------------------
# Helmholtz solve  ([UPSTREAM-RECALL] hmholtz.f cggo + setprec)
# ----------------------------------------------------------------------------
def helm_diag(g, d, bm1, h1, h2):
    """Diagonal of h1 A + h2 B per local point (setprec without the deformed-boundary cross terms)."""
    d2 = d * d                       # d2[q, i] = D(q,i)^2
    if g.shape[0] == 6:
        s = np.einsum('qi,ekjq->ekji', d2, g[0]) + np.einsum('qj,ekqi->ekji', d2, g[1]) \
            + np.einsum('qk,eqji->ekji', d2, g[2])
    else:
        s = np.einsum('qi,ejq->eji', d2, g[0]) + np.einsum('qj,eqi->eji', d2, g[1])
    return h1 * s + h2 * bm1


def cggo(rhs, g, d, glo, mask, bm1, h1, h2, tol=1e-10, maxit=500):
    """Jacobi-PCG of Nek's cggo: z = D r ; rtz = (r,z)_mult ; p = z + beta p ; w = mask dssum axhelm p ;
    alpha = rtz / (w,p)_mult ; x += alpha p ; r -= alpha w.  Returns (x, iterations, residual drop)."""
    mult = 1.0 / multiplicity(glo)
    dinv = mask / dssum(helm_diag(g, d, bm1, h1, h2), glo)
    x = np.zeros_like(rhs)
    p = np.zeros_like(rhs)
    r = rhs * mask
    z = dinv * r
    rtz1, rtz2 = float(np.sum(r * z * mult)), 1.0
    r0, rn, it = -1.0, 0.0, 0
    for it in range(1, maxit + 1):
        beta = 0.0 if it == 1 else rtz1 / rtz2
        p = z + beta * p
        w = dssum(axhelm(p, g, d, h1, h2, bm1), glo) * mask
        rho = float(np.sum(w * p * mult))
        if not rho > 0.0:
            break
        alpha = rtz1 / rho
        x = x + alpha * p
        r = r - alpha * w
        z = dinv * r
        rtz2, rtz1 = rtz1, float(np.sum(r * z * mult))
        rn = np.sqrt(abs(rtz1))
        if r0 < 0.0:
            r0 = np.sqrt(abs(rtz2))
        if rn <= tol * r0:
            break
    return x, it, (rn / r0 if r0 > 0 else 0.0)


# ----------------------------------------------------------------------------
# Dealiased convection  ([UPSTREAM-RECALL] Nek5000 convect.f: set_dealias_rx, set_convect_new,
# convect_new, intp_rstd, grad_rst on the lxd Gauss-Legendre mesh.  Pinned, together with axhelm and the geometry, by
# the steady momentum balance of the reference's own base flows: tests/test_oracle_fixtures.py, oracle/ns.py header)
# ----------------------------------------------------------------------------
def gl(n: int):
    """n Gauss-Legendre nodes and weights on [-1, 1] (Nek zwgl)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w


def _bary_weights(z):
    n = len(z)
    w = np.ones(n)
    for i in range(n):
        for j in range(n):
            if i != j:
                w[i] /= (z[i] - z[j])
    return w


def interp_matrix(zfrom, zto):
    """J[I, i] = l_i(zto_I), the Lagrange interpolant through zfrom evaluated at zto (Nek igllm)."""
    bw = _bary_weights(zfrom)
    J = np.zeros((len(zto), len(zfrom)))
    for I, x in enumerate(zto):
        d = x - zfrom
        hit = np.where(np.abs(d) < 1e-15)[0]
        if hit.size:
            J[I, hit[0]] = 1.0
        else:
            t = bw / d
            J[I] = t / t.sum()
    return J


def deriv_matrix(z):
    """D[i, j] = l_j'(z_i) on arbitrary nodes (Nek gen_dgl on the Gauss points)."""
    n = len(z)
    bw = _bary_weights(z)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = (bw[j] / bw[i]) / (z[i] - z[j])
        D[i, i] = -np.sum(D[i])
    return D


def interp_fine(u, J):
    """intp_rstd(.., idir = 0): element-local field on lx1^d GLL points -> lxd^d Gauss points."""
    if u.ndim == 4:
        return np.einsum('Kk,Jj,Ii,ekji->eKJI', J, J, J, u, optimize=True)
    return np.einsum('Jj,Ii,eji->eJI', J, J, u, optimize=True)


def project_coarse(uf, J):
    """intp_rstd(.., idir = 1): the transpose of interp_fine."""
    if uf.ndim == 4:
        return np.einsum('Kk,Jj,Ii,eKJI->ekji', J, J, J, uf, optimize=True)
    return np.einsum('Jj,Ii,eJI->eji', J, J, uf, optimize=True)


def dealias_setup(n, lxd, rst):
    """set_dealias_rx: the metrics rxm1.. (already times the Jacobian) interpolated to the fine mesh and
    multiplied by the Gauss weights, plus the interpolation and fine-mesh derivative matrices."""
    zg, _ = gll(n)
    zd, wd = gl(lxd)
    J = interp_matrix(zg, zd)
    Dg = deriv_matrix(zd)
    dim = 3 if len(rst) == 9 else 2
    w = wd[:, None, None] * wd[None, :, None] * wd[None, None, :] if dim == 3 else wd[:, None] * wd[None, :]
    rxf = [interp_fine(m, J) * w[None] for m in rst]
    return dict(J=J, Dg=Dg, rxf=rxf, lxd=lxd, dim=dim)


def set_convect(vel, dl):
    """set_convect_new: contravariant convecting field on the fine mesh, (c_r, c_s[, c_t])."""
    f = [interp_fine(v, dl['J']) for v in vel]
    d, rxf = dl['dim'], dl['rxf']
    return [sum(rxf[a * d + b] * f[b] for b in range(d)) for a in range(d)]


def convect_dealiased(u, cf, dl):
    """convect_new(bdu, u, .false., cr, cs, ct, .true.): J^T [ (c . grad_rst)(J u) ] -- the mass matrix and
    the Jacobian are inside c."""
    uf = interp_fine(u, dl['J'])
    g = grad_rst(uf, dl['Dg'])
    return project_coarse(sum(c * gi for c, gi in zip(cf, g)), dl['J'])


# ----------------------------------------------------------------------------
# EXT / BDF sums of the perturbation step  ([UPSTREAM-RECALL] Nek5000 perturb.f makextp, makebdfp)
# ----------------------------------------------------------------------------
def bdf_ext(bf, e1, e2, vlag, bm1, ab, bd, rho_over_dt):
    """In place:  ta = ab1 e1 + ab2 e2 ; e2 <- e1 ; e1 <- bf ; bf <- ab0 bf + ta   (makextp)
                  bf += (rho/dt) bm1 sum_i bd[i+1] vlag[i]                         (makebdfp; vlag[0] = current)."""
    ta = ab[1] * e1 + ab[2] * e2
    e2[...] = e1
    e1[...] = bf
    bf[...] = ab[0] * bf + ta
    tb = bd[1] * bm1 * vlag[0]
    for i in range(1, len(vlag)):
        tb = tb + bd[i + 1] * bm1 * vlag[i]
    bf += rho_over_dt * tb
    return bf


# ----------------------------------------------------------------------------
# Transposed pieces and the discrete adjoint of the scalar time-stepper
# (exponential_prop%rmatvec, core/linear_operators.f90:84-103, for Nek's scalar step cdscal [UPSTREAM-RECALL])
# ----------------------------------------------------------------------------
BD = {1: (1.0, 1.0, 0.0, 0.0), 2: (1.5, 2.0, -0.5, 0.0), 3: (11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0)}
AB = {1: (1.0, 0.0, 0.0), 2: (2.0, -1.0, 0.0), 3: (3.0, -3.0, 1.0)}


def grad_rst_t(ws, d):
    """Transpose of grad_rst: sum_a D_a^T w_a."""
    if ws[0].ndim == 4:
        return (np.einsum('li,ekjl->ekji', d, ws[0]) + np.einsum('lj,ekli->ekji', d, ws[1])
                + np.einsum('lk,elji->ekji', d, ws[2]))
    return np.einsum('li,ejl->eji', d, ws[0]) + np.einsum('lj,eli->eji', d, ws[1])


def convect_dealiased_t(v, cf, dl):
    """Exact transpose of convect_dealiased on the local points: J^T sum_a D_a^T (c_a o J v)."""
    vf = interp_fine(v, dl['J'])
    return project_coarse(grad_rst_t([c * vf for c in cf], dl['Dg']), dl['J'])


def scalar_steps(glo, mask, geo, n, cf, dl, T0, kappa, dt, nsteps, rho=1.0, tol=1e-13):
    """nsteps BDF/EXT steps (order ramp 1, 2, 3) of rho dT/dt + rho (U.grad) T = kappa lap T from a cold start:
    bq = -rho C T ; makextp / makebdfp ; dssum ; hmholtz.  cf = None: no convection."""
    d = dgll(n)
    lag = [T0.copy(), 0 * T0, 0 * T0]
    e1, e2 = 0 * T0, 0 * T0
    for s in range(1, nsteps + 1):
        o = min(s, 3)
        bq = -rho * convect_dealiased(lag[0], cf, dl) if cf is not None else 0 * T0
        bdf_ext(bq, e1, e2, lag[:o], geo['bm1'], AB[o], BD[o], rho / dt)
        Tn, _, _ = cggo(dssum(bq, glo), geo['g'], d, glo, mask, geo['bm1'], kappa, rho * BD[o][0] / dt, tol=tol, maxit=2000)
        lag = [Tn, lag[0], lag[1]]
    return lag[0]


def scalar_steps_adjoint(glo, mask, geo, n, cf, dl, v, kappa, dt, nsteps, rho=1.0, tol=1e-13):
    """Discrete BM1-adjoint of scalar_steps: <A u, w>_B = <u, A^+ w>_B for continuous masked u, w.  The transposed
    recurrence runs backwards on duals y^m in local right-hand-side form:
        y^N = B v ; for s = N..1: g = hmholtz(dssum(y^s)) ; y^(s-1-j) += ab_j (-rho C^T g) + (rho/dt) bd_(j+1) B g ;
        A^+ v = binvm1 mask dssum(y^0)."""
    d = dgll(n)
    bm1 = geo['bm1']
    y = {nsteps: bm1 * v}
    for s in range(nsteps, 0, -1):
        o = min(s, 3)
        g, _, _ = cggo(dssum(y.pop(s, 0 * v), glo), geo['g'], d, glo, mask, bm1, kappa, rho * BD[o][0] / dt, tol=tol,
                       maxit=2000)
        ct = -rho * convect_dealiased_t(g, cf, dl) if cf is not None else 0 * v
        for j in range(o):
            m = s - 1 - j
            if m < 0:
                continue
            y[m] = y.get(m, 0 * v) + AB[o][j] * ct + (rho / dt) * BD[o][j + 1] * bm1 * g
    return dssum(y.get(0, 0 * v), glo) * mask / dssum(bm1, glo)
