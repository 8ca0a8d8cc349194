"""CPU restatement (numpy) of the pressure-coupled perturbation step that sits inside ``nek_advance`` -- the body of
``exponential_prop%matvec`` (core/linear_operators.f90:225-274: ``nopcopy`` of the Krylov vector into vxp/vyp/vzp/prp,
``nek_advance`` for tau/dt steps from a cold start, copy of the final state back).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nek5000 is not vendored with the reference (SURVEY.md section
2.3), so everything here is [UPSTREAM-RECALL] of Nek5000's P_N - P_N-2 perturbation path (perturb.f: perturbv ->
advabp, makextp, makebdfp, cresvipp, ophinv, incomprp;  navier1.f: opdiv / multd, opgradt / cdtp, cdabdtp, opbinv,
ortho;  coef.f: geom2 / map12) restated from its published formulation (Maday-Patera-Ronquist splitting with the
consistent Poisson operator E = D B^-1 D^T).

Pinning.  The SPATIAL operators are pinned against data the un-vendored solver itself produced: the reference's base
flows (examples/cylinder/BF_1cyl0.f00001, examples/back_fstep/baseflow/BF_bfs0.f00001; velocity and pressure
committed in tests/golden/*_mesh.npz) are converged steady states of exactly this discretisation (lx2 = lx1 - 2,
lxd = 9), and with the operators of this file they satisfy its discrete equations to the tolerances they were
computed with -- max |D U| / bm2 = 7e-10 (cylinder; 6e-2 for a collocation divergence on the GLL mesh), assembled
momentum residual B C(U) U + nu A U - D^T p = 7e-6 of its largest term (O(1) with the other sign of the pressure
term, x 30 with 1.1 nu): tests/test_oracle_fixtures.py, and through the CUDA kernels tests/test_gpu_ns.py.  That
fixes opdiv, opgradt (mesh, metrics, weights, sign, scaling) and, with them, the dealiased convection and axhelm of
oracle/sem.py.  The TIME discretisation (BDF/EXT coefficients, the splitting and the pressure extrapolation) has no
reference data to meet and remains parity unpinned; it is held by independent mathematics in
tests/test_oracle_ns.py: adjointness <D u, p> = <u, D^T p>, exactness of D on polynomial fields, symmetry and null
space of E, discrete incompressibility of every step, linearity and temporal convergence order of the stepper.  The
pressure solver and its preconditioner are solvers: any converged one returns the same dp.

Meshes: velocity on lx1 = N+1 Gauss-Lobatto-Legendre points per direction (C0 across elements), pressure on
lx2 = lx1 - 2 Gauss-Legendre points (element-local, discontinuous).  Arrays are element-local, (e, k, j, i) / (e, j, i).
"""
from __future__ import annotations

import numpy as np

from . import sem as osem


# ----------------------------------------------------------------------------
# Pressure mesh  (coef.f geom2: rxm2 = map12(rxm1) ..., bm2 = w3m2 * jacm2)
# ----------------------------------------------------------------------------
def pressure_setup(n, geo):
    """Interpolation / derivative matrices GLL(lx1) -> GL(lx2) and the metrics on the pressure mesh.

    I12[I, i] = l_i(z2_I) (ixm12), D12 = I12 D (dxm12: derivative of the GLL interpolant at the Gauss points),
    rx2[a*dim+b] = w3m2 * map12(J dr_a/dx_b) -- the Gauss weights folded in, as multd's final col2(dx, w3m2)."""
    lx1, lx2 = n + 1, n - 1
    z1, _ = osem.gll(n)
    z2, w2 = osem.gl(lx2)
    I12 = osem.interp_matrix(z1, z2)
    D12 = I12 @ osem.dgll(n)
    rst = geo['rst']
    dim = 3 if len(rst) == 9 else 2
    w3 = w2[:, None, None] * w2[None, :, None] * w2[None, None, :] if dim == 3 else w2[:, None] * w2[None, :]
    rx2 = [osem.interp_fine(m, I12) * w3[None] for m in rst]
    bm2 = osem.interp_fine(geo['jac'], I12) * w3[None]
    return dict(I12=I12, D12=D12, rx2=rx2, bm2=bm2, dim=dim, lx1=lx1, lx2=lx2)


def _grad12(u, ps):
    """d u / d r_a evaluated on the pressure mesh (multd: dxm12 in direction a, ixm12 in the others)."""
    I, D = ps['I12'], ps['D12']
    if ps['dim'] == 3:
        return (np.einsum('Kk,Jj,Ii,ekji->eKJI', I, I, D, u, optimize=True),
                np.einsum('Kk,Jj,Ii,ekji->eKJI', I, D, I, u, optimize=True),
                np.einsum('Kk,Jj,Ii,ekji->eKJI', D, I, I, u, optimize=True))
    return (np.einsum('Jj,Ii,eji->eJI', I, D, u, optimize=True),
            np.einsum('Jj,Ii,eji->eJI', D, I, u, optimize=True))


def _grad12_t(ws, ps):
    """Transpose of _grad12: sum_a T_a^T w_a (cdtp)."""
    I, D = ps['I12'], ps['D12']
    if ps['dim'] == 3:
        return (np.einsum('Kk,Jj,Ii,eKJI->ekji', I, I, D, ws[0], optimize=True)
                + np.einsum('Kk,Jj,Ii,eKJI->ekji', I, D, I, ws[1], optimize=True)
                + np.einsum('Kk,Jj,Ii,eKJI->ekji', D, I, I, ws[2], optimize=True))
    return (np.einsum('Jj,Ii,eJI->eji', I, D, ws[0], optimize=True)
            + np.einsum('Jj,Ii,eJI->eji', D, I, ws[1], optimize=True))


def opdiv(vel, ps):
    """opdiv / multd: (D u)_q = w_q sum_b sum_a (J dr_a/dx_b)_q (du_b/dr_a)_q  = int q div u on the Gauss points."""
    d = ps['dim']
    out = 0.0
    for b in range(d):
        g = _grad12(vel[b], ps)
        for a in range(d):
            out = out + ps['rx2'][a * d + b] * g[a]
    return out


def opgradt(p, ps):
    """opgradt / cdtp: the exact transpose of opdiv, element-local (no dssum): (D^T p)_b = sum_a T_a^T (rx2[a,b] p)."""
    d = ps['dim']
    return [_grad12_t([ps['rx2'][a * d + b] * p for a in range(d)], ps) for b in range(d)]


def opbinv(ws, glo, mask, binv):
    """opbinv: mask, dssum, multiply by the assembled inverse mass matrix."""
    return [osem.dssum(w, glo) * mask * binv for w in ws]


def ortho(p):
    """ortho: remove the mean over all pressure points (all-Dirichlet velocity: E has the constants as null space)."""
    return p - np.mean(p)


def cdabdtp(p, ps, glo, mask, binv):
    """E p = D B^-1 D^T p  (cdabdtp without the h2inv scaling, which the caller applies to the right-hand side)."""
    return opdiv(opbinv(opgradt(p, ps), glo, mask, binv), ps)


def fdm_setup(n, geo, ps):
    """Element-wise fast-diagonalisation preconditioner for E (the local solves of Nek's Schwarz preconditioner,
    fast.f / hsmg.f, without overlap and without the coarse grid): every element is replaced by the box with its mean
    edge lengths L_a = 2 / mean(|J grad r_a| / J), on which E_e = sum_a c_a (E^ in direction a) x (M^ in the others) with
    the 1-D operators  E^ = D^ b^-1 D^T,  M^ = I^ b^-1 I^T  (D^ = w2 o D12, I^ = w2 o I12, b^ = GLL weights with the two
    end weights doubled = assembled with an equal neighbour).  E^ S = M^ S Lambda, S^T M^ S = 1 gives
    E_e^-1 = (S x S x S) diag(1 / sum_a c_a lambda_ia) (S x S x S)^T."""
    import scipy.linalg as sla
    d = ps['dim']
    _, w1 = osem.gll(n)
    _, w2 = osem.gl(n - 1)
    b = w1.copy()
    b[0] *= 2.0
    b[-1] *= 2.0
    Dh = w2[:, None] * ps['D12']
    Ih = w2[:, None] * ps['I12']
    Eh = (Dh / b) @ Dh.T
    Mh = (Ih / b) @ Ih.T
    lam, S = sla.eigh(Eh, Mh)
    rst, jac = geo['rst'], geo['jac']
    ne = jac.shape[0]
    L = np.zeros((ne, d))
    for a in range(d):
        g = np.sqrt(sum(rst[a * d + bb] ** 2 for bb in range(d))) / jac
        L[:, a] = 2.0 / g.reshape(ne, -1).mean(axis=1)
    c = np.zeros((ne, d))
    for a in range(d):
        c[:, a] = 2.0 / L[:, a]
        for o in range(d):
            if o != a:
                c[:, a] *= L[:, o] / 2.0
    if d == 3:   # den[e, K, J, I] = c_r lam_I + c_s lam_J + c_t lam_K
        den = (c[:, 0, None, None, None] * lam[None, None, None, :] + c[:, 1, None, None, None] * lam[None, None, :, None]
               + c[:, 2, None, None, None] * lam[None, :, None, None])
    else:
        den = c[:, 0, None, None] * lam[None, None, :] + c[:, 1, None, None] * lam[None, :, None]
    return dict(S=S, lam=lam, L=L, c=c, den=den, dim=d, coarse=None)


def coarse_setup(fd, ps, glo, mask, binv):
    """Coarse level of the preconditioner: one constant per element, E_c = R E R^T with R = sum over the element's
    pressure points.  With g_e = D^T 1_e (element-local), E_c[f, e] = sum over the velocity nodes n shared by e and f of
    binvm1_n mask_n g_f(n) . g_e(n)  -- a sparse nel x nel matrix (elements that share a node).  Solved exactly here
    (pseudo-inverse in the mean-free subspace); the device runs Jacobi-CG on it."""
    import scipy.sparse as sp
    d = ps['dim']
    ones = np.ones_like(ps['bm2'])
    g = opgradt(ones, ps)
    ne = ones.shape[0]
    nglob = int(glo.max()) + 1
    rows = np.concatenate([glo.ravel() + b * nglob for b in range(d)])
    cols = np.tile(np.repeat(np.arange(ne), glo[0].size), d)
    vals = np.concatenate([g[b].ravel() for b in range(d)])
    G = sp.csr_matrix((vals, (rows, cols)), shape=(d * nglob, ne))
    wnode = np.zeros(nglob)
    wnode[glo.ravel()] = (binv * mask).ravel()
    Ec = (G.T @ sp.diags(np.tile(wnode, d)) @ G).toarray()
    P = np.eye(ne) - 1.0 / ne
    fd['coarse'] = dict(Ec=Ec, Ecp=P @ np.linalg.pinv(P @ Ec @ P, hermitian=True, rcond=1e-10) @ P)
    return fd


def fdm_apply(r, fd):
    S = fd['S']
    if fd['dim'] == 3:
        t = np.einsum('Kk,Jj,Ii,eKJI->ekji', S, S, S, r, optimize=True) / fd['den']
        z = np.einsum('Kk,Jj,Ii,ekji->eKJI', S, S, S, t, optimize=True)
    else:
        t = np.einsum('Jj,Ii,eJI->eji', S, S, r, optimize=True) / fd['den']
        z = np.einsum('Jj,Ii,eji->eJI', S, S, t, optimize=True)
    if fd.get('coarse') is not None:
        ne = r.shape[0]
        zc = fd['coarse']['Ecp'] @ r.reshape(ne, -1).sum(axis=1)
        z = z + zc.reshape((ne,) + (1,) * (r.ndim - 1))
    return z


def esolve(rhs, ps, glo, mask, binv, tol=1e-10, maxit=2000, mean_free=True, fdm=None):
    """E dp = rhs by preconditioned conjugate gradients: z = M^-1 r ; rtz = sum r z ; p = z + beta p ; w = E p ;
    alpha = rtz / sum w p.  M^-1 = 1 / bm2 (uzawa / uzprec without the Schwarz part) or, with fdm = fdm_setup(..), the
    element-wise fast-diagonalisation solve.  Stops on sqrt(rtz) <= tol * sqrt(rtz_0).
    Returns (dp, iterations, residual drop)."""
    if fdm is None:
        bminv = 1.0 / ps['bm2']
        minv = _Diag(bminv)
    else:
        minv = _Fdm(fdm)
    x = np.zeros_like(rhs)
    p = np.zeros_like(rhs)
    r = ortho(rhs) if mean_free else rhs.copy()
    z = minv * r
    if mean_free:
        z = ortho(z)
    rtz1, rtz2 = float(np.sum(r * z)), 1.0
    r0, rn, it = -1.0, 0.0, 0
    if not rtz1 > 0.0:
        return x, 0, 0.0
    for it in range(1, maxit + 1):
        beta = 0.0 if it == 1 else rtz1 / rtz2
        p = z + beta * p
        w = cdabdtp(p, ps, glo, mask, binv)
        rho = float(np.sum(w * p))
        if not rho > 0.0:
            break
        alpha = rtz1 / rho
        x = x + alpha * p
        r = r - alpha * w
        z = minv * r
        if mean_free:
            z = ortho(z)
        rtz2, rtz1 = rtz1, float(np.sum(r * z))
        rn = np.sqrt(abs(rtz1))
        if r0 < 0.0:
            r0 = np.sqrt(abs(rtz2))
        if rn <= tol * r0:
            break
    return x, it, (rn / r0 if r0 > 0 else 0.0)


class _Diag:
    def __init__(self, d):
        self.d = d

    def __mul__(self, r):
        return self.d * r


class _Fdm:
    def __init__(self, fd):
        self.fd = fd

    def __mul__(self, r):
        return fdm_apply(r, self.fd)


# ----------------------------------------------------------------------------
# The perturbation step  (perturb.f perturbv: advabp, makextp, makebdfp, cresvipp + ophinv, incomprp)
# ----------------------------------------------------------------------------
def advabp(vel, base, cf_base, dl):
    """Explicit term of the linearised momentum equation, local weak form (mass matrix inside the dealiased
    quadrature):  bf_b = -[ (U . grad) v_b + (v . grad) U_b ]."""
    cf_v = osem.set_convect(vel, dl)
    return [-(osem.convect_dealiased(vel[b], cf_base, dl) + osem.convect_dealiased(base[b], cf_v, dl))
            for b in range(len(vel))]


def grad_base(base, geo, n):
    """bm1 * dU_c/dx_b on the GLL points (collocation, gradm1): G[c][b]."""
    d = osem.dgll(n)
    dim = len(base)
    rst, w3 = geo['rst'], geo['bm1'] / geo['jac']
    G = []
    for c in range(dim):
        g = osem.grad_rst(base[c], d)
        G.append([sum(rst[a * dim + b] * g[a] for a in range(dim)) * w3 for b in range(dim)])
    return G


def advabp_adjoint(vel, cf_base, G, dl):
    """Explicit term of the ADJOINT linearised momentum equation (advabp with ifadj; exponential_prop%rmatvec,
    core/linear_operators.f90:84-103, runs Nek's stepper in that mode): the continuous adjoint of
    -(U.grad) v - (v.grad) U is  +(U.grad) w - sum_c w_c grad U_c ; the transport term on the dealiased mesh like the
    forward one, the base-flow-gradient term pointwise on the GLL mesh."""
    dim = len(vel)
    return [osem.convect_dealiased(vel[b], cf_base, dl) - sum(G[c][b] * vel[c] for c in range(dim)) for b in range(dim)]


def ns_steps(glo, mask, geo, n, ps, dl, base, vel0, pr0, nu, dt, nsteps, tol_v=1e-13, tol_p=1e-13, mean_free=True,
             maxit=4000, info=None, fdm=None, adjoint=False, orbit=None):
    """nsteps BDF/EXT steps (order ramp 1, 2, 3, cold start) of the linearised incompressible Navier-Stokes equations
        dv/dt + (U.grad) v + (v.grad) U = -grad p + nu lap v ,  div v = 0
    in the P_N - P_N-2 splitting:
        bf   = EXT(-B C(v)) + (1/dt) B sum_i bd_i v^(n-i)                       (advabp, makextp, makebdfp)
        p*   = p^(n-1)            (order 1, 2)   |   2 p^(n-1) - p^(n-2)       (order 3; extrapprp)
        v*   = H^-1 mask dssum(bf + D^T p*),   H = nu A + (bd_0/dt) B           (cresvipp + ophinv)
        E dp = -(bd_0/dt) D v* ;  v = v* + (dt/bd_0) B^-1 D^T dp ;  p = p* + dp (incomprp)
    base = None: Stokes.  adjoint = True: the same stepper on the adjoint equations (explicit term advabp_adjoint), what
    Nek runs for exponential_prop%rmatvec.  orbit = [U_1, ..., U_nsteps]: the base flow step s linearises about
    (time-periodic base flows: the stored orbit uor / vor / wor of core/linear_operators.f90:254-275 -- step istep runs
    with what was copied into vx after step istep-1, i.e. U_1 = ubase, U_s = uor(:, s-1)); `base` is then unused.
    Returns (velocity list, pressure)."""
    d = osem.dgll(n)
    dim = ps['dim']
    bm1 = geo['bm1']
    binv = 1.0 / osem.dssum(bm1, glo)
    cf_base = osem.set_convect(base, dl) if base is not None else None
    G = grad_base(base, geo, n) if (adjoint and base is not None) else None
    lag = [[v.copy() for v in vel0]] + [[0 * v for v in vel0] for _ in range(2)]
    e1 = [0 * v for v in vel0]
    e2 = [0 * v for v in vel0]
    p, plag = pr0.copy(), 0 * pr0
    its_v = its_p = 0
    for s in range(1, nsteps + 1):
        o = min(s, 3)
        bd0 = osem.BD[o][0]
        if orbit is not None:
            base = orbit[s - 1]
            cf_base = osem.set_convect(base, dl)
            G = grad_base(base, geo, n) if adjoint else None
        if base is None:
            bf = [0 * v for v in vel0]
        elif adjoint:
            bf = advabp_adjoint(lag[0], cf_base, G, dl)
        else:
            bf = advabp(lag[0], base, cf_base, dl)
        for b in range(dim):
            osem.bdf_ext(bf[b], e1[b], e2[b], [lag[i][b] for i in range(o)], bm1, osem.AB[o], osem.BD[o], 1.0 / dt)
        pstar = p if o < 3 else 2.0 * p - plag
        gt = opgradt(pstar, ps)
        vstar = []
        for b in range(dim):
            x, it, _ = osem.cggo(osem.dssum(bf[b] + gt[b], glo), geo['g'], d, glo, mask, bm1, nu, bd0 / dt, tol=tol_v,
                                 maxit=maxit)
            its_v += it
            vstar.append(x)
        rhs = -(bd0 / dt) * opdiv(vstar, ps)
        dp, it, _ = esolve(rhs, ps, glo, mask, binv, tol=tol_p, maxit=maxit, mean_free=mean_free, fdm=fdm)
        its_p += it
        corr = opbinv(opgradt(dp, ps), glo, mask, binv)
        vnew = [vstar[b] + (dt / bd0) * corr[b] for b in range(dim)]
        plag, p = p, pstar + dp
        lag = [vnew, lag[0], lag[1]]
    if info is not None:
        info.update(helmholtz_iterations=its_v, pressure_iterations=its_p)
    return lag[0], p
