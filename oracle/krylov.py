"""Literal CPU restatement of nekStab's Krylov layer.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines (relative to /root/reference) it follows.
Loop order, the two-pass MGS, the LAPACK routines (through scipy) and the sort are
kept as in the Fortran so that the CUDA path can be compared against the reference's
*semantics*; there are no golden H / Ritz vectors in the reference (parity unpinned,
see oracle/__init__.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
from scipy.linalg import lapack


# ----------------------------------------------------------------------------
# Vector type  (core/krylov_subspace.f90:12-17, core/nek_vectors.f90:20-31)
# ----------------------------------------------------------------------------
@dataclass
class Ctx:
    """What Nek's commons provide to the vector routines."""
    bm1s: np.ndarray                 # weight, one entry per velocity point (may contain zeros)
    in_dot: Sequence[bool]           # per field: enters the inner product? (pressure: never)
    time_in_dot: bool = False        # legacy: only if uparam(1)==2.1; new API: always
    bm1s_t: Optional[np.ndarray] = None  # weight for temperature/scalars (same bm1s in Nek)


@dataclass
class KVec:
    f: List[np.ndarray]
    time: float = 0.0

    def copy(self):
        return KVec([a.copy() for a in self.f], self.time)


def k_zero_like(p: KVec) -> KVec:
    return KVec([np.zeros_like(a) for a in p.f], 0.0)


def k_dot(ctx: Ctx, p: KVec, q: KVec) -> float:
    """core/krylov_subspace.f90:26-60 / core/nek_vectors.f90:80-114.

    glsc3 accumulates sequentially over points, one component at a time
    ([UPSTREAM-RECALL] math.f); numpy's pairwise sum differs at the 1e-16 level.
    """
    alpha = 0.0
    for a, b, use in zip(p.f, q.f, ctx.in_dot):
        if use:
            alpha += float(np.sum(a.ravel() * ctx.bm1s.ravel()[:a.size] * b.ravel()))
    if ctx.time_in_dot:
        alpha += p.time * q.time
    if math.isnan(alpha):
        raise FloatingPointError('NaN detected in dot product')  # reference: call nek_end
    return alpha


def k_norm(ctx, p):                       # krylov_subspace.f90:62-73
    return math.sqrt(k_dot(ctx, p, p))


def k_cmult(p: KVec, c: float):           # :94-104
    for a in p.f:
        a *= c
    p.time *= c


def k_normalize(ctx, p: KVec) -> float:   # :75-92
    alpha = k_norm(ctx, p)
    k_cmult(p, 1.0 / alpha)
    return alpha


def k_add2(p: KVec, q: KVec):             # :106-115
    for a, b in zip(p.f, q.f):
        a += b
    p.time += q.time


def k_sub2(p: KVec, q: KVec):             # :117-127
    for a, b in zip(p.f, q.f):
        a -= b
    p.time -= q.time


def k_sub3(p: KVec, q: KVec, r: KVec):    # :129-139
    for a, b, c in zip(p.f, q.f, r.f):
        a[...] = b - c
    p.time = q.time - r.time


def k_copy(p: KVec, q: KVec):             # :152-161 (destination first)
    for a, b in zip(p.f, q.f):
        a[...] = b
    p.time = q.time


def axpby(p: KVec, alpha: float, q: KVec, beta: float, skip_time: bool = True):
    """core/nek_vectors.f90:127-139, 250-256: self <- alpha*self + beta*vec.

    The reference's real_axpby leaves ``time`` untouched (quirk);
    ``skip_time=False`` gives the k_add2-style behaviour.
    """
    for a, b in zip(p.f, q.f):
        a[...] = a * alpha + b * beta
    if not skip_time:
        p.time = p.time * alpha + q.time * beta


def k_matmul(Q: Sequence[KVec], yvec: np.ndarray, k: int) -> KVec:
    """core/krylov_subspace.f90:163-209: dq = sum_i y_i Q_i (all fields and time)."""
    dq = k_zero_like(Q[0])
    for c in range(len(dq.f)):
        stack = np.stack([Q[i].f[c].ravel() for i in range(k)], axis=1)
        dq.f[c][...] = (stack @ yvec[:k]).reshape(dq.f[c].shape)
    dq.time = float(np.dot([Q[i].time for i in range(k)], yvec[:k]))
    return dq


def forward_finite_difference_map(ctx, F: Callable[[KVec], KVec], base: KVec, q: KVec, order: int = 2,
                                  epsilon_base: float = 1e-6) -> KVec:
    """core/matvec.f90:246-379: the linearised forward map as finite differences of the nonlinear map F about the
    base state: eps0 = 1e-6 |base| (:276-277), amplitudes / coefficients of :279-289, the loop of :319-371
    (pert = amp q; F(base + pert); work *= coef; f += work) and the final 1/eps0 (:374)."""
    eps0 = epsilon_base * k_norm(ctx, base)          # epsilon_base: core/main.f90:16
    if order == 2:
        amp, coef = np.array([1.0, -1.0]), np.array([1.0, -1.0]) / 2.0
    elif order == 4:
        amp, coef = np.array([1.0, -1.0, 2.0, -2.0]), np.array([8.0, -8.0, -1.0, 1.0]) / 12.0
    else:
        raise ValueError('findiff_order is 2 or 4')
    amp = amp * eps0
    f = k_zero_like(q)
    for a, cf in zip(amp, coef):
        pert = q.copy()
        k_cmult(pert, float(a))
        x = base.copy()
        k_add2(x, pert)
        work = F(x)
        k_cmult(work, float(cf))
        k_add2(f, work)
    k_cmult(f, 1.0 / eps0)
    return f


# ----------------------------------------------------------------------------
# Arnoldi  (core/krylov_decomposition.f90)
# ----------------------------------------------------------------------------
def update_hessenberg_matrix(ctx, H: np.ndarray, f: KVec, q: Sequence[KVec], k: int,
                             alphas2: Optional[list] = None):
    """core/krylov_decomposition.f90:103-189.  H is the (k+1, k) leading block (0-based
    column k-1 is written).  MGS pass, unconditional second MGS pass, normalise."""
    for i in range(k):                       # :155-168
        wrk = q[i].copy()
        alpha = k_dot(ctx, f, wrk)
        k_cmult(wrk, alpha)
        k_sub2(f, wrk)
        H[i, k - 1] = alpha
    for i in range(k):                       # :171-180
        wrk = q[i].copy()
        alpha = k_dot(ctx, f, wrk)
        k_cmult(wrk, alpha)
        k_sub2(f, wrk)
        H[i, k - 1] += alpha
        if alphas2 is not None:
            alphas2.append(alpha)
    alpha = k_normalize(ctx, f)              # :183
    H[k, k - 1] = alpha                      # :186


def arnoldi_factorization(ctx, matvec: Callable[[KVec], KVec], Q: List[KVec], H: np.ndarray,
                          mstart: int, mend: int, ksize: int):
    """core/krylov_decomposition.f90:2-99 (1-based mstart/mend as in the reference)."""
    if ksize == 0:
        raise ValueError('Krylov base dimension == 0')
    for mstep in range(mstart, mend + 1):
        f = matvec(Q[mstep - 1])                                           # :75
        update_hessenberg_matrix(ctx, H[:mstep + 1, :mstep], f, Q[:mstep], mstep)  # :78
        k_copy(Q[mstep], f)                                                # :81


# ----------------------------------------------------------------------------
# lapack_wrapper.f90 semantics on scipy's LAPACK (same routines)
# ----------------------------------------------------------------------------
def eig(A: np.ndarray):
    """core/lapack_wrapper.f90:114-177: dgeev + complexification + sort by |lambda| desc."""
    n = A.shape[0]
    wr, wi, _, vr, info = lapack.dgeev(np.array(A, order='F'), compute_vl=0, compute_vr=1)
    vals = wr + 1j * wi
    vecs = vr.astype(np.complex128)
    for i in range(n - 1):                   # :167-173
        if wi[i] > 0:
            vecs[:, i] = vr[:, i] + 1j * vr[:, i + 1]
            vecs[:, i + 1] = vr[:, i] - 1j * vr[:, i + 1]
        elif wi[i] == 0:
            vecs[:, i] = vr[:, i]
    sort_eigendecomp(vals, vecs)
    return vecs, vals


def sort_eigendecomp(vals, vecs):
    """core/lapack_wrapper.f90:181-228: selection-style exchange sort, decreasing |lambda|."""
    n = vals.shape[0]
    norm = np.sqrt(vals.real ** 2 + vals.imag ** 2)
    for k in range(n - 1):
        for l in range(k + 1, n):
            if norm[k] < norm[l]:
                norm[k], norm[l] = norm[l], norm[k]
                vals[k], vals[l] = vals[l], vals[k]
                tmp = vecs[:, k].copy()
                vecs[:, k] = vecs[:, l]
                vecs[:, l] = tmp


def select_eigvals(wr, wi):                  # core/lapack_wrapper.f90:232-244
    return math.sqrt(wr * wr + wi * wi) > 0.9


def schur(A: np.ndarray):
    """core/lapack_wrapper.f90:3-55: dgees('V','S',select_eigvals).  Returns T, vecs, vals."""
    T, sdim, wr, wi, vs, work, info = lapack.dgees(select_eigvals, np.array(A, order='F'),
                                                   compute_v=1, sort_t=1)
    return T, vs, wr + 1j * wi


def ordschur(T: np.ndarray, Qm: np.ndarray, selected: np.ndarray):
    """core/lapack_wrapper.f90:59-111: dtrsen('N','V',selected,...)."""
    res = lapack.dtrsen(np.asarray(selected, dtype=np.int32), np.array(T, order='F'),
                        np.array(Qm, order='F'), job='N', wantq=1)
    return res[0], res[1]


def lstsq(A: np.ndarray, b: np.ndarray):
    """core/lapack_wrapper.f90:248-300: dgels('N')."""
    m, n = A.shape
    bb = np.array(b, dtype=np.float64).reshape(m, 1)
    lqr, x, info = lapack.dgels(np.array(A, order='F'), np.array(bb, order='F'))
    return x[:n, 0].copy()


def quicksort2_idx(arr: np.ndarray) -> np.ndarray:
    """Index order produced by core/utils.f90:29-138 (ascending).  Ties are resolved by the
    Numerical-Recipes quicksort there; for distinct magnitudes this equals a stable argsort."""
    return np.argsort(arr, kind='stable')


def select_eigenvalues(vals: np.ndarray, delta: float, nev: int):
    """core/eigensolvers.f90:688-754.  Returns (selected mask, count)."""
    n = vals.shape[0]
    idx = quicksort2_idx(np.abs(vals))
    selected = np.abs(vals) >= (1.0 - delta)            # :743
    selected[idx[n - (nev + 3) - 1:n]] = True           # :746  idx(n-(nev+3):n), 1-based
    a = vals[idx[n - (nev + 3) - 1]].imag
    b = vals[idx[n - (nev + 4) - 1]].imag
    if a == -b:                                         # :747-749
        selected[idx[n - (nev + 4) - 1]] = True
    return selected, int(np.count_nonzero(selected))


def schur_condensation(mstart: int, H: np.ndarray, Q: List[KVec], ksize: int,
                       schur_del: float, schur_tgt: int) -> int:
    """core/eigensolvers.f90:363-468.  Returns the new (1-based) mstart."""
    b_vec = np.zeros(ksize)
    b_vec[ksize - 1] = H[ksize, ksize - 1]                          # :403
    T, vecs, vals = schur(H[:ksize, :ksize])                        # :407
    selected, mstart = select_eigenvalues(vals, schur_del, schur_tgt, )  # :410
    T, vecs = ordschur(T, vecs, selected)                           # :414
    H[:ksize, :ksize] = T
    H[:mstart, mstart:ksize] = 0.0                                  # :417
    H[mstart:ksize + 1, :] = 0.0                                    # :418
    for c in range(len(Q[0].f)):                                    # :421-442  Q(:,1:k) <- Q(:,1:k) Z
        stack = np.stack([Q[i].f[c].ravel() for i in range(ksize)], axis=1)
        stack = stack @ vecs
        for i in range(ksize):
            Q[i].f[c][...] = stack[:, i].reshape(Q[i].f[c].shape)
    # the reference does not rotate %time (it is not packed into qx..qt): keep as is.
    b_vec = b_vec @ vecs                                            # :446
    H[mstart, :] = b_vec                                            # :447
    mstart += 1                                                     # :450
    for c in range(len(Q[0].f)):                                    # :452-453 Q(mstart) <- Q(k+1)
        Q[mstart - 1].f[c][...] = Q[ksize].f[c]
    return mstart


@dataclass
class KSResult:
    vals: np.ndarray
    vecs: np.ndarray
    residual: np.ndarray
    cnt: int
    schur_cnt: int
    H: np.ndarray
    Q: List[KVec] = field(default_factory=list)


def krylov_schur(ctx, matvec, q0: KVec, k_dim=100, schur_tgt=2, eigen_tol=1e-6, schur_del=0.10,
                 max_restarts=200) -> KSResult:
    """core/eigensolvers.f90:120-359 (seed handling reduced to a given unit-norm-able q0)."""
    Q = [k_zero_like(q0) for _ in range(k_dim + 1)]
    H = np.zeros((k_dim + 1, k_dim))
    k_copy(Q[0], q0)
    mstart, schur_cnt = 1, 0
    while True:
        arnoldi_factorization(ctx, matvec, Q, H, mstart, k_dim, k_dim)      # :297
        vecs, vals = eig(H[:k_dim, :k_dim])                                 # :306
        residual = np.abs(H[k_dim, k_dim - 1] * vecs[k_dim - 1, :])         # :309
        cnt = int(np.count_nonzero(residual < eigen_tol))                   # :310
        if schur_tgt <= 0 or cnt >= schur_tgt or schur_cnt >= max_restarts:  # :314-331
            break
        schur_cnt += 1
        mstart = schur_condensation(mstart, H, Q, k_dim, schur_del, schur_tgt)
    return KSResult(vals, vecs, residual, cnt, schur_cnt, H, Q)


# ----------------------------------------------------------------------------
# GMRES  (core/newton_krylov.f90:170-326)
# ----------------------------------------------------------------------------
def ts_gmres(ctx, matvec, rhs: KVec, maxiter: int, ksize: int, tol: float):
    """core/newton_krylov.f90:170-299.  Returns (sol, residual history, matvec calls).

    The reference's use of ``k`` after loop exhaustion (k_dim+1) is not
    replicated: min(k, ksize) columns are combined.
    """
    sol = k_zero_like(rhs)
    Q = [k_zero_like(rhs) for _ in range(ksize + 1)]
    k_copy(Q[0], rhs)
    beta = k_normalize(ctx, Q[0])                                           # :241-242
    hist, calls = [], 0
    for _ in range(maxiter):                                                # :245
        H = np.zeros((ksize + 1, ksize))
        evec = np.zeros(ksize + 1)
        evec[0] = beta
        for j in range(1, ksize + 1):
            Q[j] = k_zero_like(rhs)
        kk = ksize
        yvec = np.zeros(ksize)
        for k in range(1, ksize + 1):                                       # :250
            arnoldi_factorization(ctx, matvec, Q, H, k, k, ksize)           # :252
            calls += 1
            yvec[:k] = lstsq(H[:k + 1, :k], evec[:k + 1])                   # :255
            beta = float(np.linalg.norm(evec[:k + 1] - H[:k + 1, :k] @ yvec[:k]))  # :258
            if beta ** 2 < tol:                                             # :266
                kk = k
                break
        dq = k_matmul(Q, yvec, kk)                                          # :279
        k_add2(sol, dq)                                                     # :280
        k_copy(Q[0], sol)                                                   # :283
        f = matvec(Q[0])                                                    # :303-326
        calls += 1
        k_sub2(f, rhs)
        k_cmult(f, -1.0)
        beta = k_normalize(ctx, f)
        k_copy(Q[0], f)
        hist.append(beta ** 2)
        if beta ** 2 < tol:                                                 # :293
            break
    return sol, hist, calls


# ----------------------------------------------------------------------------
# Newton-Krylov fixed-point iteration  (core/newton_krylov.f90:1-168)
# ----------------------------------------------------------------------------
def newton_krylov(ctx, forward_map, jacobian_for, q: KVec, maxiter_newton: int, maxiter_gmres: int, ksize: int,
                  tol: float):
    """core/newton_krylov.f90:52-133: f = F(q) (:102); residual = |f|^2 (:107); stop when residual < tol (:117);
    dq = ts_gmres(J, f) (:125); q -= dq (:130).  ``jacobian_for(q)`` returns the matvec of the linearisation
    about q (prepare_linearized_solver, :71).  Returns (q, residual history, linear-solver calls)."""
    hist, calls = [], 0
    for _ in range(maxiter_newton):
        f = forward_map(q)
        residual = k_norm(ctx, f) ** 2
        hist.append(residual)
        if residual < tol:
            break
        dq, _, c = ts_gmres(ctx, jacobian_for(q), f, maxiter_gmres, ksize, tol)
        calls += c
        k_sub2(q, dq)
    return q, hist, calls


# ----------------------------------------------------------------------------
# Seed noise  (core/utils.f90:297-359, 408-418)
# ----------------------------------------------------------------------------
def mth_rand(ix, iy, iz, ieg, xl, fc, if3d):
    """core/utils.f90:408-418 (1-based ix,iy,iz,ieg; xl = point coordinates)."""
    r = fc[0] * (ieg + xl[0] * np.sin(xl[1])) + fc[1] * ix * iy + fc[2] * ix
    if if3d:
        r = fc[0] * (ieg + xl[2] * np.sin(r)) + fc[1] * iz * ix + fc[2] * iz
    return np.cos(1.0e3 * np.sin(1.0e3 * np.sin(r)))


NOISE_FC = ((3.0e4, -1.5e3, 0.5e5), (2.3e4, 2.3e3, -2.0e5), (2.0e4, 1.0e3, 1.0e5))  # utils.f90:321-329


def op_add_noise(coords, if3d=True):
    """Raw noise fields of core/utils.f90:312-336 (before dssum/vmult/mask), element ids 1-based."""
    x = coords[0]
    nel = x.shape[0]
    lx = x.shape[-1]
    ieg = np.arange(1, nel + 1).reshape((nel,) + (1,) * (x.ndim - 1))
    i1 = np.arange(1, lx + 1)
    if if3d:
        ix, iy, iz = i1[None, None, None, :], i1[None, None, :, None], i1[None, :, None, None]
    else:
        ix, iy, iz = i1[None, None, :], i1[None, :, None], 1
    out = []
    for c in range(3 if if3d else 2):
        out.append(mth_rand(ix, iy, iz, ieg, coords, NOISE_FC[c], if3d) + np.zeros_like(x))
    return out


# ----------------------------------------------------------------------------
# BM1-weighted QR of BoostConv  (core/fixedp.f90:331-385)
# ----------------------------------------------------------------------------
def qr_dec(bm1, X):
    """X: list of KVec (velocity fields only).  Single-pass MGS with the bm1 weight; returns Q (list of
    KVec) and rr (k x k) exactly as the reference loops do (norm^2 < 1e-60 -> zero column, rr(j,j) = 1)."""
    k = len(X)
    ctx = Ctx(bm1s=bm1, in_dot=[True] * len(X[0].f))
    rr = np.zeros((k, k))
    Q = []
    for j in range(k):
        dum = X[j].copy()
        for i in range(j):
            rr[i, j] = k_dot(ctx, dum, Q[i])
            t = Q[i].copy()
            k_cmult(t, rr[i, j])
            k_sub2(dum, t)
        norma = k_dot(ctx, dum, dum)
        if norma < 1e-60:
            norma = 1.0
            dum = k_zero_like(dum)
        else:
            k_cmult(dum, 1.0 / math.sqrt(norma))
        Q.append(dum)
        rr[j, j] = math.sqrt(norma)
    return Q, rr


# ----------------------------------------------------------------------------
# LightKrylov-style step-wise eigensolver (call site core/linear_stab.f90:66)
# ----------------------------------------------------------------------------
def eigs(ctx, matvec, q0: KVec, k_dim: int, nev: int, tol: float):
    """[UPSTREAM-RECALL] LightKrylov eigs (not vendored; parity unpinned): Arnoldi one step at a time,
    eig(H_k) and residuals |H(k+1,k) y_k| after every step, stop when nev pairs are below tol."""
    Q = [k_zero_like(q0) for _ in range(k_dim + 1)]
    k_copy(Q[0], q0)
    H = np.zeros((k_dim + 1, k_dim))
    vecs = vals = residual = None
    k = 0
    for k in range(1, k_dim + 1):
        arnoldi_factorization(ctx, matvec, Q, H, k, k, k_dim)
        vecs, vals = eig(H[:k, :k])
        residual = np.abs(H[k, k - 1] * vecs[k - 1, :])
        if int(np.count_nonzero(residual < tol)) >= nev:
            break
    return vals, vecs, residual, k, H


# ----------------------------------------------------------------------------
# LightKrylov-style step-wise singular-value solver (call site core/linear_stab.f90:112)
# ----------------------------------------------------------------------------
def svds(ctx, matvec, rmatvec, u0: KVec, k_dim: int, nev: int, tol: float):
    """[UPSTREAM-RECALL] LightKrylov svds + lanczos_bidiagonalization (not vendored; parity unpinned):
    v_k = A^T u_k, one modified Gram-Schmidt sweep against V(1:k-1), alpha = |v_k| = B(k,k);
    u_k+1 = A v_k, one sweep against U(1:k), beta = |u_k+1| = B(k+1,k); svd(B(1:k,1:k)) and residuals
    |beta * vvecs(k,:)| after every step, stop when nev triplets are below tol."""
    U = [k_zero_like(u0) for _ in range(k_dim + 1)]
    V = [k_zero_like(u0) for _ in range(k_dim)]
    k_copy(U[0], u0)
    B = np.zeros((k_dim + 1, k_dim))
    sig = uv = vv = residual = None
    kdone = 0
    for k in range(1, k_dim + 1):
        v = rmatvec(U[k - 1])
        for j in range(k - 1):
            g = k_dot(ctx, v, V[j])
            axpby(v, 1.0, V[j], -g, skip_time=False)
        alpha = k_norm(ctx, v)
        B[k - 1, k - 1] = alpha
        if not alpha > tol:
            break
        k_cmult(v, 1.0 / alpha)
        k_copy(V[k - 1], v)
        u = matvec(V[k - 1])
        for j in range(k):
            g = k_dot(ctx, u, U[j])
            axpby(u, 1.0, U[j], -g, skip_time=False)
        beta = k_norm(ctx, u)
        B[k, k - 1] = beta
        if beta > tol:
            k_cmult(u, 1.0 / beta)
        k_copy(U[k], u)
        kdone = k
        uv, sig, vt = np.linalg.svd(B[:k, :k])
        vv = vt.T
        residual = np.abs(beta * vv[k - 1, :])
        if int(np.count_nonzero(residual < tol)) >= nev or not beta > tol:
            break
    return sig, uv, vv, residual, kdone, B
