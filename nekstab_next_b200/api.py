"""Host-side mirror of the reference's vector / operator / solver interface over the C ABI.

Names, argument meaning and error behaviour follow the reference so that tests read like its own
routines (paths relative to the nekStab repository root):

* ``nek_dvector``                      -- ``real_nek_vector`` (core/nek_vectors.f90:20-31) /
                                          ``krylov_vector`` (core/krylov_subspace.f90:12-17)
* ``k_dot, k_norm, k_normalize, ...``  -- core/krylov_subspace.f90:26-209
* ``arnoldi_factorization``            -- core/krylov_decomposition.f90:2  (1-based mstart/mend)
* ``krylov_schur, schur_condensation`` -- core/eigensolvers.f90:120, 363
* ``ts_gmres``                         -- core/newton_krylov.f90:170
* ``eig, schur, ordschur, lstsq``      -- core/lapack_wrapper.f90

Everything computes on the GPU through libnekstab_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import check, c_double_p, c_i64_p, c_int_p

ORTH_MGS2_REF, ORTH_CGS2, ORTH_DGKS = 0, 1, 2
AXPBY_SKIP_TIME = 1


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def _f64(a, copy=False) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.copy() if copy else a


# ------------------------------------------------------------------------------------------------
# LAPACK provider: scipy's cython_lapack entry points (same routines as lapack_wrapper.f90)
# ------------------------------------------------------------------------------------------------
_lapack_set = False


def set_lapack_from_scipy():
    global _lapack_set
    if _lapack_set:
        return
    import scipy.linalg.cython_lapack as cl
    api = C.pythonapi
    api.PyCapsule_GetName.restype = C.c_char_p
    api.PyCapsule_GetName.argtypes = [C.py_object]
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    ptrs = []
    for name in ('dgeev', 'dgees', 'dtrsen', 'dgels', 'dgesvd'):
        cap = cl.__pyx_capi__[name]
        ptrs.append(api.PyCapsule_GetPointer(cap, api.PyCapsule_GetName(cap)))
    check(_capi.load().nsb_set_lapack(*ptrs[:4]))
    check(_capi.load().nsb_set_lapack_svd(ptrs[4]))
    _lapack_set = True


# ------------------------------------------------------------------------------------------------
# context / layout / basis
# ------------------------------------------------------------------------------------------------
class Context:
    """One per process / GPU (the Nek rank)."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, unique_id: Optional[bytes] = None):
        self.lib = _capi.load()
        h = C.c_void_p()
        buf = None
        if unique_id is not None:
            buf = C.create_string_buffer(unique_id, 128)
        check(self.lib.nsb_init(device, rank, nranks, buf, C.byref(h)))
        self.h = h
        self.rank, self.nranks, self.device = rank, nranks, device

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_capi.load().nsb_get_unique_id(buf))
        return buf.raw

    def connect_peers(self, allgather, halo_bytes: int = 64 << 20):
        """Map every rank's NVLink mailbox.  ``allgather(bytes) -> list[bytes]`` is the host's own
        transport (torch.distributed.all_gather_object in the tests, MPI_Allgather in Nek)."""
        if self.nranks == 1:
            return False
        buf = C.create_string_buffer(64)
        check(self.lib.nsb_p2p_mailbox_create(self.h, int(halo_bytes), buf))
        handles = allgather(buf.raw)
        allb = C.create_string_buffer(b''.join(handles), 64 * self.nranks)
        check(self.lib.nsb_p2p_mailbox_connect(self.h, allb))
        return True

    def p2p_enabled(self) -> bool:
        v = C.c_int()
        check(self.lib.nsb_p2p_enabled(self.h, C.byref(v)))
        return bool(v.value)

    def sync(self):
        check(self.lib.nsb_sync(self.h))

    def set_dgks_eta(self, eta: float):
        """DGKS threshold: second projection when |w'| < eta |w| (default 1/sqrt 2)."""
        check(self.lib.nsb_set_dgks_eta(self.h, float(eta)))

    def launch_count(self) -> int:
        n = C.c_int64()
        check(self.lib.nsb_launch_count(self.h, C.byref(n)))
        return n.value

    def timer_start(self):
        check(self.lib.nsb_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(self.lib.nsb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(self.lib.nsb_flush_l2(self.h))

    PROF_CLASSES = ('multidot', 'update', 'normalize', 'axhelm', 'gather_scatter', 'blas1', 'small',
                    'rotate', 'gemv', 'dot', 'fused_update_dot')

    def prof_enable(self, on: bool = True):
        check(self.lib.nsb_prof_enable(self.h, int(on)))

    def prof_report(self) -> dict:
        """{class: dict(ms, launches, bytes)} for the launches recorded since prof_enable."""
        out = {}
        for i, name in enumerate(self.PROF_CLASSES):
            ms, n, b = C.c_double(), C.c_int64(), C.c_double()
            check(self.lib.nsb_prof_get(self.h, i, C.byref(ms), C.byref(n), C.byref(b)))
            if n.value:
                out[name] = dict(ms=ms.value, launches=n.value, bytes=b.value)
        return out

    def allreduce(self, x: np.ndarray) -> np.ndarray:
        x = _f64(x, copy=True).ravel()
        check(self.lib.nsb_allreduce_host(self.h, _dp(x), x.size))
        return x

    def close(self):
        if self.h:
            self.lib.nsb_finalize(self.h)
            self.h = None


class Layout:
    """Field lengths of one state vector {vx, vy, [vz], [pr], [t..]} + which enter the dot."""

    def __init__(self, ctx: Context, field_len: Sequence[int], field_in_dot: Sequence[bool],
                 time_in_dot: bool = False, c0_sem: 'Optional[Sem]' = None, n_c0: int = 0):
        """``c0_sem`` / ``n_c0``: store the first n_c0 (continuous) fields on the distinct nodes of that mesh
        (nsb_layout_create_c0); field_len stays the element-local length, the host interface is unchanged."""
        self.ctx, self.lib = ctx, ctx.lib
        self.field_len = [int(n) for n in field_len]
        self.field_in_dot = [bool(b) for b in field_in_dot]
        self.nfields = len(self.field_len)
        fl = (C.c_int64 * self.nfields)(*self.field_len)
        fd = (C.c_int * self.nfields)(*[int(b) for b in self.field_in_dot])
        h = C.c_void_p()
        if c0_sem is not None:
            check(self.lib.nsb_layout_create_c0(ctx.h, c0_sem.h, self.nfields, fl, fd, int(time_in_dot),
                                                int(n_c0 or self.nfields), C.byref(h)))
            self._keep = c0_sem
        else:
            check(self.lib.nsb_layout_create(ctx.h, self.nfields, fl, fd, int(time_in_dot), C.byref(h)))
        self.h = h
        nc0, rows = C.c_int(), C.c_int64()
        check(self.lib.nsb_layout_is_c0(h, C.byref(nc0), C.byref(rows)))
        self.n_c0, self.c0_rows = nc0.value, rows.value
        ld, ndot, ndof = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.nsb_layout_info(h, C.byref(ld), C.byref(ndot), C.byref(ndof)))
        self.ld, self.ndot, self.ndof_dot = ld.value, ndot.value, ndof.value

    def set_weight(self, weights: Sequence[np.ndarray]):
        """One bm1s array per in-dot field (pass the same array several times to reuse it)."""
        dot_fields = [i for i, b in enumerate(self.field_in_dot) if b]
        if len(weights) != len(dot_fields):
            raise ValueError(f'need {len(dot_fields)} weight arrays, got {len(weights)}')
        keep = []
        for w, f in zip(weights, dot_fields):
            a = _f64(w).ravel()
            if a.size != self.field_len[f]:
                raise ValueError(f'weight for field {f} has {a.size} entries, expected {self.field_len[f]}')
            keep.append(a)
        arr = (c_double_p * len(keep))(*[_dp(a) for a in keep])
        check(self.lib.nsb_layout_set_weight(self.h, arr))

    def close(self):
        if self.h:
            self.lib.nsb_layout_destroy(self.h)
            self.h = None


class Basis:
    """Device-resident column-major tall-skinny fp64 array; column c is the vector ``self[c]``."""

    def __init__(self, layout: Layout, ncols: int):
        self.layout, self.lib, self.ncols = layout, layout.lib, int(ncols)
        h = C.c_void_p()
        check(self.lib.nsb_basis_create(layout.h, self.ncols, C.byref(h)))
        self.h = h

    def __getitem__(self, col: int) -> 'nek_dvector':
        if col < 0:
            col += self.ncols
        return nek_dvector(self, col)

    def __len__(self):
        return self.ncols

    def col_ptr(self, col: int) -> int:
        p = C.c_uint64()
        check(self.lib.nsb_basis_col_ptr(self.h, col, C.byref(p)))
        return p.value

    def gram(self, k: int) -> np.ndarray:
        G = np.zeros((k, k), order='F')
        check(self.lib.nsb_basis_gram(self.h, k, _dp(G), k))
        return G

    def qr(self, k: int, mode: int = ORTH_CGS2) -> np.ndarray:
        """In-place BM1-weighted QR of the first k columns (qr_dec, core/fixedp.f90:331-385); returns R."""
        R = np.zeros((k, k), order='F')
        check(self.lib.nsb_basis_qr(self.h, k, mode, _dp(R), k))
        return R

    def rotate(self, k: int, Z: np.ndarray, rotate_time: bool = False):
        Zf = np.asfortranarray(Z, dtype=np.float64)
        check(self.lib.nsb_basis_rotate(self.h, k, _dp(Zf), Zf.shape[0], int(rotate_time)))

    def close(self):
        if self.h:
            self.lib.nsb_basis_destroy(self.h)
            self.h = None


class nek_dvector:
    """A (basis, column) pair: the device vector with the reference's type-bound procedures."""

    __slots__ = ('basis', 'col')

    def __init__(self, basis: Basis, col: int):
        self.basis, self.col = basis, col

    # -- host transfer ------------------------------------------------------------------------
    def upload(self, fields: Sequence[Optional[np.ndarray]], time: float = 0.0):
        L = self.basis.layout
        keep, ptrs = [], []
        for f, n in zip(fields, L.field_len):
            if f is None:
                ptrs.append(None)
                continue
            a = _f64(f).ravel()
            if a.size != n:
                raise ValueError(f'field has {a.size} entries, layout expects {n}')
            keep.append(a)
            ptrs.append(_dp(a))
        arr = (c_double_p * L.nfields)(*ptrs)
        check(self.basis.lib.nsb_vec_upload(self.basis.h, self.col, arr, float(time)))
        return self

    def download(self):
        L = self.basis.layout
        out = [np.empty(n) for n in L.field_len]
        arr = (c_double_p * L.nfields)(*[_dp(a) for a in out])
        t = C.c_double()
        check(self.basis.lib.nsb_vec_download(self.basis.h, self.col, arr, C.byref(t)))
        return out, t.value

    # -- type-bound procedures (core/nek_vectors.f90:27-30) -----------------------------------
    def zero(self):
        check(self.basis.lib.nsb_vec_zero(self.basis.h, self.col))

    def dot(self, vec: 'nek_dvector') -> float:
        a = C.c_double()
        check(self.basis.lib.nsb_vec_dot(self.basis.h, self.col, vec.basis.h, vec.col, C.byref(a)))
        return a.value

    def norm(self) -> float:
        a = C.c_double()
        check(self.basis.lib.nsb_vec_norm(self.basis.h, self.col, C.byref(a)))
        return a.value

    def scal(self, alpha: float):
        check(self.basis.lib.nsb_vec_scal(self.basis.h, self.col, float(alpha)))

    def axpby(self, alpha: float, vec: 'nek_dvector', beta: float, skip_time: bool = True):
        """self <- alpha*self + beta*vec; like real_axpby, %time is left alone by default."""
        check(self.basis.lib.nsb_vec_axpby(self.basis.h, self.col, float(alpha), vec.basis.h, vec.col,
                                           float(beta), AXPBY_SKIP_TIME if skip_time else 0))


# -- legacy free functions (core/krylov_subspace.f90:26-209) -----------------------------------
def k_dot(p: nek_dvector, q: nek_dvector) -> float:
    return p.dot(q)


def k_norm(p: nek_dvector) -> float:
    return p.norm()


def k_normalize(p: nek_dvector) -> float:
    a = C.c_double()
    check(p.basis.lib.nsb_vec_normalize(p.basis.h, p.col, C.byref(a)))
    return a.value


def k_cmult(p: nek_dvector, c: float):
    p.scal(c)


def k_add2(p: nek_dvector, q: nek_dvector):
    check(p.basis.lib.nsb_vec_add2(p.basis.h, p.col, q.basis.h, q.col))


def k_sub2(p: nek_dvector, q: nek_dvector):
    check(p.basis.lib.nsb_vec_sub2(p.basis.h, p.col, q.basis.h, q.col))


def k_sub3(p: nek_dvector, q: nek_dvector, r: nek_dvector):
    check(p.basis.lib.nsb_vec_sub3(p.basis.h, p.col, q.basis.h, q.col, r.basis.h, r.col))


def k_zero(p: nek_dvector):
    p.zero()


def k_copy(p: nek_dvector, q: nek_dvector):
    """Destination first, like the reference."""
    check(p.basis.lib.nsb_vec_copy(p.basis.h, p.col, q.basis.h, q.col))


def k_matmul(dq: nek_dvector, Q: Basis, yvec: np.ndarray, k: int):
    y = _f64(yvec).ravel()
    check(Q.lib.nsb_basis_gemv(Q.h, int(k), _dp(y), dq.basis.h, dq.col))


def orthonormalize(Q: Basis, k: int, col_w: int, mode: int = ORTH_CGS2):
    """update_hessenberg_matrix on the device: returns (h[0..k], passes)."""
    h = np.zeros(k + 1)
    passes = C.c_int()
    check(Q.lib.nsb_orthonormalize(Q.h, int(k), int(col_w), int(mode), _dp(h), C.byref(passes)))
    return h, passes.value


# ------------------------------------------------------------------------------------------------
# spectral-element mesh + operators
# ------------------------------------------------------------------------------------------------
def gll(N: int):
    lib = _capi.load()
    z, w, D = np.zeros(N + 1), np.zeros(N + 1), np.zeros((N + 1, N + 1), order='F')
    check(lib.nsb_gll(N, _dp(z), _dp(w), _dp(D)))
    return z, w, np.ascontiguousarray(D)  # D[i, j] = dxm1(i, j)


class Sem:
    """A mesh partition on the device: geometry, masks, gather-scatter lists."""

    SEL = dict(bm1=0, jac=1, binvm1=2, vmult=3, mask=4, g1=10, g2=11, g3=12, g4=13, g5=14, g6=15)

    def __init__(self, ctx: Context, N: int, x, y, z=None, mask=None, glo_num=None):
        self.ctx, self.lib = ctx, ctx.lib
        x = _f64(x)
        self.dim = 3 if z is not None else 2
        self.N, self.lx = N, N + 1
        self.shape = x.shape
        self.nel = x.shape[0]
        if glo_num is None:
            raise ValueError('Sem: glo_num (Nek\'s global node numbering of every local point) is required')
        glo = np.ascontiguousarray(glo_num, dtype=np.int64)
        h = C.c_void_p()
        check(self.lib.nsb_sem_create(ctx.h, self.dim, N, self.nel, _dp(x), _dp(_f64(y)),
                                      _dp(_f64(z)) if z is not None else None,
                                      _dp(_f64(mask)) if mask is not None else None,
                                      glo.ctypes.data_as(c_i64_p), C.byref(h)))
        self.h = h
        self.npts = self.lib.nsb_sem_npts(h)

    def setup_exchange(self):
        check(self.lib.nsb_sem_setup_exchange(self.h))

    def get(self, name: str) -> np.ndarray:
        out = np.empty(self.npts)
        check(self.lib.nsb_sem_get(self.h, self.SEL[name], _dp(out)))
        return out.reshape(self.shape)

    def axhelm(self, vin: nek_dvector, vout: nek_dvector, field: int, h1: float, h2: float):
        check(self.lib.nsb_sem_axhelm(self.h, vin.basis.h, vin.col, vout.basis.h, vout.col, field, h1, h2))

    def dssum(self, v: nek_dvector, field: int):
        check(self.lib.nsb_sem_dssum(self.h, v.basis.h, v.col, field))

    def col2(self, v: nek_dvector, field: int, which: str):
        check(self.lib.nsb_sem_col2(self.h, v.basis.h, v.col, field, self.SEL[which]))

    def hmholtz(self, rhs: nek_dvector, x: nek_dvector, field: int, h1: float, h2: float, tol: float = 1e-10,
                maxit: int = 500):
        """Jacobi-PCG solve of (h1 A + h2 B) x = rhs (Nek's hmholtz/cggo); returns (iterations, residual drop)."""
        it, res = C.c_int(), C.c_double()
        check(self.lib.nsb_sem_hmholtz(self.h, rhs.basis.h, rhs.col, x.basis.h, x.col, field, h1, h2, tol, maxit,
                                       C.byref(it), C.byref(res)))
        return it.value, res.value

    def hmholtz_vec(self, rhs: nek_dvector, x: nek_dvector, field0: int, nf: int, h1: float, h2: float,
                    tol: float = 1e-10, maxit: int = 500):
        """nf <= 3 Helmholtz systems side by side (Nek's ophinv); returns (iterations[nf], residual drops[nf])."""
        it = (C.c_int * nf)()
        res = (C.c_double * nf)()
        check(self.lib.nsb_sem_hmholtz_vec(self.h, rhs.basis.h, rhs.col, x.basis.h, x.col, int(field0), int(nf), h1, h2,
                                           tol, maxit, it, res))
        return list(it), list(res)

    def ax(self, vin: nek_dvector, vout: nek_dvector, field: int, h1: float, h2: float):
        check(self.lib.nsb_sem_ax(self.h, vin.basis.h, vin.col, vout.basis.h, vout.col, field, h1, h2))

    # -- pressure mesh of the P_N - P_N-2 splitting ([UPSTREAM-RECALL] navier1.f opdiv / opgradt / cdabdtp / uzawa) --
    def pressure_setup(self) -> int:
        """Metrics on the lx2 = lx1 - 2 Gauss-Legendre mesh; returns the number of pressure points of this rank."""
        check(self.lib.nsb_sem_pressure_setup(self.h))
        self.n2 = int(self.lib.nsb_sem_npres(self.h))
        return self.n2

    def pressure_get(self, name: str) -> np.ndarray:
        n2 = int(self.lib.nsb_sem_npres(self.h))
        which = dict(rx2=0, bm2inv=1)[name]
        out = np.empty(self.dim * self.dim * n2 if which == 0 else n2)
        check(self.lib.nsb_sem_pressure_get(self.h, which, _dp(out)))
        return out.reshape(self.dim * self.dim, n2) if which == 0 else out

    def norm_grad(self, vec: nek_dvector) -> float:
        """norm_grad (core/utils.f90:446-486): sum_c sum_b (du_c/dx_b, du_c/dx_b)_bm1s of the velocity fields, no
        square root -- the spurious-mode measure of outpost_ks."""
        a = C.c_double()
        check(self.lib.nsb_sem_norm_grad(self.h, vec.basis.h, vec.col, C.byref(a)))
        return a.value

    def compute_cfl(self, vec: nek_dvector, dt: float = 1.0) -> float:
        """compute_cfl(cfl, vx, vy, vz, dt) of the velocity fields (call sites core/linear_stab.f90:222,231)."""
        a = C.c_double()
        check(self.lib.nsb_sem_cfl(self.h, vec.basis.h, vec.col, float(dt), C.byref(a)))
        return a.value

    def opdiv(self, vin: nek_dvector, vout: nek_dvector):
        """opdiv: pressure field of vout <- D (velocity fields of vin)."""
        check(self.lib.nsb_sem_opdiv(self.h, vin.basis.h, vin.col, vout.basis.h, vout.col))

    def opgradt(self, vin: nek_dvector, vout: nek_dvector):
        """opgradt: velocity fields of vout <- D^T (pressure field of vin), element-local."""
        check(self.lib.nsb_sem_opgradt(self.h, vin.basis.h, vin.col, vout.basis.h, vout.col))

    def cdabdtp(self, vin: nek_dvector, vout: nek_dvector):
        """cdabdtp: pressure field of vout <- D B^-1 D^T (pressure field of vin)."""
        check(self.lib.nsb_sem_cdabdtp(self.h, vin.basis.h, vin.col, vout.basis.h, vout.col))

    def esolve(self, rhs: nek_dvector, x: nek_dvector, tol: float = 1e-10, maxit: int = 2000, mean_free: bool = True,
               precond: int = 1):
        """E x = rhs on the pressure fields by preconditioned CG (precond 0: 1 / bm2; 1: element-wise fast
        diagonalisation + coarse level); returns (iterations, residual drop)."""
        it, res = C.c_int(), C.c_double()
        check(self.lib.nsb_sem_esolve(self.h, rhs.basis.h, rhs.col, x.basis.h, x.col, float(tol), int(maxit),
                                      int(bool(mean_free)), int(precond), C.byref(it), C.byref(res)))
        return it.value, res.value

    # -- time-stepper pieces around ax ([UPSTREAM-RECALL] convect.f, perturb.f) --
    def dealias_setup(self, lxd: int = 0):
        """Fine-mesh metrics for the dealiased convection (lxd Gauss-Legendre points; 0 = 3 lx1 / 2)."""
        check(self.lib.nsb_sem_dealias_setup(self.h, int(lxd)))

    def set_convect(self, slot: int, vel: nek_dvector, field0: int = 0):
        """set_convect_new: slot <- contravariant fine-mesh form of the velocity in fields field0..field0+2."""
        check(self.lib.nsb_sem_set_convect(self.h, int(slot), vel.basis.h, vel.col, int(field0)))

    def convect(self, slot: int, vin: nek_dvector, vout: nek_dvector, field0: int = 0, nf: int = 1,
                scale: float = 1.0, accumulate: bool = False):
        """convect_new: vout (+)= scale * J^T[(c_slot . grad)(J vin)] on nf fields from field0."""
        check(self.lib.nsb_sem_convect(self.h, int(slot), vin.basis.h, vin.col, vout.basis.h, vout.col,
                                       int(field0), int(nf), float(scale), int(accumulate)))

    def convect_t(self, slot: int, vin: nek_dvector, vout: nek_dvector, field0: int = 0, nf: int = 1,
                  scale: float = 1.0, accumulate: bool = False):
        """Exact transpose of ``convect`` on the local points (convective term of the adjoint stepper)."""
        check(self.lib.nsb_sem_convect_t(self.h, int(slot), vin.basis.h, vin.col, vout.basis.h, vout.col,
                                         int(field0), int(nf), float(scale), int(accumulate)))

    def bdf_ext(self, bf: nek_dvector, e1: nek_dvector, e2: nek_dvector, vlag, ab, bd, rho_over_dt: float,
                field0: int = 0, nf: int = 1):
        """makextp + makebdfp in one pass; all vectors are columns of the same basis, vlag[0] = current."""
        cols = (C.c_int * len(vlag))(*[v.col for v in vlag])
        abv, bdv = _f64(ab), _f64(bd)
        assert abv.size >= 3 and bdv.size >= len(vlag) + 1
        check(self.lib.nsb_sem_bdf_ext(self.h, bf.basis.h, bf.col, e1.col, e2.col, cols, len(vlag), int(field0),
                                       int(nf), _dp(abv), _dp(bdv), float(rho_over_dt)))

    def close(self):
        if self.h:
            self.lib.nsb_sem_destroy(self.h)
            self.h = None


class LinearOperator:
    """abstract_linop: ``matvec(vec_in, vec_out)`` (core/linear_operators.f90:17-23)."""

    def __init__(self, lib, h, keep=None):
        self.lib, self.h, self._keep = lib, h, keep

    def matvec(self, vec_in: nek_dvector, vec_out: nek_dvector):
        check(self.lib.nsb_op_apply(self.h, vec_in.basis.h, vec_in.col, vec_out.basis.h, vec_out.col))

    def count(self) -> int:
        n = C.c_int64()
        check(self.lib.nsb_op_count(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            self.lib.nsb_op_destroy(self.h)
            self.h = None


def dealias_matrices(N: int, lxd: int):
    """Host only: (zd, wd, J[lxd, N+1], Dg[lxd, lxd]) of the dealiased convection (Gauss-Legendre fine mesh)."""
    zd, wd = np.zeros(lxd), np.zeros(lxd)
    J, Dg = np.zeros((lxd, N + 1)), np.zeros((lxd, lxd))
    check(_capi.load().nsb_dealias_matrices(int(N), int(lxd), _dp(zd), _dp(wd), _dp(J), _dp(Dg)))
    return zd, wd, J, Dg


def sem_operator(sem: Sem, nfields: int, alpha: float, beta: float, h1: float, h2: float,
                 conv=None) -> LinearOperator:
    """out = alpha*in + beta * binvm1*mask*dssum(h1 A in + h2 B in [+ B (c.grad) in])."""
    h = C.c_void_p()
    cs = [None, None, None]
    keep = []
    if conv is not None:
        for i, c in enumerate(conv):
            a = _f64(c)
            keep.append(a)
            cs[i] = _dp(a)
    check(sem.lib.nsb_op_create_sem(sem.h, nfields, alpha, beta, h1, h2, cs[0], cs[1], cs[2], C.byref(h)))
    return LinearOperator(sem.lib, h, keep)


def stepper_operator(sem: Sem, layout: Layout, nfields: int, slot: int, kappa: float, dt: float, nsteps: int,
                     rho: float = 1.0, tol: float = 1e-12, maxit: int = 1000, adjoint: bool = False) -> LinearOperator:
    """Device time-stepper operator: nsteps BDF3/EXT3 advection-diffusion steps from a cold start (the structure
    of exponential_prop%matvec, core/linear_operators.f90:225-274, for Nek's scalar step); slot -1 = no flow."""
    h = C.c_void_p()
    create = sem.lib.nsb_op_create_stepper_adjoint if adjoint else sem.lib.nsb_op_create_stepper
    check(create(sem.h, layout.h, int(nfields), int(slot), float(kappa), float(rho),
                 float(dt), int(nsteps), float(tol), int(maxit), C.byref(h)))
    return LinearOperator(sem.lib, h, keep=(sem, layout))


def ns_stepper_operator(sem: Sem, layout: Layout, base: nek_dvector | None, nu: float, dt: float, nsteps: int,
                        tol_v: float = 1e-12, tol_p: float = 1e-12, maxit: int = 4000,
                        mean_free: bool = True, precond: int = 1, adjoint: bool = False) -> LinearOperator:
    """exponential_prop%matvec for the linearised incompressible Navier-Stokes equations on the device
    (core/linear_operators.f90:225-274): nsteps pressure-coupled BDF/EXT steps of the P_N - P_N-2 splitting from a
    cold start.  Layout: fields 0..dim-1 velocity, field dim pressure; base = the base flow (None: Stokes);
    adjoint = True: %rmatvec, the same stepper on the adjoint equations."""
    h = C.c_void_p()
    create = sem.lib.nsb_op_create_ns_stepper_adjoint if adjoint else sem.lib.nsb_op_create_ns_stepper
    check(create(sem.h, layout.h, base.basis.h if base is not None else None,
                 base.col if base is not None else 0, float(nu), float(dt), int(nsteps),
                 float(tol_v), float(tol_p), int(maxit), int(bool(mean_free)), int(precond), C.byref(h)))
    return LinearOperator(sem.lib, h, keep=(sem, layout))


def ns_set_orbit(op: LinearOperator, orbit: Basis | None, col0: int = 0, stride: int = 1):
    """Time-periodic base flow: step n of the Navier-Stokes stepper linearises about column col0 + (n-1) stride of
    ``orbit`` (the stored uor / vor / wor of core/linear_operators.f90:254-275); None: the steady base flow again."""
    check(op.lib.nsb_op_ns_set_orbit(op.h, orbit.h if orbit is not None else None, int(col0), int(stride)))
    op._keep = (op._keep, orbit)


def ns_iterations(op: LinearOperator):
    """(Helmholtz iterations, pressure iterations) a Navier-Stokes stepper operator has spent so far."""
    a, b = C.c_int64(), C.c_int64()
    check(op.lib.nsb_op_ns_iterations(op.h, C.byref(a), C.byref(b)))
    return a.value, b.value


def pressure_matrices(N: int):
    """Host-only: (z2, w2, I12, D12) of the lx2 = N - 1 Gauss-Legendre pressure mesh (ixm12, dxm12)."""
    from ._capi import load
    lib = load()
    l1, l2 = N + 1, N - 1
    z2, w2, I12, D12 = np.empty(l2), np.empty(l2), np.empty((l2, l1)), np.empty((l2, l1))
    check(lib.nsb_pressure_matrices(int(N), _dp(z2), _dp(w2), _dp(I12), _dp(D12)))
    return z2, w2, I12, D12


def fdm_matrices(N: int):
    """Host-only: (S, lam) of the element-wise pressure solves, E^ S = M^ S diag(lam), S^T M^ S = 1."""
    l2 = N - 1
    S, lam = np.empty((l2, l2)), np.empty(l2)
    check(_capi.load().nsb_fdm_matrices(int(N), _dp(S), _dp(lam)))
    return S, lam


def compose_operators(layout: Layout, outer: LinearOperator, inner: LinearOperator) -> LinearOperator:
    """out = outer(inner(in)) -- e.g. transient_growth_map = adjoint(forward(q)) (core/matvec.f90:478-495)."""
    h = C.c_void_p()
    check(layout.lib.nsb_op_create_compose(layout.h, outer.h, inner.h, C.byref(h)))
    return LinearOperator(layout.lib, h, keep=(outer, inner))


def axpby_operator(layout: Layout, A: LinearOperator | None, B: LinearOperator | None, alpha: float,
                   beta: float) -> LinearOperator:
    """out = alpha A(in) + beta B(in), None = identity (LightKrylov's axpby_linop / identity_linop,
    core/linear_operators.f90:364-403): newton_linearized_map = axpby_operator(lay, A, None, 1, -1)
    (core/matvec.f90:520-541), ts_force_sensitivity_map = axpby_operator(lay, None, A_adj, 1, -1) (:499-516)."""
    h = C.c_void_p()
    check(layout.lib.nsb_op_create_axpby(layout.h, A.h if A is not None else None, B.h if B is not None else None,
                                         float(alpha), float(beta), C.byref(h)))
    return LinearOperator(layout.lib, h, keep=(layout, A, B))


def frechet_operator(layout: Layout, F: LinearOperator, base: nek_dvector, order: int = 2,
                     epsilon_base: float | None = None) -> LinearOperator:
    """forward_finite_difference_map (core/matvec.f90:246-379): finite-difference approximation of the Frechet
    derivative of the nonlinear map F about ``base`` (read at every application), findiff_order 2 or 4."""
    h = C.c_void_p()
    check(layout.lib.nsb_op_create_frechet_fd(layout.h, F.h, base.basis.h, base.col, int(order), C.byref(h)))
    if epsilon_base is not None:
        check(layout.lib.nsb_op_frechet_set_epsilon(h, float(epsilon_base)))
    return LinearOperator(layout.lib, h, keep=(layout, F, base.basis))


def host_operator(layout: Layout, fn, linear: bool = False) -> LinearOperator:
    """Wrap a host matvec ``fn(fields_in, time_in) -> (fields_out, time_out)`` (the reference's
    time-stepper lives on the host); vectors cross PCIe around every call.  ``linear=True`` declares
    M(a x) = a M(x): the Arnoldi loop then hands over un-normalised vectors while its last sweep still runs."""
    lens = layout.field_len

    hold = {}

    def tramp(user, pin, tin, pout, tout):
        try:
            fin = [np.ctypeslib.as_array(pin[i], shape=(lens[i],)) if lens[i] else np.empty(0)
                   for i in range(len(lens))]
            res, t = fn(fin, tin)
            for i in range(len(lens)):
                a = res[i]
                if lens[i] == 0:
                    continue
                if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous \
                        and a.size == lens[i]:
                    hold[i] = a                       # zero-copy: hand the caller's array over
                    pout[i] = a.ctypes.data_as(c_double_p)
                else:
                    np.ctypeslib.as_array(pout[i], shape=(lens[i],))[:] = np.asarray(a).ravel()
            tout[0] = t
            return 0
        except Exception:  # noqa: BLE001  (must not propagate through the C frame)
            import traceback
            traceback.print_exc()
            return 1

    cb = _capi.HOST_MATVEC(tramp)
    h = C.c_void_p()
    check(layout.lib.nsb_op_create_host(layout.h, cb, None, C.byref(h)))
    if linear:
        check(layout.lib.nsb_op_set_linear(h, 1))
    return LinearOperator(layout.lib, h, keep=cb)


# ------------------------------------------------------------------------------------------------
# solvers
# ------------------------------------------------------------------------------------------------
def arnoldi_factorization(Q: Basis, H: np.ndarray, mstart: int, mend: int, ksize: int,
                          op: LinearOperator, orth_mode: int = ORTH_CGS2):
    """core/krylov_decomposition.f90:2 -- 1-based mstart/mend like the reference; H is a Fortran-
    ordered (ksize+1, ksize) array updated in place."""
    if ksize == 0:
        raise ValueError('Krylov base dimension == 0! Increase it.')
    if not (H.flags.f_contiguous and H.dtype == np.float64):
        raise ValueError('H must be a Fortran-ordered float64 array')
    check(Q.lib.nsb_arnoldi(Q.h, op.h, mstart - 1, mend - 1, orth_mode, _dp(H), H.shape[0]))


def arnoldi_passes(Q: Basis, mstart: int, mend: int, orth_mode: int) -> np.ndarray:
    """Projection passes per step (1-based mstart..mend) of the last device-resident factorisation."""
    out = np.zeros(mend - mstart + 1, dtype=np.int32)
    check(Q.lib.nsb_arnoldi_passes(Q.h, mstart - 1, mend - 1, orth_mode, out.ctypes.data_as(c_int_p)))
    return out


def build_id() -> str:
    return _capi.load().nsb_build_id().decode()


def eig(A: np.ndarray):
    set_lapack_from_scipy()
    n = A.shape[0]
    Af = np.asfortranarray(A, dtype=np.float64)
    vecs = np.zeros((n, n), dtype=np.complex128, order='F')
    vals = np.zeros(n, dtype=np.complex128)
    check(_capi.load().nsb_eig(_dp(Af), n, n, vecs.ctypes.data_as(c_double_p), vals.ctypes.data_as(c_double_p)))
    return vecs, vals


def schur(A: np.ndarray):
    set_lapack_from_scipy()
    n = A.shape[0]
    T = np.array(A, dtype=np.float64, order='F')
    vecs = np.zeros((n, n), order='F')
    vals = np.zeros(n, dtype=np.complex128)
    check(_capi.load().nsb_schur(_dp(T), n, n, _dp(vecs), vals.ctypes.data_as(c_double_p)))
    return T, vecs, vals


def ordschur(T: np.ndarray, Qm: np.ndarray, selected):
    set_lapack_from_scipy()
    n = T.shape[0]
    Tf = np.array(T, dtype=np.float64, order='F')
    Qf = np.array(Qm, dtype=np.float64, order='F')
    sel = np.ascontiguousarray(selected, dtype=np.int32)
    check(_capi.load().nsb_ordschur(_dp(Tf), n, _dp(Qf), n, sel.ctypes.data_as(c_int_p), n))
    return Tf, Qf


def lstsq(A: np.ndarray, b: np.ndarray):
    set_lapack_from_scipy()
    m, n = A.shape
    Af = np.asfortranarray(A, dtype=np.float64)
    bb = _f64(b)
    x = np.zeros(n)
    check(_capi.load().nsb_lstsq(_dp(Af), m, m, n, _dp(bb), _dp(x)))
    return x


def select_eigenvalues(vals: np.ndarray, delta: float, nev: int):
    v = np.ascontiguousarray(vals, dtype=np.complex128)
    n = v.size
    sel = np.zeros(n, dtype=np.int32)
    cnt = C.c_int()
    check(_capi.load().nsb_select_eigenvalues(sel.ctypes.data_as(c_int_p), C.byref(cnt),
                                              v.ctypes.data_as(c_double_p), delta, nev, n))
    return sel.astype(bool), cnt.value


def schur_condensation(mstart: int, H: np.ndarray, Q: Basis, ksize: int, schur_del: float,
                       schur_tgt: int) -> int:
    """core/eigensolvers.f90:363 -- returns the new 1-based mstart."""
    set_lapack_from_scipy()
    m = C.c_int(mstart - 1)
    check(Q.lib.nsb_schur_condensation(Q.h, C.byref(m), _dp(H), H.shape[0], ksize, schur_del, schur_tgt))
    return m.value + 1


class KSResult:
    def __init__(self, vals, vecs, residual, cnt, schur_cnt, H):
        self.vals, self.vecs, self.residual, self.cnt, self.schur_cnt, self.H = \
            vals, vecs, residual, cnt, schur_cnt, H


def krylov_schur(Q: Basis, op: LinearOperator, k_dim: int = 100, schur_tgt: int = 2,
                 eigen_tol: float = 1e-6, schur_del: float = 0.10, orth_mode: int = ORTH_CGS2,
                 max_restarts: int = 200) -> KSResult:
    """core/eigensolvers.f90:120 -- Q[0] holds the unit-norm seed (defaults: core/main.f90:9-31)."""
    set_lapack_from_scipy()
    H = np.zeros((k_dim + 1, k_dim), order='F')
    vals = np.zeros(k_dim, dtype=np.complex128)
    vecs = np.zeros((k_dim, k_dim), dtype=np.complex128, order='F')
    res = np.zeros(k_dim)
    cnt, scnt = C.c_int(), C.c_int()
    check(Q.lib.nsb_krylov_schur(Q.h, op.h, k_dim, schur_tgt, eigen_tol, schur_del, orth_mode,
                                 max_restarts, _dp(H), k_dim + 1, vals.ctypes.data_as(c_double_p),
                                 vecs.ctypes.data_as(c_double_p), _dp(res), C.byref(cnt), C.byref(scnt)))
    return KSResult(vals, vecs, res, cnt.value, scnt.value, H)


def eigs(Q: Basis, op: LinearOperator, k_dim: int, nev: int, tol: float, orth_mode: int = ORTH_CGS2):
    """LightKrylov-style step-wise eigensolver (call site core/linear_stab.f90:66); Q[0] = unit seed.
    Returns (vals[k], vecs[k, k], residual[k], k, nconv, H)."""
    set_lapack_from_scipy()
    H = np.zeros((k_dim + 1, k_dim), order='F')
    vals = np.zeros(k_dim, dtype=np.complex128)
    vecs = np.zeros((k_dim, k_dim), dtype=np.complex128, order='F')
    res = np.zeros(k_dim)
    ku, nc = C.c_int(), C.c_int()
    check(Q.lib.nsb_eigs(Q.h, op.h, k_dim, nev, tol, orth_mode, _dp(H), k_dim + 1,
                         vals.ctypes.data_as(c_double_p), vecs.ctypes.data_as(c_double_p), _dp(res),
                         C.byref(ku), C.byref(nc)))
    k = ku.value
    return vals[:k].copy(), vecs[:k, :k].copy(), res[:k].copy(), k, nc.value, H


def newton_krylov(Q: Basis, fop: LinearOperator, jop: LinearOperator, q: nek_dvector, f: nek_dvector,
                  dq: nek_dvector, maxiter_newton: int, maxiter_gmres: int, ksize: int, tol: float,
                  orth_mode: int = ORTH_CGS2):
    """core/newton_krylov.f90:1 -- q is updated in place; returns (residual history, linear-solver calls).
    fop = forward map F (may be nonlinear), jop = its linearisation about the current q (the owner of the two
    host callbacks re-linearises when F is called); f and dq are work vectors of one basis."""
    set_lapack_from_scipy()
    assert f.basis is dq.basis
    hist = np.zeros(maxiter_newton)
    it, calls = C.c_int(), C.c_int()
    check(Q.lib.nsb_newton_krylov(Q.h, fop.h, jop.h, q.basis.h, q.col, f.basis.h, f.col, dq.col, maxiter_newton,
                                  maxiter_gmres, ksize, tol, orth_mode, C.byref(it), _dp(hist), C.byref(calls)))
    return hist[:it.value].copy(), calls.value


def ritz_vector(Q: Basis, k: int, y, out_re: nek_dvector, out_im: nek_dvector, normalize: bool = True):
    """fp = Q(:,1:k) y for complex y (core/eigensolvers.f90:565-585): real part -> out_re, imaginary part ->
    out_im (columns of the same basis), both scaled by 1/sqrt(|Re|^2 + |Im|^2); returns (|Re|, |Im|)."""
    yc = np.ascontiguousarray(y, dtype=np.complex128).ravel()
    assert out_re.basis is out_im.basis and yc.size >= k
    ar, ai = C.c_double(), C.c_double()
    check(Q.lib.nsb_ritz_vector(Q.h, int(k), yc.ctypes.data_as(c_double_p), out_re.basis.h, out_re.col, out_im.col,
                                int(normalize), C.byref(ar), C.byref(ai)))
    return ar.value, ai.value


def outpost_ks(Q: Basis, sem: Sem, k: int, vals, vecs, converged: int, work: Basis, speriod: float,
               maxmodes: int = 20, spurious_limit: float = 1.1, on_mode=None):
    """The device side of outpost_ks (core/eigensolvers.f90:553-618): for each of the first ``converged`` Ritz pairs
    assemble fp = Q(:,1:k) vecs(:,i) (real part -> work[0], imaginary part -> work[1]), measure |Re|, |Im| and
    norm_grad of both parts, skip the pair as spurious when either norm_grad exceeds 1.1 (:591-594), otherwise scale
    both parts by 1/sqrt(|Re|^2 + |Im|^2) (:584-585, 607-614) and hand them to ``on_mode(outp, work[0], work[1])``
    (the reference writes the field files here).  At most maxmodes modes are kept (:556-559).  Returns one dict per
    converged pair: index, kept, norms, norm_grads, sigma / omega = log_transform(val) / speriod (:598-603)."""
    vals = np.asarray(vals, dtype=np.complex128)
    vecs = np.asarray(vecs, dtype=np.complex128)
    out, outp = [], 0
    for i in range(int(converged)):
        if outp >= maxmodes:
            out.append(dict(index=i, kept=False, reason='maxmodes'))
            continue
        ar, ai = ritz_vector(Q, k, vecs[:k, i], work[0], work[1], normalize=False)
        gr, gi = sem.norm_grad(work[0]), sem.norm_grad(work[1])
        lam = np.log(vals[i])
        if vals[i].imag == 0:
            lam = complex(lam.real, 0.0)                           # log_transform, core/eigensolvers.f90:860-869
        rec = dict(index=i, norms=(ar, ai), norm_grads=(gr, gi), sigma=lam.real / speriod, omega=lam.imag / speriod)
        if gr > spurious_limit or gi > spurious_limit:
            rec.update(kept=False, reason='spurious')
            out.append(rec)
            continue
        beta = 1.0 / np.sqrt(ar * ar + ai * ai)
        work[0].scal(beta)
        work[1].scal(beta)
        outp += 1
        rec.update(kept=True, outp=outp)
        if on_mode is not None:
            on_mode(outp, work[0], work[1])
        out.append(rec)
    return out


def set_linear_solver(sem: Sem, base: nek_dvector, T: float, ctarg: float = 0.5):
    """The time step of the linearised solver as set_linear_solver chooses it (core/linear_stab.f90:214-236): target
    CFL above 1 is limited to 0.5; dt = ctarg / compute_cfl(base, 1); nsteps = ceiling(T / dt); dt = T / nsteps.
    Returns (dt, nsteps, cfl at that dt)."""
    import math
    if ctarg > 1.0:
        ctarg = 0.5
    dt = ctarg / sem.compute_cfl(base, 1.0)
    nsteps = int(math.ceil(T / dt))
    dt = T / nsteps
    return dt, nsteps, sem.compute_cfl(base, dt)


def svd(A):
    """Thin SVD through the injected dgesvd: returns (U, S, V) with A = U diag(S) V^T."""
    set_lapack_from_scipy()
    m, n = A.shape
    r = min(m, n)
    Af = np.asfortranarray(A, dtype=np.float64)
    U = np.zeros((m, r), order='F')
    S = np.zeros(r)
    V = np.zeros((n, r), order='F')
    check(_capi.load().nsb_svd(_dp(Af), m, m, n, _dp(U), _dp(S), _dp(V)))
    return U, S, V


def svds(U: Basis, V: Basis, op: LinearOperator, op_adj: LinearOperator, k_dim: int, nev: int, tol: float,
         orth_mode: int = ORTH_CGS2):
    """LightKrylov-style step-wise singular-value solver (call site core/linear_stab.f90:112);
    U[0] = unit seed, op_adj = A%rmatvec.  Returns (sigma[k], uvecs[k, k], vvecs[k, k], residual[k],
    k, nconv, B)."""
    set_lapack_from_scipy()
    B = np.zeros((k_dim + 1, k_dim), order='F')
    sigma = np.zeros(k_dim)
    uv = np.zeros((k_dim, k_dim), order='F')
    vv = np.zeros((k_dim, k_dim), order='F')
    res = np.zeros(k_dim)
    ku, nc = C.c_int(), C.c_int()
    check(U.lib.nsb_svds(U.h, V.h, op.h, op_adj.h, k_dim, nev, tol, orth_mode, _dp(B), k_dim + 1, _dp(sigma),
                         _dp(uv), _dp(vv), _dp(res), C.byref(ku), C.byref(nc)))
    k = ku.value
    return sigma[:k].copy(), uv[:k, :k].copy(), vv[:k, :k].copy(), res[:k].copy(), k, nc.value, B


def ts_gmres(Q: Basis, op: LinearOperator, rhs: nek_dvector, sol: nek_dvector, maxiter: int,
             ksize: int, tol: float, orth_mode: int = ORTH_CGS2):
    """core/newton_krylov.f90:170 -- returns (residual history, calls)."""
    set_lapack_from_scipy()
    hist = np.zeros(maxiter)
    calls, nh = C.c_int(), C.c_int()
    check(Q.lib.nsb_ts_gmres(Q.h, op.h, rhs.basis.h, rhs.col, sol.basis.h, sol.col, maxiter, ksize,
                             tol, orth_mode, C.byref(calls), _dp(hist), C.byref(nh)))
    return hist[:nh.value].copy(), calls.value


# ------------------------------------------------------------------------------------------------
# restart wire formats through the C ABI (the Python module `checkpoint` holds the writers / spectra files)
# ------------------------------------------------------------------------------------------------
def hessenberg_write(path, H: np.ndarray, k: int):
    Hf = np.asfortranarray(H, dtype=np.float64)
    check(_capi.load().nsb_hessenberg_write(str(path).encode(), _dp(Hf), Hf.shape[0], int(k)))


def hessenberg_read(path, k_dim: int, mstart: int) -> np.ndarray:
    H = np.zeros((k_dim + 1, k_dim), order='F')
    check(_capi.load().nsb_hessenberg_read(str(path).encode(), int(k_dim), int(mstart), _dp(H), k_dim + 1))
    return H


def fld_read_into(vec: nek_dvector, path, nel_local: int, ufield0: int = 0, pfield: int = -1, tfield: int = -1,
                  lglel=None) -> float:
    """Nek field file -> device vector (load_fld + the copies of load_files, core/IO.f90:58-68); returns the
    header's time."""
    t = C.c_double()
    lg = None if lglel is None else np.ascontiguousarray(lglel, dtype=np.int64)
    check(vec.basis.lib.nsb_fld_read_into(vec.basis.h, vec.col, str(path).encode(),
                                          None if lg is None else lg.ctypes.data_as(c_i64_p), int(nel_local),
                                          int(ufield0), int(pfield), int(tfield), C.byref(t)))
    return t.value


def restart_load(Q: Basis, directory, session: str, mstart: int, k_dim: int, nel_local: int, ufield0: int = 0,
                 pfield: int = -1, tfield: int = -1, lglel=None):
    """The restart branch of krylov_schur (core/eigensolvers.f90:240-285); returns (H, next 1-based mstart)."""
    H = np.zeros((k_dim + 1, k_dim), order='F')
    nxt = C.c_int()
    lg = None if lglel is None else np.ascontiguousarray(lglel, dtype=np.int64)
    check(Q.lib.nsb_restart_load(Q.h, str(directory).encode(), session.encode(), int(mstart), int(k_dim),
                                 None if lg is None else lg.ctypes.data_as(c_i64_p), int(nel_local), int(ufield0),
                                 int(pfield), int(tfield), _dp(H), k_dim + 1, C.byref(nxt)))
    return H, nxt.value + 1
