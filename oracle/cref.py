"""ctypes front-end of oracle/_ref/libref_cpu.so (C restatement of the reference CPU path).
TEST INFRASTRUCTURE ONLY -- see oracle/ref_cpu.c."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / '_ref' / 'libref_cpu.so'
_dp = C.POINTER(C.c_double)
_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(['make', '-C', str(HERE), '-s'], check=True)
        _lib = C.CDLL(str(LIB))
        _lib.ref_glsc3.restype = C.c_double
        _lib.ref_k_dot.restype = C.c_double
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def p(a):
    return a.ctypes.data_as(_dp)


def gs_lists(glo):
    """CSR lists (offsets int64, indices int32) of all points per unique global node."""
    flat = glo.ravel()
    order = np.argsort(flat, kind='stable')
    sg = flat[order]
    starts = np.flatnonzero(np.r_[True, sg[1:] != sg[:-1]])
    off = np.r_[starts, flat.size].astype(np.int64)
    return off, order.astype(np.int32)


def num_threads():
    return load().ref_num_threads()


def set_num_threads(n):
    load().ref_set_num_threads(C.c_int(int(n)))


def axhelm3d(u, g, bm1, D, h1, h2):
    lib = load()
    nel, lx = u.shape[0], u.shape[-1]
    out = np.empty_like(u)
    Df = np.asfortranarray(D)
    lib.ref_axhelm3d(p(out), p(np.ascontiguousarray(u)), p(np.ascontiguousarray(g)),
                     p(np.ascontiguousarray(bm1)), p(Df), C.c_int(lx), C.c_int64(nel),
                     C.c_double(h1), C.c_double(h2))
    return out


def dssum(u, off, idx):
    lib = load()
    out = np.ascontiguousarray(u).copy()
    lib.ref_dssum(p(out), off.ctypes.data_as(C.POINTER(C.c_int64)), idx.ctypes.data_as(C.POINTER(C.c_int32)),
                  C.c_int64(off.size - 1))
    return out


def update_hessenberg(Q, f, bm1s, npts, ncomp, k):
    """Q: (ncols, n) C-contiguous rows = vectors; f modified in place; returns h[0..k]."""
    lib = load()
    n = npts * ncomp
    wrk = np.empty(n)
    h = np.zeros(k + 1)
    lib.ref_update_hessenberg(p(Q), C.c_int64(Q.shape[1]), p(f), p(wrk), p(bm1s), C.c_int64(npts),
                              C.c_int(ncomp), C.c_int(k), p(h))
    return h


def arnoldi(Q, H, mstart, mend, bm1s, g, bm1, binv, mask, D, lx, nel, ncomp, off, idx, h1, h2, alpha, beta):
    """Q: (ncols, n) rows = vectors, H: Fortran (ldh, k) array; 0-based inclusive steps."""
    lib = load()
    n = Q.shape[1]
    wrk, tmp = np.empty(n), np.empty(n // ncomp)
    Df = np.asfortranarray(D)
    lib.ref_arnoldi(p(Q), C.c_int64(n), p(H), C.c_int(H.shape[0]), C.c_int(mstart), C.c_int(mend), p(wrk),
                    p(tmp), p(bm1s), p(g), p(bm1), p(binv), p(mask), p(Df), C.c_int(lx), C.c_int64(nel),
                    C.c_int(ncomp), off.ctypes.data_as(C.POINTER(C.c_int64)),
                    idx.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(off.size - 1), C.c_double(h1),
                    C.c_double(h2), C.c_double(alpha), C.c_double(beta))
