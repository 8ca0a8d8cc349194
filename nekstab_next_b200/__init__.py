"""nekstab_next_b200 -- B200-native Arnoldi / Newton-Krylov hot path of nekStab.

The compute path is libnekstab_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/nekstab_b200.h).  This package is the thin host-side mirror of the reference's
vector / operator / solver interface used by the tests and the benchmark; importing it never
falls back to a CPU implementation.
"""
from ._capi import NsbError, load, LIB_PATH  # noqa: F401
from .api import (  # noqa: F401
    ORTH_MGS2_REF, ORTH_CGS2, ORTH_DGKS, Context, Layout, Basis, nek_dvector, Sem, LinearOperator,
    sem_operator, dealias_matrices, host_operator, compose_operators, axpby_operator, frechet_operator, stepper_operator, ns_stepper_operator, ns_iterations, ns_set_orbit, pressure_matrices, fdm_matrices, gll, k_dot, k_norm, k_normalize, k_cmult, k_add2, k_sub2, k_sub3,
    k_zero, k_copy, k_matmul, orthonormalize, arnoldi_factorization, eig, schur, ordschur, lstsq,
    select_eigenvalues, schur_condensation, krylov_schur, arnoldi_passes, build_id, hessenberg_write, hessenberg_read, fld_read_into, restart_load, eigs, newton_krylov, ritz_vector, outpost_ks, set_linear_solver, svd, svds, ts_gmres, set_lapack_from_scipy, KSResult,
)
from . import mesh, seed, checkpoint, linear_stab  # noqa: F401

__all__ = [n for n in dir() if not n.startswith('_')]
