"""Host-side mirror of the reference's analysis drivers (core/linear_stab.f90): the sequence of calls
``linear_stability_analysis`` and ``transient_growth_analysis`` make, with every vector operation, the propagators
and the Krylov drivers on the device.  Nothing here computes on the CPU beyond the k x k problems LAPACK solves and
the text files the reference writes.

    prepare_base_flow   -> the caller's base-flow vector (a column of a device basis; checkpoint.read_fld /
                           nsb_fld_read_into load the reference's BF_<session>0.f00001)
    set_linear_solver   -> api.set_linear_solver (compute_cfl on the device)
    exponential_prop    -> api.ns_stepper_operator (matvec) / adjoint=True (rmatvec), ns_set_orbit for 'periodic'
    prepare_seed        -> the caller's seed in X[0] (seed.seed_noise or a loaded mode), normalised here
    eigs / svds         -> api.eigs / api.svds
    outpost_*           -> checkpoint.write_spectrum / write_singvals, k_matmul (get_vec) per mode
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import api, checkpoint

__all__ = ['linear_stability_analysis', 'transient_growth_analysis']


def _propagators(sem, layout, base, nu, T, ctarg, mode, orbit, solver, want_adjoint, want_forward=True):
    dt, nsteps, cfl = api.set_linear_solver(sem, base, T, ctarg)                       # core/linear_stab.f90:220-236
    ops = {}
    for name, adj in (('forward', False), ('adjoint', True)):
        if (adj and not want_adjoint) or (not adj and not want_forward):
            continue
        op = api.ns_stepper_operator(sem, layout, base, nu, dt, nsteps, adjoint=adj, **solver)
        if 'periodic' in mode:                                                          # Floquet: the stored orbit
            if orbit is None or orbit.ncols < nsteps:
                op.close()
                raise ValueError(f"'periodic' needs the base-flow orbit as a basis of at least nsteps = {nsteps} columns")
            api.ns_set_orbit(op, orbit, 0, 1)
        ops[name] = op
    return dt, nsteps, cfl, ops


def _prepare_seed(X):
    alpha = api.k_norm(X[0])                                                            # core/linear_stab.f90:287-289
    if not alpha > 0.0:
        raise ValueError('prepare_seed: the seed in X[0] has zero norm')
    X[0].scal(1.0 / alpha)


def linear_stability_analysis(sem, layout, X, base, nu: float, T: float, solver_mode: str = 'steady',
                              solver_type: str = 'direct', k_dim: int | None = None, schur_tgt: int = 2,
                              eigen_tol: float = 1e-6, ctarg: float = 0.5, maxmodes: int = 20, orbit=None,
                              work=None, outdir=None, on_mode=None, solver: dict | None = None):
    """linear_stability_analysis(solver_mode, solver_type) (core/linear_stab.f90:12-80).  X: basis of k_dim + 1 columns
    with the seed in X[0]; base: the base flow; nu = 1 / Re; T = fintim.  solver_type 'direct' / 'forward' iterates on
    exponential_prop%matvec, 'adjoint' / 'backward' on %rmatvec (eigs(..., transpose=.true.), :66-67).  Writes
    Spectrum_H<evop>.dat and Spectrum_NS<evop>.dat into outdir (:69-73) and hands the first maxmodes Ritz vectors
    (get_vec of the real and of the imaginary part, :352-375, into work[0] / work[1]) to on_mode(i, re, im).
    Returns a dict with the Ritz values, log(vals) / T, residuals, the Krylov dimension reached, dt and nsteps."""
    if 'steady' not in solver_mode and 'periodic' not in solver_mode:
        raise ValueError("solver_mode is 'steady' or 'periodic'")
    adjoint = 'adjoint' in solver_type or 'backward' in solver_type
    if not adjoint and not ('direct' in solver_type or 'forward' in solver_type):
        raise ValueError("solver_type is 'direct' / 'forward' or 'adjoint' / 'backward'")
    evop = 'a' if adjoint else 'd'
    k_dim = X.ncols - 1 if k_dim is None else int(k_dim)
    if X.ncols < k_dim + 1:
        raise ValueError(f'X needs k_dim + 1 = {k_dim + 1} columns')
    dt, nsteps, cfl, ops = _propagators(sem, layout, base, nu, T, ctarg, solver_mode, orbit, solver or {},
                                        want_adjoint=adjoint, want_forward=not adjoint)
    A = ops['adjoint' if adjoint else 'forward']
    try:
        _prepare_seed(X)
        vals, vecs, res, k, nconv, H = api.eigs(X, A, k_dim, schur_tgt, eigen_tol)
        vals_ns = checkpoint.log_transform(vals) / T                                    # eigvals = log(eigvals)/A%t, :71
        if outdir is not None:
            out = Path(outdir)
            checkpoint.write_spectrum(out / f'Spectrum_H{evop}.dat', vals, res)
            checkpoint.write_spectrum(out / f'Spectrum_NS{evop}.dat', vals_ns, res)
        if on_mode is not None:
            if work is None or work.ncols < 2:
                raise ValueError('on_mode needs a work basis of two columns')
            for i in range(min(maxmodes, k)):                                           # outpost_eigenvectors, :345-375
                api.k_matmul(work[0], X, np.ascontiguousarray(vecs[:k, i].real), k)
                api.k_matmul(work[1], X, np.ascontiguousarray(vecs[:k, i].imag), k)
                on_mode(i + 1, work[0], work[1])
        return dict(evop=evop, eigvals=vals, eigvals_ns=vals_ns, eigvecs=vecs, residuals=res, k=k, nconv=nconv, H=H,
                    dt=dt, nsteps=nsteps, cfl=cfl, matvecs=A.count())
    finally:
        for op in ops.values():
            op.close()


def transient_growth_analysis(sem, layout, U, V, base, nu: float, T: float, solver_mode: str = 'steady',
                              k_dim: int | None = None, schur_tgt: int = 2, eigen_tol: float = 1e-6,
                              ctarg: float = 0.5, maxmodes: int = 20, orbit=None, work=None, outdir=None,
                              on_mode=None, solver: dict | None = None):
    """transient_growth_analysis(solver_mode) (core/linear_stab.f90:82-119): svds of exponential_prop with its
    rmatvec, the gains sigma**2 (:112) written to Spectrum_Sp.dat (:113), the first maxmodes left / right singular
    vectors (get_vec, :404-425) handed to on_mode(i, u, v).  U: k_dim + 1 columns with the seed in U[0], V: k_dim."""
    if 'steady' not in solver_mode and 'periodic' not in solver_mode:
        raise ValueError("solver_mode is 'steady' or 'periodic'")
    k_dim = U.ncols - 1 if k_dim is None else int(k_dim)
    if U.ncols < k_dim + 1 or V.ncols < k_dim:
        raise ValueError(f'U needs {k_dim + 1} columns, V {k_dim}')
    dt, nsteps, cfl, ops = _propagators(sem, layout, base, nu, T, ctarg, solver_mode, orbit, solver or {},
                                        want_adjoint=True)
    try:
        _prepare_seed(U)
        sig, uv, vv, res, k, nconv, B = api.svds(U, V, ops['forward'], ops['adjoint'], k_dim, schur_tgt, eigen_tol)
        gains = sig ** 2
        if outdir is not None:
            checkpoint.write_singvals(Path(outdir) / 'Spectrum_Sp.dat', gains, res)
        if on_mode is not None:
            if work is None or work.ncols < 2:
                raise ValueError('on_mode needs a work basis of two columns')
            for i in range(min(maxmodes, k)):                                           # outpost_singvectors
                api.k_matmul(work[0], U, np.ascontiguousarray(uv[:k, i]), k)
                api.k_matmul(work[1], V, np.ascontiguousarray(vv[:k, i]), k)
                on_mode(i + 1, work[0], work[1])
        return dict(evop='p', sigma=sig, gains=gains, uvecs=uv, vvecs=vv, residuals=res, k=k, nconv=nconv, B=B,
                    dt=dt, nsteps=nsteps, cfl=cfl, matvecs=(ops['forward'].count(), ops['adjoint'].count()))
    finally:
        for op in ops.values():
            op.close()
