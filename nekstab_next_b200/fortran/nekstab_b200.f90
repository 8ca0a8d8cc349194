!> ISO_C_BINDING layer over include/nekstab_b200.h for the nekStab host code.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (SURVEY.md
!! section 0.5).  The file is deliberately declarative -- bind(C) interfaces plus thin type-bound
!! wrappers -- so that the only logic that cannot be tested here is argument marshalling.
!! INTEGRATION.md shows where each piece replaces reference code.
!!
!! Replaces / mirrors (paths relative to the nekStab repository root):
!!   type nek_dvector          <- real_nek_vector (core/nek_vectors.f90:20-31),
!!                                krylov_vector   (core/krylov_subspace.f90:12-17)
!!   k_dot, k_norm, ...        <- core/krylov_subspace.f90:26-209
!!   arnoldi_factorization(Q, H, mstart, mend, ksize)   <- core/krylov_decomposition.f90:2   (same signature)
!!   update_hessenberg_matrix(H, f, q, k)               <- core/krylov_decomposition.f90:103 (same signature)
!!   k_matmul(dq, Q, yvec, k)                           <- core/krylov_subspace.f90:163      (same signature)
!!   schur_condensation(mstart, H, Q, ksize)            <- core/eigensolvers.f90:363         (same signature)
!!   ts_gmres(rhs, sol, maxiter, ksize, calls)          <- core/newton_krylov.f90:170        (same signature)
!!   krylov_schur (driver part)                         <- core/eigensolvers.f90:120
!! The reference routines find the linear operator through `matvec` (core/matvec.f90:56) and their tolerances
!! in the NEKSTAB commons; here the operator is the module variable nsb_current_op (set with nsb_set_operator
!! where the reference calls prepare_linearized_solver) and schur_del / schur_tgt / the GMRES tolerance are the
!! module variables below, so the CALL SITES core/eigensolvers.f90:297,318 and core/newton_krylov.f90:125,252
!! compile unchanged once Q is declared `type(nek_dvector)` and bound to the device basis (nsb_bind_basis).
module nekstab_b200
   use, intrinsic :: iso_c_binding
   implicit none
   private

   integer(c_int), parameter, public :: NSB_ORTH_MGS2_REF = 0, NSB_ORTH_CGS2 = 1, NSB_ORTH_DGKS = 2
   integer(c_int), parameter, public :: NSB_AXPBY_SKIP_TIME = 1

   !> One per MPI rank: device context, state-vector layout, Krylov basis, work vectors.
   type(c_ptr), save, public :: nsb_ctx = c_null_ptr, nsb_layout = c_null_ptr
   type(c_ptr), save, public :: nsb_Q = c_null_ptr      !< Krylov basis, k_dim+1 columns
   type(c_ptr), save, public :: nsb_work = c_null_ptr   !< pool of stand-alone vectors (f, wrk, sol, dq ...)

   !> The device vector: a (basis, column) pair with the reference's type-bound procedures.
   type, public :: nek_dvector
      type(c_ptr) :: basis = c_null_ptr
      integer(c_int) :: col = 0            !< 0-based column
   contains
      procedure, pass(self), public :: zero => dv_zero
      procedure, pass(self), public :: dot => dv_dot
      procedure, pass(self), public :: norm => dv_norm
      procedure, pass(self), public :: scal => dv_scal
      procedure, pass(self), public :: axpby => dv_axpby
   end type nek_dvector

   !> State the reference keeps in commons / finds through `matvec`: the operator every solver below applies,
   !! the orthogonalisation mode, Schur parameters (core/NEKSTAB: schur_del, schur_tgt) and the GMRES tolerance
   !! (max(param(21), param(22)), core/newton_krylov.f90:232).
   type(c_ptr), save, public :: nsb_current_op = c_null_ptr
   integer(c_int), save, public :: nsb_orth_mode = NSB_ORTH_CGS2
   real(c_double), save, public :: nsb_schur_del = 0.10d0, nsb_gmres_tol = 1.0d-8
   integer, save, public :: nsb_schur_tgt = 2

   !> k_matmul keeps the reference's four-argument form; the three-argument form uses the module basis nsb_Q.
   interface k_matmul
      module procedure k_matmul_q, k_matmul_d
   end interface k_matmul

   public :: nsb_set_operator, nsb_bind_basis
   public :: arnoldi_factorization, update_hessenberg_matrix, schur_condensation, ts_gmres
   public :: nsb_check, nsb_startup, nsb_shutdown, nsb_upload, nsb_download
   public :: nsb_p2p_mailbox_create, nsb_p2p_mailbox_connect
   public :: k_dot, k_norm, k_normalize, k_cmult, k_add2, k_sub2, k_sub3, k_zero, k_copy, k_matmul
   public :: nsb_vec_zero, nsb_vec_copy, nsb_vec_scal, nsb_vec_axpby, nsb_vec_dot   ! for nekstab_b200_lightkrylov
   public :: arnoldi_factorization_d, schur_condensation_d, krylov_schur_d, ts_gmres_d, eigs_d, svds_d
   public :: nsb_ritz_vector, nsb_newton_krylov

   interface
      function nsb_last_error() bind(C, name='nsb_last_error') result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      function nsb_get_unique_id(id) bind(C, name='nsb_get_unique_id') result(ierr)
         import :: c_int, c_char
         character(kind=c_char) :: id(128)
         integer(c_int) :: ierr
      end function
      function nsb_init(device, rank, nranks, id, ctx) bind(C, name='nsb_init') result(ierr)
         import :: c_int, c_char, c_ptr
         integer(c_int), value :: device, rank, nranks
         character(kind=c_char) :: id(128)
         type(c_ptr) :: ctx
         integer(c_int) :: ierr
      end function
      function nsb_p2p_mailbox_create(ctx, halo_bytes, handle) bind(C, name='nsb_p2p_mailbox_create') result(ierr)
         import :: c_int, c_int64_t, c_char, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: halo_bytes
         character(kind=c_char) :: handle(64)
         integer(c_int) :: ierr
      end function
      function nsb_p2p_mailbox_connect(ctx, all_handles) bind(C, name='nsb_p2p_mailbox_connect') result(ierr)
         import :: c_int, c_char, c_ptr
         type(c_ptr), value :: ctx
         character(kind=c_char) :: all_handles(64, *)
         integer(c_int) :: ierr
      end function
      function nsb_finalize(ctx) bind(C, name='nsb_finalize') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int) :: ierr
      end function
      function nsb_layout_create(ctx, nfields, field_len, field_in_dot, time_in_dot, layout) &
         bind(C, name='nsb_layout_create') result(ierr)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: nfields, time_in_dot
         integer(c_int64_t) :: field_len(*)
         integer(c_int) :: field_in_dot(*)
         type(c_ptr) :: layout
         integer(c_int) :: ierr
      end function
      function nsb_layout_set_weight(layout, w) bind(C, name='nsb_layout_set_weight') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: layout
         type(c_ptr) :: w(*)            !< c_loc(bm1s) once per in-dot field
         integer(c_int) :: ierr
      end function
      function nsb_basis_create(layout, ncols, basis) bind(C, name='nsb_basis_create') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: layout
         integer(c_int), value :: ncols
         type(c_ptr) :: basis
         integer(c_int) :: ierr
      end function
      function nsb_basis_destroy(basis) bind(C, name='nsb_basis_destroy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: basis
         integer(c_int) :: ierr
      end function
      function nsb_vec_upload(b, col, fields, time) bind(C, name='nsb_vec_upload') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: col
         type(c_ptr) :: fields(*)       !< c_loc(vx), c_loc(vy), ...
         real(c_double), value :: time
         integer(c_int) :: ierr
      end function
      function nsb_vec_download(b, col, fields, time) bind(C, name='nsb_vec_download') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: col
         type(c_ptr) :: fields(*)
         real(c_double) :: time
         integer(c_int) :: ierr
      end function
      function nsb_vec_zero(b, col) bind(C, name='nsb_vec_zero') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: b
         integer(c_int), value :: col
         integer(c_int) :: ierr
      end function
      function nsb_vec_copy(bd, cd, bs, cs) bind(C, name='nsb_vec_copy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: bd, bs
         integer(c_int), value :: cd, cs
         integer(c_int) :: ierr
      end function
      function nsb_vec_scal(b, col, alpha) bind(C, name='nsb_vec_scal') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: col
         real(c_double), value :: alpha
         integer(c_int) :: ierr
      end function
      function nsb_vec_axpby(bx, cx, alpha, by, cy, beta, flags) bind(C, name='nsb_vec_axpby') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: bx, by
         integer(c_int), value :: cx, cy, flags
         real(c_double), value :: alpha, beta
         integer(c_int) :: ierr
      end function
      function nsb_vec_add2(bp, cp, bq, cq) bind(C, name='nsb_vec_add2') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: bp, bq
         integer(c_int), value :: cp, cq
         integer(c_int) :: ierr
      end function
      function nsb_vec_sub2(bp, cp, bq, cq) bind(C, name='nsb_vec_sub2') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: bp, bq
         integer(c_int), value :: cp, cq
         integer(c_int) :: ierr
      end function
      function nsb_vec_sub3(bp, cp, bq, cq, br, cr) bind(C, name='nsb_vec_sub3') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: bp, bq, br
         integer(c_int), value :: cp, cq, cr
         integer(c_int) :: ierr
      end function
      function nsb_vec_dot(ba, ca, bb, cb, alpha) bind(C, name='nsb_vec_dot') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ba, bb
         integer(c_int), value :: ca, cb
         real(c_double) :: alpha
         integer(c_int) :: ierr
      end function
      function nsb_vec_normalize(b, col, alpha) bind(C, name='nsb_vec_normalize') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: col
         real(c_double) :: alpha
         integer(c_int) :: ierr
      end function
      function nsb_basis_gemv(b, k, y, bout, cout) bind(C, name='nsb_basis_gemv') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b, bout
         integer(c_int), value :: k, cout
         real(c_double) :: y(*)
         integer(c_int) :: ierr
      end function
      function nsb_op_create_host(layout, fn, user, op) bind(C, name='nsb_op_create_host') result(ierr)
         import :: c_int, c_ptr, c_funptr
         type(c_ptr), value :: layout, user
         type(c_funptr), value :: fn
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      function nsb_arnoldi(Q, op, mstart, mend, mode, H, ldh) bind(C, name='nsb_arnoldi') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: Q, op
         integer(c_int), value :: mstart, mend, mode, ldh
         real(c_double) :: H(ldh, *)
         integer(c_int) :: ierr
      end function
      function nsb_set_lapack(dgeev, dgees, dtrsen, dgels) bind(C, name='nsb_set_lapack') result(ierr)
         import :: c_int, c_funptr
         type(c_funptr), value :: dgeev, dgees, dtrsen, dgels
         integer(c_int) :: ierr
      end function
      function nsb_schur_condensation(Q, mstart, H, ldh, ksize, schur_del, schur_tgt) &
         bind(C, name='nsb_schur_condensation') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: Q
         integer(c_int) :: mstart
         integer(c_int), value :: ldh, ksize, schur_tgt
         real(c_double) :: H(ldh, *)
         real(c_double), value :: schur_del
         integer(c_int) :: ierr
      end function
      function nsb_krylov_schur(Q, op, k_dim, schur_tgt, eigen_tol, schur_del, mode, max_restarts, H, ldh, &
                                vals, vecs, residual, cnt, schur_cnt) bind(C, name='nsb_krylov_schur') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: Q, op
         integer(c_int), value :: k_dim, schur_tgt, mode, max_restarts, ldh
         real(c_double), value :: eigen_tol, schur_del
         real(c_double) :: H(ldh, *), residual(*)
         complex(c_double_complex) :: vals(*), vecs(k_dim, *)
         integer(c_int) :: cnt, schur_cnt, ierr
      end function
      function nsb_eigs(Q, op, k_dim, nev, tol, mode, H, ldh, vals, vecs, residual, kused, nconv) &
         bind(C, name='nsb_eigs') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: Q, op
         integer(c_int), value :: k_dim, nev, mode, ldh
         real(c_double), value :: tol
         real(c_double) :: H(ldh, *), residual(*)
         complex(c_double_complex) :: vals(*), vecs(k_dim, *)
         integer(c_int) :: kused, nconv, ierr
      end function
      function nsb_newton_krylov(Q, fop, jop, bq, cq, bw, cf, cdq, maxiter_newton, maxiter_gmres, ksize, tol, mode, &
                                 iters, hist, calls) bind(C, name='nsb_newton_krylov') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: Q, fop, jop, bq, bw
         integer(c_int), value :: cq, cf, cdq, maxiter_newton, maxiter_gmres, ksize, mode
         real(c_double), value :: tol
         integer(c_int) :: iters, calls
         real(c_double) :: hist(*)
         integer(c_int) :: ierr
      end function
      function nsb_ritz_vector(Q, k, y, bout, cre, cim, normalize, alpha_re, alpha_im) &
         bind(C, name='nsb_ritz_vector') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: Q, bout
         integer(c_int), value :: k, cre, cim, normalize
         complex(c_double_complex) :: y(*)
         real(c_double) :: alpha_re, alpha_im
         integer(c_int) :: ierr
      end function
      function nsb_set_lapack_svd(dgesvd) bind(C, name='nsb_set_lapack_svd') result(ierr)
         import :: c_int, c_funptr
         type(c_funptr), value :: dgesvd
         integer(c_int) :: ierr
      end function
      function nsb_svds(U, V, op, op_adj, k_dim, nev, tol, mode, B, ldb, sigma, uvecs, vvecs, residual, &
                        kused, nconv) bind(C, name='nsb_svds') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: U, V, op, op_adj
         integer(c_int), value :: k_dim, nev, mode, ldb
         real(c_double), value :: tol
         real(c_double) :: B(ldb, *), sigma(*), uvecs(k_dim, *), vvecs(k_dim, *), residual(*)
         integer(c_int) :: kused, nconv, ierr
      end function
      function nsb_ts_gmres(Q, op, brhs, crhs, bsol, csol, maxiter, ksize, tol, mode, calls, hist, nhist) &
         bind(C, name='nsb_ts_gmres') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: Q, op, brhs, bsol
         integer(c_int), value :: crhs, csol, maxiter, ksize, mode
         real(c_double), value :: tol
         integer(c_int) :: calls, nhist
         real(c_double) :: hist(*)
         integer(c_int) :: ierr
      end function
   end interface

   !> Device-side operators, dense mirrors and stand-alone kernels (same header, include/nekstab_b200.h).
   !> Not bound here because they serve the C / Python test and bench harness only: nsb_prof_*, nsb_profiler_*,
   !> nsb_timer_*, nsb_flush_l2, nsb_launch_count, nsb_stream, nsb_rank, nsb_version, nsb_gll, nsb_dealias_matrices,
   !> nsb_host_alloc /
   !> nsb_host_free, nsb_host_gs_plan, nsb_host_exchange_plan, nsb_allreduce_host, nsb_orthonormalize_async,
   !> nsb_p2p_enabled, nsb_basis_ncols, nsb_basis_col_ptr, nsb_layout_info, nsb_layout_destroy, nsb_sem_get,
   !> nsb_sem_npts.
   interface
      function nsb_sync(ctx) bind(C, name='nsb_sync') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int) :: ierr
      end function
      function nsb_vec_norm(b, col, alpha) bind(C, name='nsb_vec_norm') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: col
         real(c_double) :: alpha
         integer(c_int) :: ierr
      end function
      function nsb_orthonormalize(b, k, col_w, mode, h, passes) bind(C, name='nsb_orthonormalize') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: k, col_w, mode
         real(c_double) :: h(*)
         integer(c_int) :: passes, ierr
      end function
      function nsb_basis_gram(b, k, G, ldg) bind(C, name='nsb_basis_gram') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: k, ldg
         real(c_double) :: G(ldg, *)
         integer(c_int) :: ierr
      end function
      function nsb_basis_qr(b, k, mode, R, ldr) bind(C, name='nsb_basis_qr') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: k, mode, ldr
         real(c_double) :: R(ldr, *)
         integer(c_int) :: ierr
      end function
      function nsb_basis_rotate(b, k, Z, ldz, rotate_time) bind(C, name='nsb_basis_rotate') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: b
         integer(c_int), value :: k, ldz, rotate_time
         real(c_double) :: Z(ldz, *)
         integer(c_int) :: ierr
      end function
      function nsb_sem_create(ctx, dim, N, nel, x, y, z, mask, glo_num, sem) bind(C, name='nsb_sem_create') &
         result(ierr)
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: dim, N
         integer(c_int64_t), value :: nel
         real(c_double) :: x(*), y(*), z(*), mask(*)
         integer(c_int64_t) :: glo_num(*)
         type(c_ptr) :: sem
         integer(c_int) :: ierr
      end function
      function nsb_sem_destroy(sem) bind(C, name='nsb_sem_destroy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem
         integer(c_int) :: ierr
      end function
      function nsb_sem_setup_exchange(sem) bind(C, name='nsb_sem_setup_exchange') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem
         integer(c_int) :: ierr
      end function
      function nsb_sem_axhelm(sem, bin, cin, bout, cout, field, h1, h2) bind(C, name='nsb_sem_axhelm') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, bin, bout
         integer(c_int), value :: cin, cout, field
         real(c_double), value :: h1, h2
         integer(c_int) :: ierr
      end function
      function nsb_sem_ax(sem, bin, cin, bout, cout, field, h1, h2) bind(C, name='nsb_sem_ax') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, bin, bout
         integer(c_int), value :: cin, cout, field
         real(c_double), value :: h1, h2
         integer(c_int) :: ierr
      end function
      function nsb_sem_dssum(sem, b, col, field) bind(C, name='nsb_sem_dssum') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem, b
         integer(c_int), value :: col, field
         integer(c_int) :: ierr
      end function
      function nsb_sem_col2(sem, b, col, field, which) bind(C, name='nsb_sem_col2') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem, b
         integer(c_int), value :: col, field, which
         integer(c_int) :: ierr
      end function
      function nsb_sem_hmholtz(sem, brhs, crhs, bx, cx, field, h1, h2, tol, maxit, iters, res) &
         bind(C, name='nsb_sem_hmholtz') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, brhs, bx
         integer(c_int), value :: crhs, cx, field, maxit
         real(c_double), value :: h1, h2, tol
         integer(c_int) :: iters
         real(c_double) :: res
         integer(c_int) :: ierr
      end function
      function nsb_sem_hmholtz_vec(sem, brhs, crhs, bx, cx, field0, nf, h1, h2, tol, maxit, iters, res) &
         bind(C, name='nsb_sem_hmholtz_vec') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, brhs, bx
         integer(c_int), value :: crhs, cx, field0, nf, maxit
         real(c_double), value :: h1, h2, tol
         integer(c_int) :: iters(*)
         real(c_double) :: res(*)
         integer(c_int) :: ierr
      end function
      function nsb_sem_dealias_setup(sem, lxd) bind(C, name='nsb_sem_dealias_setup') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem
         integer(c_int), value :: lxd
         integer(c_int) :: ierr
      end function
      function nsb_sem_set_convect(sem, slot, b, col, field0) bind(C, name='nsb_sem_set_convect') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem, b
         integer(c_int), value :: slot, col, field0
         integer(c_int) :: ierr
      end function
      function nsb_sem_convect(sem, slot, bin, cin, bout, cout, field0, nf, scale, accumulate) &
         bind(C, name='nsb_sem_convect') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, bin, bout
         integer(c_int), value :: slot, cin, cout, field0, nf, accumulate
         real(c_double), value :: scale
         integer(c_int) :: ierr
      end function
      function nsb_sem_bdf_ext(sem, b, col_bf, col_e1, col_e2, col_vlag, nbd, field0, nf, ab, bd, rho_over_dt) &
         bind(C, name='nsb_sem_bdf_ext') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, b
         integer(c_int), value :: col_bf, col_e1, col_e2, nbd, field0, nf
         integer(c_int) :: col_vlag(*)
         real(c_double) :: ab(*), bd(*)
         real(c_double), value :: rho_over_dt
         integer(c_int) :: ierr
      end function
      function nsb_op_create_sem(sem, nfields_apply, alpha, beta, h1, h2, cx, cy, cz, op) &
         bind(C, name='nsb_op_create_sem') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem
         integer(c_int), value :: nfields_apply
         real(c_double), value :: alpha, beta, h1, h2
         type(c_ptr), value :: cx, cy, cz        !< c_loc of the convecting velocity, or c_null_ptr
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      function nsb_op_create_stepper(sem, layout, nfields_apply, slot, kappa, rho, dt, nsteps, tol, maxit, op) &
         bind(C, name='nsb_op_create_stepper') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, layout
         integer(c_int), value :: nfields_apply, slot, nsteps, maxit
         real(c_double), value :: kappa, rho, dt, tol
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      function nsb_sem_pressure_setup(sem) bind(C, name='nsb_sem_pressure_setup') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: sem
         integer(c_int) :: ierr
      end function
      function nsb_op_create_ns_stepper(sem, layout, base, col_base, nu, dt, nsteps, tol_v, tol_p, maxit, mean_free, &
                                        precond, op) &
         bind(C, name='nsb_op_create_ns_stepper') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, layout, base       !< base = c_null_ptr: Stokes operator
         integer(c_int), value :: col_base, nsteps, maxit, mean_free, precond
         real(c_double), value :: nu, dt, tol_v, tol_p
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      !> exponential_prop%rmatvec (core/linear_operators.f90:84-103): the same stepper in adjoint mode
      function nsb_op_create_ns_stepper_adjoint(sem, layout, base, col_base, nu, dt, nsteps, tol_v, tol_p, maxit, &
                                                mean_free, precond, op) &
         bind(C, name='nsb_op_create_ns_stepper_adjoint') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, layout, base
         integer(c_int), value :: col_base, nsteps, maxit, mean_free, precond
         real(c_double), value :: nu, dt, tol_v, tol_p
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      !> stored orbit of a time-periodic base flow (core/linear_operators.f90:254-275): step n uses column col0 + (n-1) stride
      function nsb_op_ns_set_orbit(op, orbit, col0, stride) bind(C, name='nsb_op_ns_set_orbit') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: op, orbit
         integer(c_int), value :: col0, stride
         integer(c_int) :: ierr
      end function
      !> norm_grad (core/utils.f90:446-486) of the velocity fields of (b, col), summed over all ranks
      function nsb_sem_norm_grad(sem, b, col, norma) bind(C, name='nsb_sem_norm_grad') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, b
         integer(c_int), value :: col
         real(c_double) :: norma
         integer(c_int) :: ierr
      end function
      !> compute_cfl(cfl, vx, vy, vz, dt) of the velocity fields of (b, col) (core/linear_stab.f90:222,231)
      function nsb_sem_cfl(sem, b, col, dt, cfl) bind(C, name='nsb_sem_cfl') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: sem, b
         integer(c_int), value :: col
         real(c_double), value :: dt
         real(c_double) :: cfl
         integer(c_int) :: ierr
      end function
      function nsb_op_create_compose(layout, outer, inner, op) bind(C, name='nsb_op_create_compose') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: layout, outer, inner
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      !> alpha A + beta B, c_null_ptr = identity: LightKrylov's axpby_linop (core/linear_operators.f90:403),
      !! newton_linearized_map / ts_force_sensitivity_map of core/matvec.f90:499-541
      function nsb_op_create_axpby(layout, A, B, alpha, beta, op) bind(C, name='nsb_op_create_axpby') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: layout, A, B
         real(c_double), value :: alpha, beta
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      !> forward_finite_difference_map (core/matvec.f90:246-379): order = findiff_order (2 or 4)
      function nsb_op_create_frechet_fd(layout, F, base, col_base, order, op) &
         bind(C, name='nsb_op_create_frechet_fd') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: layout, F, base
         integer(c_int), value :: col_base, order
         type(c_ptr) :: op
         integer(c_int) :: ierr
      end function
      function nsb_op_frechet_set_epsilon(op, epsilon_base) bind(C, name='nsb_op_frechet_set_epsilon') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: op
         real(c_double), value :: epsilon_base   !< core/main.f90:16
         integer(c_int) :: ierr
      end function
      function nsb_op_apply(op, bin, cin, bout, cout) bind(C, name='nsb_op_apply') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: op, bin, bout
         integer(c_int), value :: cin, cout
         integer(c_int) :: ierr
      end function
      function nsb_op_destroy(op) bind(C, name='nsb_op_destroy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: op
         integer(c_int) :: ierr
      end function
      function nsb_op_count(op, napply) bind(C, name='nsb_op_count') result(ierr)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: op
         integer(c_int64_t) :: napply
         integer(c_int) :: ierr
      end function
      function nsb_eig(A, lda, n, vecs, vals) bind(C, name='nsb_eig') result(ierr)
         import :: c_int, c_double, c_double_complex
         integer(c_int), value :: lda, n
         real(c_double) :: A(lda, *)
         complex(c_double_complex) :: vecs(n, *), vals(*)
         integer(c_int) :: ierr
      end function
      function nsb_schur(A, lda, n, vecs, vals) bind(C, name='nsb_schur') result(ierr)
         import :: c_int, c_double, c_double_complex
         integer(c_int), value :: lda, n
         real(c_double) :: A(lda, *), vecs(n, *)
         complex(c_double_complex) :: vals(*)
         integer(c_int) :: ierr
      end function
      function nsb_ordschur(T, ldt, Q, ldq, selected, n) bind(C, name='nsb_ordschur') result(ierr)
         import :: c_int, c_double
         integer(c_int), value :: ldt, ldq, n
         real(c_double) :: T(ldt, *), Q(ldq, *)
         integer(c_int) :: selected(*)
         integer(c_int) :: ierr
      end function
      function nsb_lstsq(A, lda, m, n, b, x) bind(C, name='nsb_lstsq') result(ierr)
         import :: c_int, c_double
         integer(c_int), value :: lda, m, n
         real(c_double) :: A(lda, *), b(*), x(*)
         integer(c_int) :: ierr
      end function
      function nsb_svd(A, lda, m, n, U, S, V) bind(C, name='nsb_svd') result(ierr)
         import :: c_int, c_double
         integer(c_int), value :: lda, m, n
         real(c_double) :: A(lda, *), U(m, *), S(*), V(n, *)
         integer(c_int) :: ierr
      end function
      function nsb_set_dgks_eta(ctx, eta) bind(C, name='nsb_set_dgks_eta') result(ierr)
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: ctx
         real(c_double), value :: eta
         integer(c_int) :: ierr
      end function
      function nsb_op_set_linear(op, linear) bind(C, name='nsb_op_set_linear') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: op
         integer(c_int), value :: linear
         integer(c_int) :: ierr
      end function
      function nsb_arnoldi_passes(Q, mstart, mend, orth_mode, passes) bind(C, name='nsb_arnoldi_passes') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: Q
         integer(c_int), value :: mstart, mend, orth_mode
         integer(c_int) :: passes(*)
         integer(c_int) :: ierr
      end function
      function nsb_hessenberg_write(path, H, ldh, k) bind(C, name='nsb_hessenberg_write') result(ierr)
         import :: c_int, c_double, c_char
         character(kind=c_char) :: path(*)
         integer(c_int), value :: ldh, k
         real(c_double) :: H(ldh, *)
         integer(c_int) :: ierr
      end function
      function nsb_hessenberg_read(path, k_dim, mstart, H, ldh) bind(C, name='nsb_hessenberg_read') result(ierr)
         import :: c_int, c_double, c_char
         character(kind=c_char) :: path(*)
         integer(c_int), value :: k_dim, mstart, ldh
         real(c_double) :: H(ldh, *)
         integer(c_int) :: ierr
      end function
      function nsb_restart_load(Q, dir, session, mstart, k_dim, lglel, nel_local, ufield0, pfield, tfield, H, ldh, &
                                mstart_next) bind(C, name='nsb_restart_load') result(ierr)
         import :: c_int, c_int64_t, c_double, c_char, c_ptr
         type(c_ptr), value :: Q
         character(kind=c_char) :: dir(*), session(*)
         integer(c_int), value :: mstart, k_dim, ufield0, pfield, tfield, ldh
         integer(c_int64_t) :: lglel(*)
         integer(c_int64_t), value :: nel_local
         real(c_double) :: H(ldh, *)
         integer(c_int) :: mstart_next
         integer(c_int) :: ierr
      end function
      function nsb_select_eigenvalues(selected, cnt, vals, delta, nev, n) bind(C, name='nsb_select_eigenvalues') &
         result(ierr)
         import :: c_int, c_double, c_double_complex
         integer(c_int) :: selected(*), cnt
         complex(c_double_complex) :: vals(*)
         real(c_double), value :: delta
         integer(c_int), value :: nev, n
         integer(c_int) :: ierr
      end function
   end interface
   public :: nsb_sync, nsb_vec_norm, nsb_orthonormalize, nsb_basis_gram, nsb_basis_qr, nsb_basis_rotate
   public :: nsb_sem_create, nsb_sem_destroy, nsb_sem_setup_exchange, nsb_sem_axhelm, nsb_sem_ax, nsb_sem_dssum
   public :: nsb_sem_col2, nsb_sem_hmholtz, nsb_sem_hmholtz_vec, nsb_sem_dealias_setup, nsb_sem_set_convect, nsb_sem_convect
   public :: nsb_sem_bdf_ext, nsb_op_create_sem, nsb_op_create_stepper, nsb_op_create_compose, nsb_op_create_axpby
   public :: nsb_op_create_frechet_fd, nsb_op_frechet_set_epsilon, nsb_op_ns_set_orbit, nsb_sem_norm_grad
   public :: nsb_sem_cfl
   public :: nsb_sem_pressure_setup, nsb_op_create_ns_stepper, nsb_op_create_ns_stepper_adjoint
   public :: nsb_op_apply, nsb_op_destroy, nsb_op_count
   public :: nsb_eig, nsb_schur, nsb_ordschur, nsb_lstsq, nsb_svd, nsb_select_eigenvalues
   public :: nsb_set_dgks_eta, nsb_op_set_linear, nsb_arnoldi_passes, nsb_hessenberg_write, nsb_hessenberg_read
   public :: nsb_restart_load


contains

   !> The reference prints and calls nek_end on fatal conditions (core/nek_vectors.f90:108-111);
   !! the C ABI returns a code instead, mapped back here.
   subroutine nsb_check(ierr, where)
      integer(c_int), intent(in) :: ierr
      character(len=*), intent(in) :: where
      character(kind=c_char), pointer :: msg(:)
      integer :: n
      if (ierr == 0) return
      call c_f_pointer(nsb_last_error(), msg, [1024])
      n = 1
      do while (n < 1024 .and. msg(n) /= c_null_char)
         n = n + 1
      end do
      write (6, *) 'nekstab_b200 error ', ierr, ' in ', where, ': ', msg(1:n - 1)
      call nek_end
   end subroutine nsb_check

   !> Called once from nekStab_init (core/main.f90:77-136) after bm1s <- bm1 (:108).
   !! nid/np_ are Nek's rank and size; the unique id travels through Nek's own bcast.
   subroutine nsb_startup(nid, np_, local_device, nfields, field_len, field_in_dot, time_in_dot, bm1s, k_dim, nwork)
      integer, intent(in) :: nid, np_, local_device, nfields, time_in_dot, k_dim, nwork
      integer(c_int64_t), intent(in) :: field_len(nfields)
      integer(c_int), intent(in) :: field_in_dot(nfields)
      real(c_double), target, intent(in) :: bm1s(*)
      character(kind=c_char) :: id(128)
      type(c_ptr) :: w(nfields)
      integer :: i, nd
      external :: dgeev, dgees, dtrsen, dgels, dgesvd
      if (nid == 0) call nsb_check(nsb_get_unique_id(id), 'nsb_get_unique_id')
      call bcast(id, 128)                                   ! Nek5000 comm_mpi.f
      call nsb_check(nsb_init(int(local_device, c_int), int(nid, c_int), int(np_, c_int), id, nsb_ctx), 'nsb_init')
      call nsb_check(nsb_layout_create(nsb_ctx, int(nfields, c_int), field_len, field_in_dot, &
                                       int(time_in_dot, c_int), nsb_layout), 'nsb_layout_create')
      nd = 0
      do i = 1, nfields                                     ! the same bm1s for vx, vy, vz, t (k_dot)
         if (field_in_dot(i) /= 0) then
            nd = nd + 1
            w(nd) = c_loc(bm1s)
         end if
      end do
      call nsb_check(nsb_layout_set_weight(nsb_layout, w), 'nsb_layout_set_weight')
      ! k_dim + 1 Krylov vectors + one work column (ts_gmres needs ksize + 2)
      call nsb_check(nsb_basis_create(nsb_layout, int(k_dim + 2, c_int), nsb_Q), 'nsb_basis_create Q')
      call nsb_check(nsb_basis_create(nsb_layout, int(nwork, c_int), nsb_work), 'nsb_basis_create work')
      ! the k x k dense step stays on the host through the LAPACK the case already links
      ! (core/lapack_wrapper.f90:49,108,158,288)
      call nsb_check(nsb_set_lapack(c_funloc(dgeev), c_funloc(dgees), c_funloc(dtrsen), c_funloc(dgels)), &
                     'nsb_set_lapack')
      call nsb_check(nsb_set_lapack_svd(c_funloc(dgesvd)), 'nsb_set_lapack_svd')   ! svds (LightKrylov's svd)
   end subroutine nsb_startup

   subroutine nsb_shutdown()
      integer(c_int) :: ierr
      ierr = nsb_basis_destroy(nsb_Q)
      ierr = nsb_basis_destroy(nsb_work)
      ierr = nsb_finalize(nsb_ctx)
   end subroutine nsb_shutdown

   !> krylov_vector -> device column (fields in layout order) and back.
   subroutine nsb_upload(v, vx, vy, vz, pr, t, time)
      type(nek_dvector), intent(in) :: v
      real(c_double), target, intent(in) :: vx(*), vy(*), vz(*), pr(*), t(*)
      real(c_double), intent(in) :: time
      type(c_ptr) :: f(5)
      f = [c_loc(vx), c_loc(vy), c_loc(vz), c_loc(pr), c_loc(t)]
      call nsb_check(nsb_vec_upload(v%basis, v%col, f, time), 'nsb_vec_upload')
   end subroutine nsb_upload

   subroutine nsb_download(v, vx, vy, vz, pr, t, time)
      type(nek_dvector), intent(in) :: v
      real(c_double), target, intent(inout) :: vx(*), vy(*), vz(*), pr(*), t(*)
      real(c_double), intent(out) :: time
      type(c_ptr) :: f(5)
      f = [c_loc(vx), c_loc(vy), c_loc(vz), c_loc(pr), c_loc(t)]
      call nsb_check(nsb_vec_download(v%basis, v%col, f, time), 'nsb_vec_download')
   end subroutine nsb_download

   ! ---- type-bound procedures (core/nek_vectors.f90:27-30) ------------------------------------
   subroutine dv_zero(self)
      class(nek_dvector), intent(inout) :: self
      call nsb_check(nsb_vec_zero(self%basis, self%col), 'zero')
   end subroutine dv_zero

   real(c_double) function dv_dot(self, vec) result(alpha)
      class(nek_dvector), intent(in) :: self
      class(nek_dvector), intent(in) :: vec
      call nsb_check(nsb_vec_dot(self%basis, self%col, vec%basis, vec%col, alpha), 'dot')
   end function dv_dot

   real(c_double) function dv_norm(self) result(alpha)
      class(nek_dvector), intent(in) :: self
      call nsb_check(nsb_vec_dot(self%basis, self%col, self%basis, self%col, alpha), 'norm')
      alpha = sqrt(alpha)
   end function dv_norm

   subroutine dv_scal(self, alpha)
      class(nek_dvector), intent(inout) :: self
      real(c_double), intent(in) :: alpha
      call nsb_check(nsb_vec_scal(self%basis, self%col, alpha), 'scal')
   end subroutine dv_scal

   !> self <- alpha*self + beta*vec; %time untouched, like real_axpby (core/nek_vectors.f90:127-139)
   subroutine dv_axpby(self, alpha, vec, beta)
      class(nek_dvector), intent(inout) :: self
      class(nek_dvector), intent(in) :: vec
      real(c_double), intent(in) :: alpha, beta
      call nsb_check(nsb_vec_axpby(self%basis, self%col, alpha, vec%basis, vec%col, beta, NSB_AXPBY_SKIP_TIME), 'axpby')
   end subroutine dv_axpby

   ! ---- legacy free functions (core/krylov_subspace.f90:26-209), same argument order ----------
   subroutine k_dot(alpha, p, q)
      real(c_double), intent(out) :: alpha
      type(nek_dvector), intent(in) :: p, q
      call nsb_check(nsb_vec_dot(p%basis, p%col, q%basis, q%col, alpha), 'k_dot')
   end subroutine k_dot

   subroutine k_norm(alpha, p)
      real(c_double), intent(out) :: alpha
      type(nek_dvector), intent(in) :: p
      call k_dot(alpha, p, p)
      alpha = dsqrt(alpha)
   end subroutine k_norm

   subroutine k_normalize(p, alpha)
      type(nek_dvector), intent(inout) :: p
      real(c_double), intent(out) :: alpha
      call nsb_check(nsb_vec_normalize(p%basis, p%col, alpha), 'k_normalize')
   end subroutine k_normalize

   subroutine k_cmult(p, c)
      type(nek_dvector), intent(inout) :: p
      real(c_double), intent(in) :: c
      call nsb_check(nsb_vec_scal(p%basis, p%col, c), 'k_cmult')
   end subroutine k_cmult

   subroutine k_add2(p, q)
      type(nek_dvector), intent(inout) :: p
      type(nek_dvector), intent(in) :: q
      call nsb_check(nsb_vec_add2(p%basis, p%col, q%basis, q%col), 'k_add2')
   end subroutine k_add2

   subroutine k_sub2(p, q)
      type(nek_dvector), intent(inout) :: p
      type(nek_dvector), intent(in) :: q
      call nsb_check(nsb_vec_sub2(p%basis, p%col, q%basis, q%col), 'k_sub2')
   end subroutine k_sub2

   subroutine k_sub3(p, q, r)
      type(nek_dvector), intent(inout) :: p
      type(nek_dvector), intent(in) :: q, r
      call nsb_check(nsb_vec_sub3(p%basis, p%col, q%basis, q%col, r%basis, r%col), 'k_sub3')
   end subroutine k_sub3

   subroutine k_zero(p)
      type(nek_dvector), intent(inout) :: p
      call nsb_check(nsb_vec_zero(p%basis, p%col), 'k_zero')
   end subroutine k_zero

   subroutine k_copy(p, q)        ! destination first, like the reference
      type(nek_dvector), intent(inout) :: p
      type(nek_dvector), intent(in) :: q
      call nsb_check(nsb_vec_copy(p%basis, p%col, q%basis, q%col), 'k_copy')
   end subroutine k_copy

   !> dq = Q(1:k) * yvec (core/krylov_subspace.f90:163-209); Q is the device basis.
   subroutine k_matmul_d(dq, yvec, k)
      type(nek_dvector), intent(inout) :: dq
      integer, intent(in) :: k
      real(c_double), intent(in) :: yvec(k)
      call nsb_check(nsb_basis_gemv(nsb_Q, int(k, c_int), yvec, dq%basis, dq%col), 'k_matmul')
   end subroutine k_matmul_d

   ! ---- the reference's own signatures --------------------------------------------------------
   !> Where the reference calls prepare_linearized_solver / selects the operator behind `matvec`
   !! (core/matvec.f90:56-146): every solver below applies this handle (nsb_op_create_host wraps the host
   !! time-stepper, nsb_op_create_sem / _stepper / _compose are device operators).
   subroutine nsb_set_operator(op)
      type(c_ptr), intent(in) :: op
      nsb_current_op = op
   end subroutine nsb_set_operator

   !> Q(i) <- column i-1 of `basis`: the replacement of allocate(Q(k_dim+1)) (core/eigensolvers.f90:149) --
   !! the array of vectors the reference passes around becomes a view of the device-resident basis.
   subroutine nsb_bind_basis(Q, basis, ncols)
      integer, intent(in) :: ncols
      type(nek_dvector), intent(out) :: Q(ncols)
      type(c_ptr), intent(in) :: basis
      integer :: i
      do i = 1, ncols
         Q(i)%basis = basis
         Q(i)%col = int(i - 1, c_int)
      end do
   end subroutine nsb_bind_basis

   !> Q must be consecutive columns of one device basis, Q(1) its column 0 (nsb_bind_basis).
   subroutine nsb_require_view(Q, n, where)
      integer, intent(in) :: n
      type(nek_dvector), intent(in) :: Q(n)
      character(len=*), intent(in) :: where
      integer :: i
      do i = 1, n
         if (.not. c_associated(Q(i)%basis, Q(1)%basis) .or. Q(i)%col /= i - 1) then
            write (6, *) where, ': Q is not a view of one device basis (use nsb_bind_basis)'
            call nek_end
         end if
      end do
   end subroutine nsb_require_view

   !> arnoldi_factorization(Q, H, mstart, mend, ksize) -- core/krylov_decomposition.f90:2-99, same arguments.
   subroutine arnoldi_factorization(Q, H, mstart, mend, ksize)
      integer, intent(in) :: mstart, mend, ksize
      type(nek_dvector), dimension(ksize + 1) :: Q
      real(c_double), dimension(ksize + 1, ksize) :: H
      if (ksize == 0) then                                  ! core/krylov_decomposition.f90:59-62
         write (6, *) 'Krylov base dimension == 0! Increase it.'
         call nek_end
      end if
      call nsb_require_view(Q, ksize + 1, 'arnoldi_factorization')
      call nsb_check(nsb_arnoldi(Q(1)%basis, nsb_current_op, int(mstart - 1, c_int), int(mend - 1, c_int), &
                                 nsb_orth_mode, H, int(ksize + 1, c_int)), 'arnoldi_factorization')
   end subroutine arnoldi_factorization

   !> update_hessenberg_matrix(H, f, q, k) -- core/krylov_decomposition.f90:103-189, same arguments; f must be
   !! a column of q's basis at or behind column k (it is orthonormalised in place against q(1:k)).
   subroutine update_hessenberg_matrix(H, f, q, k)
      integer, intent(in) :: k
      real(c_double), dimension(k + 1, k) :: H
      type(nek_dvector) :: f
      type(nek_dvector), dimension(k) :: q
      real(c_double) :: h_col(k + 1)
      integer(c_int) :: passes
      call nsb_require_view(q, k, 'update_hessenberg_matrix')
      if (.not. c_associated(f%basis, q(1)%basis)) then
         write (6, *) 'update_hessenberg_matrix: f must live in the basis of q'
         call nek_end
      end if
      call nsb_check(nsb_orthonormalize(f%basis, int(k, c_int), f%col, nsb_orth_mode, h_col, passes), &
                     'update_hessenberg_matrix')
      H(1:k + 1, k) = h_col
   end subroutine update_hessenberg_matrix

   !> k_matmul(dq, Q, yvec, k) -- core/krylov_subspace.f90:163-209, same arguments.
   subroutine k_matmul_q(dq, Q, yvec, k)
      integer, intent(in) :: k
      type(nek_dvector) :: dq
      type(nek_dvector), dimension(k) :: Q
      real(c_double), dimension(k) :: yvec
      call nsb_require_view(Q, k, 'k_matmul')
      call nsb_check(nsb_basis_gemv(Q(1)%basis, int(k, c_int), yvec, dq%basis, dq%col), 'k_matmul')
   end subroutine k_matmul_q

   !> schur_condensation(mstart, H, Q, ksize) -- core/eigensolvers.f90:363-468, same arguments; schur_del and
   !! schur_tgt come from the module variables (the reference reads them from the NEKSTAB commons).
   subroutine schur_condensation(mstart, H, Q, ksize)
      integer, intent(inout) :: mstart
      integer, intent(in) :: ksize
      real(c_double), dimension(ksize + 1, ksize), intent(inout) :: H
      type(nek_dvector), dimension(ksize + 1) :: Q
      integer(c_int) :: m
      call nsb_require_view(Q, ksize + 1, 'schur_condensation')
      m = int(mstart - 1, c_int)
      call nsb_check(nsb_schur_condensation(Q(1)%basis, m, H, int(ksize + 1, c_int), int(ksize, c_int), &
                                            nsb_schur_del, int(nsb_schur_tgt, c_int)), 'schur_condensation')
      mstart = m + 1
   end subroutine schur_condensation

   !> ts_gmres(rhs, sol, maxiter, ksize, calls) -- core/newton_krylov.f90:170-299, same arguments.  The reference
   !! allocates its Krylov basis inside; here the module basis nsb_Q (>= ksize + 2 columns) is used.
   subroutine ts_gmres(rhs, sol, maxiter, ksize, calls)
      type(nek_dvector), intent(in) :: rhs
      type(nek_dvector), intent(inout) :: sol
      integer, intent(in) :: maxiter, ksize
      integer, intent(out) :: calls
      call ts_gmres_d(nsb_current_op, rhs, sol, maxiter, ksize, nsb_gmres_tol, calls)
   end subroutine ts_gmres

   ! ---- solver entry points -------------------------------------------------------------------
   !> arnoldi_factorization(Q, H, mstart, mend, ksize) (core/krylov_decomposition.f90:2-99):
   !! 1-based mstart/mend as in the reference; `op` wraps the host matvec (nsb_op_create_host).
   subroutine arnoldi_factorization_d(op, H, mstart, mend, ksize)
      type(c_ptr), intent(in) :: op
      integer, intent(in) :: mstart, mend, ksize
      real(c_double), intent(inout) :: H(ksize + 1, ksize)
      call nsb_check(nsb_arnoldi(nsb_Q, op, int(mstart - 1, c_int), int(mend - 1, c_int), NSB_ORTH_CGS2, &
                                 H, int(ksize + 1, c_int)), 'arnoldi_factorization')
   end subroutine arnoldi_factorization_d

   !> schur_condensation(mstart, H, Q, ksize) (core/eigensolvers.f90:363-468)
   subroutine schur_condensation_d(mstart, H, ksize, schur_del, schur_tgt)
      integer, intent(inout) :: mstart
      integer, intent(in) :: ksize, schur_tgt
      real(c_double), intent(inout) :: H(ksize + 1, ksize)
      real(c_double), intent(in) :: schur_del
      integer(c_int) :: m
      m = int(mstart - 1, c_int)
      call nsb_check(nsb_schur_condensation(nsb_Q, m, H, int(ksize + 1, c_int), int(ksize, c_int), schur_del, &
                                            int(schur_tgt, c_int)), 'schur_condensation')
      mstart = m + 1
   end subroutine schur_condensation_d

   !> The restart loop of krylov_schur (core/eigensolvers.f90:295-333); Q(1) must hold the seed.
   subroutine krylov_schur_d(op, k_dim, schur_tgt, eigen_tol, schur_del, H, vals, vecs, residual, cnt, schur_cnt)
      type(c_ptr), intent(in) :: op
      integer, intent(in) :: k_dim, schur_tgt
      real(c_double), intent(in) :: eigen_tol, schur_del
      real(c_double), intent(inout) :: H(k_dim + 1, k_dim), residual(k_dim)
      complex(c_double_complex), intent(out) :: vals(k_dim), vecs(k_dim, k_dim)
      integer, intent(out) :: cnt, schur_cnt
      integer(c_int) :: c1, c2
      call nsb_check(nsb_krylov_schur(nsb_Q, op, int(k_dim, c_int), int(schur_tgt, c_int), eigen_tol, schur_del, &
                                      NSB_ORTH_CGS2, 200_c_int, H, int(k_dim + 1, c_int), vals, vecs, residual, &
                                      c1, c2), 'krylov_schur')
      cnt = c1
      schur_cnt = c2
   end subroutine krylov_schur_d

   !> eigs(A, X, eigvecs, eigvals, residuals, info, nev, tolerance) (call site core/linear_stab.f90:66)
   subroutine eigs_d(op, k_dim, eigvecs, eigvals, residuals, info, nev, tolerance)
      type(c_ptr), intent(in) :: op
      integer, intent(in) :: k_dim
      complex(c_double_complex), intent(out) :: eigvecs(k_dim, k_dim), eigvals(k_dim)
      real(c_double), intent(out) :: residuals(k_dim)
      integer, intent(out) :: info
      integer, intent(in) :: nev
      real(c_double), intent(in) :: tolerance
      real(c_double), allocatable :: H(:, :)
      integer(c_int) :: ku, nc
      allocate (H(k_dim + 1, k_dim))
      call nsb_check(nsb_eigs(nsb_Q, op, int(k_dim, c_int), int(nev, c_int), tolerance, NSB_ORTH_CGS2, H, &
                              int(k_dim + 1, c_int), eigvals, eigvecs, residuals, ku, nc), 'eigs')
      info = ku
      deallocate (H)
   end subroutine eigs_d

   !> svds(A, U, V, uvecs, vvecs, sigma, residuals, info, nev, tolerance) (call site core/linear_stab.f90:112);
   !> U, V are device bases (k_dim+1 and k_dim columns), op_adj applies A%rmatvec
   subroutine svds_d(op, op_adj, U, V, k_dim, uvecs, vvecs, sigma, residuals, info, nev, tolerance)
      type(c_ptr), intent(in) :: op, op_adj, U, V
      integer, intent(in) :: k_dim
      real(c_double), intent(out) :: uvecs(k_dim, k_dim), vvecs(k_dim, k_dim), sigma(k_dim), residuals(k_dim)
      integer, intent(out) :: info
      integer, intent(in) :: nev
      real(c_double), intent(in) :: tolerance
      real(c_double), allocatable :: B(:, :)
      integer(c_int) :: ku, nc
      allocate (B(k_dim + 1, k_dim))
      call nsb_check(nsb_svds(U, V, op, op_adj, int(k_dim, c_int), int(nev, c_int), tolerance, NSB_ORTH_CGS2, B, &
                              int(k_dim + 1, c_int), sigma, uvecs, vvecs, residuals, ku, nc), 'svds')
      info = ku
      deallocate (B)
   end subroutine svds_d

   !> ts_gmres(rhs, sol, maxiter, ksize, calls) (core/newton_krylov.f90:170-299)
   subroutine ts_gmres_d(op, rhs, sol, maxiter, ksize, tol, calls)
      type(c_ptr), intent(in) :: op
      type(nek_dvector), intent(in) :: rhs
      type(nek_dvector), intent(inout) :: sol
      integer, intent(in) :: maxiter, ksize
      real(c_double), intent(in) :: tol
      integer, intent(out) :: calls
      real(c_double) :: hist(maxiter)
      integer(c_int) :: c, nh
      call nsb_check(nsb_ts_gmres(nsb_Q, op, rhs%basis, rhs%col, sol%basis, sol%col, int(maxiter, c_int), &
                                  int(ksize, c_int), tol, NSB_ORTH_CGS2, c, hist, nh), 'ts_gmres')
      calls = c
   end subroutine ts_gmres_d

end module nekstab_b200
