/*
 * nekstab_b200.h -- C ABI of the B200-native nekStab Arnoldi / Newton-Krylov hot path.
 *
 * This is the boundary the reference's Fortran host code binds through ISO_C_BINDING
 * (see nekstab_next_b200/fortran/nekstab_b200.f90 and INTEGRATION.md).  Every entry point names the reference
 * routine it replaces (paths relative to the nekStab repository root).
 *
 * Conventions
 *  - plain C types only; all handles are opaque pointers owned by the library;
 *  - every function returns 0 on success and a negative NSB_E* code otherwise, never aborts
 *    (the reference prints and calls nek_end -- the Fortran shim maps nonzero to that);
 *    nsb_last_error() gives the message of the last failure on the calling thread;
 *  - pointers are HOST pointers unless the parameter name ends in _d;
 *  - indices (columns, k, mstart, mend) are 0-based on this side; the Fortran shim converts;
 *  - matrices (H, Z, y) are column-major doubles with an explicit leading dimension,
 *    exactly as Fortran passes them;
 *  - one CUDA stream per context; calls that return host scalars synchronise that stream,
 *    the others only enqueue work.
 *  - there is NO CPU fallback: without a CUDA device nsb_init fails with NSB_ENODEVICE.
 */
#ifndef NEKSTAB_B200_H
#define NEKSTAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSB_VERSION 100

/* error codes */
#define NSB_OK 0
#define NSB_EINVAL (-1)     /* bad argument */
#define NSB_ECUDA (-2)      /* CUDA runtime error */
#define NSB_ENODEVICE (-3)  /* no usable CUDA device (there is no CPU fallback) */
#define NSB_ENAN (-4)       /* NaN in an inner product (reference: nek_end) */
#define NSB_ENCCL (-5)      /* NCCL error / NCCL not loadable */
#define NSB_ELAPACK (-6)    /* LAPACK provider missing or LAPACK info != 0 */
#define NSB_EBREAKDOWN (-7) /* Krylov breakdown (zero residual norm) */

typedef struct nsb_context_s *nsb_context_t;
typedef struct nsb_layout_s *nsb_layout_t;
typedef struct nsb_basis_s *nsb_basis_t;
typedef struct nsb_sem_s *nsb_sem_t;
typedef struct nsb_op_s *nsb_op_t;

const char *nsb_last_error(void);
int nsb_version(void);
/* Hash of the sources this binary was compiled from (sha256 over csrc/ and this header, first 16 hex digits),
 * so that a measured number can be tied to a source tree: bench.py prints it, build.py compares it. */
const char *nsb_build_id(void);

/* ---------------------------------------------------------------------------------------------
 * Context: one per process / MPI rank / GPU.  (Reference: the Nek rank, nid/np in SIZE/PARALLEL;
 * collectives replace gop -> MPI_Allreduce inside glsc3, core/nek_vectors.f90:7-12.)
 * ------------------------------------------------------------------------------------------- */
#define NSB_UNIQUE_ID_BYTES 128
/* rank 0 calls this and broadcasts the bytes with the host's own transport (MPI_Bcast in Nek). */
int nsb_get_unique_id(void *id_out /* NSB_UNIQUE_ID_BYTES */);
/* unique_id may be NULL when nranks == 1. */
int nsb_init(int device, int rank, int nranks, const void *unique_id, nsb_context_t *ctx);
int nsb_finalize(nsb_context_t ctx);
/* NVLink peer-memory mailbox (optional, single node, <= 16 ranks).  Every rank creates a mailbox and
 * gets a 64-byte CUDA IPC handle; the host all-gathers the handles with its own transport
 * (MPI_Allgather in Nek) and every rank maps all of them.  Afterwards the all-reduces of the inner
 * products (gop -> MPI_Allreduce in the reference) and the dssum interface exchange are kernels
 * that store straight into the peers' memory -- no NCCL launch on the hot path.  halo_bytes sizes
 * the interface area (2 slots x 8 fields x shared nodes x 8 B per neighbour; falls back to NCCL if
 * too small).  Call before nsb_sem_setup_exchange. */
#define NSB_IPC_HANDLE_BYTES 64
int nsb_p2p_mailbox_create(nsb_context_t ctx, int64_t halo_bytes, void *handle_out);
int nsb_p2p_mailbox_connect(nsb_context_t ctx, const void *all_handles /* nranks x 64 B */);
int nsb_p2p_enabled(nsb_context_t ctx, int *enabled);
int nsb_sync(nsb_context_t ctx);
int nsb_rank(nsb_context_t ctx, int *rank, int *nranks);
/* CUDA stream of the context as an integer handle (for event timing by the caller). */
int nsb_stream(nsb_context_t ctx, uint64_t *stream);
/* Kernels launched by this context since creation (for the bench's gpu_launches). */
int nsb_launch_count(nsb_context_t ctx, int64_t *count);
/* Device-side event timing on the context stream. */
int nsb_timer_start(nsb_context_t ctx);
int nsb_timer_stop(nsb_context_t ctx, double *elapsed_ms); /* synchronises */
/* Per-kernel-class device timing for the roofline report: when enabled, every launch is bracketed
 * by CUDA events on the context stream.  Classes: 0 multidot (h = V^T W w), 1 update (w -= V h),
 * 2 normalize, 3 axhelm, 4 gather-scatter (dssum), 5 BLAS-1, 6 small reductions, 7 rotate,
 * 8 gemv, 9 single dot, 10 fused update+multidot (w -= V h1 ; h2 = V^T W w).  bytes = algorithmic bytes summed over the recorded launches. */
int nsb_prof_enable(nsb_context_t ctx, int on); /* also clears the records */
int nsb_prof_get(nsb_context_t ctx, int cls, double *ms, int64_t *launches, double *bytes);
/* cudaProfilerStart / Stop, so `ncu --profile-from-start off` captures only a delimited region. */
int nsb_profiler_start(void);
int nsb_profiler_stop(void);
/* Sum-allreduce n doubles held on the host across ranks (gop(x,'+')); no-op for one rank. */
int nsb_allreduce_host(nsb_context_t ctx, double *x, int n);
/* Write a buffer larger than L2 (bench hygiene). */
int nsb_flush_l2(nsb_context_t ctx);

/* ---------------------------------------------------------------------------------------------
 * Layout of one state vector = the fields of krylov_vector / real_nek_vector
 * (core/krylov_subspace.f90:12-17, core/nek_vectors.f90:20-31): vx, vy, [vz], [pr], [t(:,m)], time.
 *   field_len[i]    active length of field i on this rank (nv, n2, nt ...)
 *   field_in_dot[i] 1 if the field enters the inner product (velocity, temperature, scalars),
 *                   0 otherwise (pressure never does: core/krylov_subspace.f90:40-49)
 *   time_in_dot     1: dot adds time*time (new API always, core/nek_vectors.f90:105-107;
 *                   legacy only when uparam(1)==2.1, core/krylov_subspace.f90:52-54)
 * ------------------------------------------------------------------------------------------- */
int nsb_layout_create(nsb_context_t ctx, int nfields, const int64_t *field_len,
                      const int *field_in_dot, int time_in_dot, nsb_layout_t *layout);
int nsb_layout_destroy(nsb_layout_t layout);
/* rows of one column (padded leading dimension), rows covered by the inner product */
int nsb_layout_info(nsb_layout_t layout, int64_t *ld, int64_t *ndot, int64_t *ndof_dot);
/* Inner-product weight bm1s (core/NEKSTAB:86-89; bm1s <- bm1 in core/main.f90:108; zeroed inside
 * sponges, core/forcing.f90:102-104).  One weight array per in-dot field, field order; pass the
 * same pointer several times to reuse bm1s for vx, vy, vz like the reference does. */
int nsb_layout_set_weight(nsb_layout_t layout, const double *const *weight_per_dot_field);

/* C0 layout (optional): the first n_c0 fields are CONTINUOUS fields on `sem`'s mesh and are stored once per
 * distinct node instead of once per element-local point (Nek duplicates shared nodes: 16.78 M local points for
 * 11.39 M nodes on the 32^3, N = 7 box).  For continuous fields the reference's inner product
 * sum_local a bm1s b (core/krylov_subspace.f90:40-49) equals sum_nodes a (QQ^T bm1s) b, which is what the
 * kernels then compute -- every sweep over the basis moves 32 % fewer bytes, results are the same to rounding.
 * The host side is unchanged: field_len[] are the element-local lengths, nsb_layout_set_weight takes the
 * element-local bm1s, nsb_vec_upload / nsb_vec_download exchange element-local arrays (upload keeps the first
 * copy of a shared node -- the field must be continuous; use nsb_layout_create for anything else).  Device
 * operators: nsb_op_create_sem works on the layout's mesh; the time-stepper pieces need the element-local layout.
 * Call after nsb_sem_setup_exchange. */
int nsb_layout_create_c0(nsb_context_t ctx, nsb_sem_t sem, int nfields, const int64_t *field_len,
                         const int *field_in_dot, int time_in_dot, int n_c0, nsb_layout_t *layout);
/* n_c0 (0 for an element-local layout) and the rows stored per C0 field */
int nsb_layout_is_c0(nsb_layout_t layout, int *n_c0, int64_t *stored_rows_per_field);

/* ---------------------------------------------------------------------------------------------
 * Basis: device-resident column-major tall-skinny fp64 array V[ld, ncols]; a (basis, col) pair
 * is one nek_dvector.  Replaces allocate(Q(k_dim+1)) (core/eigensolvers.f90:149,
 * core/linear_stab.f90:60) and stand-alone work vectors (f, wrk, sol, dq ...).
 * ------------------------------------------------------------------------------------------- */
int nsb_basis_create(nsb_layout_t layout, int ncols, nsb_basis_t *basis);
int nsb_basis_destroy(nsb_basis_t basis);
int nsb_basis_ncols(nsb_basis_t basis, int *ncols);
/* device address of a column (for callers that run their own kernels on it) */
int nsb_basis_col_ptr(nsb_basis_t basis, int col, uint64_t *ptr_d);

/* host <-> device; fields[i] has field_len[i] doubles (NULL: field zero-filled / skipped) */
int nsb_vec_upload(nsb_basis_t b, int col, const double *const *fields, double time);
int nsb_vec_download(nsb_basis_t b, int col, double *const *fields, double *time);

/* BLAS-1 set, one kernel each over the whole column.
 *   zero   : real_zero / k_zero            (core/nek_vectors.f90:70-78, krylov_subspace.f90:141-150)
 *   copy   : k_copy(dst, src), dest first  (core/krylov_subspace.f90:152-161)
 *   scal   : real_scal / k_cmult           (core/nek_vectors.f90:116-125, krylov_subspace.f90:94-104)
 *   axpby  : self <- alpha*self + beta*vec (core/nek_vectors.f90:127-139, 250-256)
 *   add2/sub2/sub3 : k_add2, k_sub2, k_sub3 (core/krylov_subspace.f90:106-139)
 * NSB_AXPBY_SKIP_TIME reproduces real_axpby's quirk of leaving %time untouched. */
#define NSB_AXPBY_SKIP_TIME 1
int nsb_vec_zero(nsb_basis_t b, int col);
int nsb_vec_copy(nsb_basis_t bdst, int cdst, nsb_basis_t bsrc, int csrc);
int nsb_vec_scal(nsb_basis_t b, int col, double alpha);
int nsb_vec_axpby(nsb_basis_t bself, int cself, double alpha, nsb_basis_t bvec, int cvec,
                  double beta, int flags);
int nsb_vec_add2(nsb_basis_t bp, int cp, nsb_basis_t bq, int cq);
int nsb_vec_sub2(nsb_basis_t bp, int cp, nsb_basis_t bq, int cq);
int nsb_vec_sub3(nsb_basis_t bp, int cp, nsb_basis_t bq, int cq, nsb_basis_t br, int cr);

/* BM1-weighted inner product incl. the allreduce over ranks:
 *   real_dot / k_dot / inner_product (core/nek_vectors.f90:80-114, krylov_subspace.f90:26-60,
 *   eigensolvers.f90:3-56); NaN -> NSB_ENAN.   norm: k_norm (:62-73); normalize: k_normalize (:75-92). */
int nsb_vec_dot(nsb_basis_t ba, int ca, nsb_basis_t bb, int cb, double *alpha);
int nsb_vec_norm(nsb_basis_t b, int col, double *alpha);
int nsb_vec_normalize(nsb_basis_t b, int col, double *alpha);

/* ---------------------------------------------------------------------------------------------
 * Orthogonalisation of column col_w against columns 0..k-1 of the same basis, then
 * normalisation: update_hessenberg_matrix (core/krylov_decomposition.f90:103-189).
 *   h[0..k-1] = H(1:k,k) (sum over passes), h[k] = H(k+1,k) = ||w|| after orthogonalisation.
 * Modes:
 *   NSB_ORTH_MGS2_REF  literal reference: column-by-column MGS, unconditional second pass
 *                      (2k dots + 2k updates; slow, for parity tests)
 *   NSB_ORTH_CGS2      fused multi-column: h1 = V^T W w ; {w -= V h1 ; h2 = V^T W w} ; w -= V h2
 *                      (same two-pass semantics, H = h1 + h2; V crosses HBM three times and there
 *                      are 3 all-reduces per step)
 *   NSB_ORTH_DGKS      as CGS2, second pass only if ||w'|| < eta ||w|| (eta = 1/sqrt 2).  The test is a
 *                      device-side predicate: the last CTA of the fused sweep compares the two norms and
 *                      the third sweep exits at once when the flag says so -- no host round trip, V
 *                      crosses HBM twice instead of three times in the common case
 * ------------------------------------------------------------------------------------------- */
#define NSB_ORTH_MGS2_REF 0
#define NSB_ORTH_CGS2 1
#define NSB_ORTH_DGKS 2
/* Threshold of the DGKS test, 0 < eta < 1 (default 1/sqrt 2, Daniel-Gragg-Kaufman-Stewart): the second
 * projection is taken when ||w'|| < eta ||w||.  One classical Gram-Schmidt pass against a basis that is
 * orthonormal to rounding leaves an orthogonality error of O(eps / eta), so smaller values (0.1) still meet
 * the 1e-10 bound while operators of the form I - tau L -- whose Krylov vectors always lose more than
 * 1 - 1/sqrt 2 of their norm in the projection -- take one pass instead of two. */
int nsb_set_dgks_eta(nsb_context_t ctx, double eta);
int nsb_orthonormalize(nsb_basis_t b, int k, int col_w, int mode, double *h, int *passes);
/* Asynchronous variant: h stays on the device until nsb_sync / the next synchronising call;
 * h_pinned must be memory from nsb_host_alloc with room for k + 2 doubles (h[0..k], and for NSB_ORTH_DGKS
 * the number of passes taken in h_pinned[k+1]).  Used by the device-resident Arnoldi loop. */
int nsb_orthonormalize_async(nsb_basis_t b, int k, int col_w, int mode, double *h_pinned);
int nsb_host_alloc(void **ptr, int64_t bytes);
int nsb_host_free(void *ptr);
/* Gram matrix G = V(:,0:k)^T W V(:,0:k) (k x k, ldg) -- the orthonormality.dat check of
 * core/eigensolvers.f90:335-345 in one pass. */
int nsb_basis_gram(nsb_basis_t b, int k, double *G, int ldg);

/* BM1-weighted QR of the first k columns, in place (X = Q R): qr_dec of BoostConv
 * (core/fixedp.f90:331-385) on the same weighted Gram-Schmidt kernels; R is k x k upper triangular
 * (ldr).  Columns with residual norm^2 < 1e-60 are zeroed with R(j,j) = 1, like the reference. */
int nsb_basis_qr(nsb_basis_t b, int k, int orth_mode, double *R, int ldr);
/* dq = sum_i y_i Q_i: k_matmul (core/krylov_subspace.f90:163-209), Ritz vectors
 * (core/eigensolvers.f90:565-574, one call for Re and one for Im coefficients). */
int nsb_basis_gemv(nsb_basis_t b, int k, const double *y, nsb_basis_t bout, int col_out);
/* Q(:,0:k) <- Q(:,0:k) * Z, in place: schur_condensation (core/eigensolvers.f90:421-442).
 * With rotate_time == 0 the %time component is left alone, as the reference does. */
int nsb_basis_rotate(nsb_basis_t b, int k, const double *Z, int ldz, int rotate_time);

/* ---------------------------------------------------------------------------------------------
 * Spectral-element operator pieces ([UPSTREAM-RECALL] Nek5000, reached by the reference through
 * nek_advance, core/linear_operators.f90:247).
 * ------------------------------------------------------------------------------------------- */
/* GLL nodes, weights, derivative matrix D[i + (N+1)*j] = dl_j/dx(z_i) (Fortran dxm1(i,j)). */
int nsb_gll(int N, double *z, double *w, double *D);

/* A mesh partition resident on the device.
 *   dim      2 or 3;  N polynomial order (lx1 = N+1);  nel local elements
 *   x,y,z    nodal coordinates, element-local Nek layout (i fastest), nel*lx1^dim each (z NULL in 2-D)
 *   mask     Dirichlet mask (v1mask), 0/1 doubles, or NULL for all ones
 *   glo_num  global node ids (Nek glo_num, any non-negative int64), identical ids <=> same node
 * Geometry (jac, bm1, G1..G6, rx..tz) is computed on the device (coef.f glmapm1/geodat1). */
int nsb_sem_create(nsb_context_t ctx, int dim, int N, int64_t nel, const double *x, const double *y,
                   const double *z, const double *mask, const int64_t *glo_num, nsb_sem_t *sem);
int nsb_sem_destroy(nsb_sem_t sem);
/* which = 0 bm1, 1 jac, 2 binvm1 (1/dssum(bm1)), 3 vmult (1/multiplicity), 4 mask,
 *         10..15 G1..G6 (2-D: 10,11,13 = G1,G2,G4) ; out has nel*lx1^dim doubles */
int nsb_sem_get(nsb_sem_t sem, int which, double *out);
int64_t nsb_sem_npts(nsb_sem_t sem);
/* Multi-rank: neighbours are found from glo_num; call once after every rank created its sem. */
int nsb_sem_setup_exchange(nsb_sem_t sem);

/* Host-only planning steps of the gather-scatter (no CUDA, no NCCL): the same code the device
 * path runs inside nsb_sem_create / nsb_sem_setup_exchange, exposed so the multi-rank logic can be
 * tested on CPU-only machines.
 *   gs_plan: unique nodes owning >= 1 element-boundary point as CSR lists (call with off/idx/gid
 *            NULL to get nnodes/nnz first);
 *   exchange_plan: given every rank's sorted node ids, which of my nodes are interface nodes,
 *            their position after the private-first reorder, and per peer the interface-relative
 *            node indices in ascending global id (the order of the packed exchange buffers). */
int nsb_host_gs_plan(int dim, int N, int64_t nel, const int64_t *glo_num, int64_t *nnodes,
                     int64_t *nnz, int64_t *off, int32_t *idx, int64_t *gid);
int nsb_host_exchange_plan(int rank, int nranks, int64_t nnodes, const int64_t *gid,
                           const int64_t *counts, const int64_t *all_sorted, int64_t mx,
                           int64_t *newpos, int64_t *n_local, int64_t *peer_count,
                           int32_t *peer_nodes);

/* Kernels on field `field` of column `col` (element-local array of nel*lx1^dim points):
 *   axhelm : w = h1 * D^T G D u + h2 * bm1 * u                (hmholtz.f axhelm, no dssum)
 *   dssum  : u <- QQ^T u  incl. the inter-rank exchange        (dssum / gs_op add)
 *   col2   : u <- u * c with c = mask, binvm1, vmult or bm1    (col2)
 *   ax     : axhelm + dssum + mask                             (Nek ax(w,x,h1,h2,n)) */
int nsb_sem_axhelm(nsb_sem_t sem, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout, int field,
                   double h1, double h2);
int nsb_sem_dssum(nsb_sem_t sem, nsb_basis_t b, int col, int field);
int nsb_sem_col2(nsb_sem_t sem, nsb_basis_t b, int col, int field, int which);
int nsb_sem_ax(nsb_sem_t sem, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout, int field,
               double h1, double h2);

/* Jacobi-preconditioned conjugate gradients for (h1 A + h2 B) x = rhs on one field: Nek5000's
 * hmholtz/cggo with setprec ([UPSTREAM-RECALL]; the solve nek_advance runs per velocity component,
 * SURVEY.md section 8 f-3), i.e. the loop {axhelm, dssum, mask, glsc3 with vmult}.  Stops when the
 * preconditioned residual norm sqrt((r, D r)_mult) has dropped by `tol` or after maxit iterations. */
int nsb_sem_hmholtz(nsb_sem_t sem, nsb_basis_t brhs, int crhs, nsb_basis_t bx, int cx, int field,
                    double h1, double h2, double tol, int maxit, int *iters, double *res);
/* The same solve for nf <= 3 fields field0..field0+nf-1 side by side (Nek's ophinv: the three velocity
 * components): one axhelm and one gather-scatter launch per iteration read the geometric factors once for
 * all systems; every system keeps its own CG scalars and convergence test (iters[nf], res[nf]) and follows
 * exactly the iteration sequence it would follow alone. */
int nsb_sem_hmholtz_vec(nsb_sem_t sem, nsb_basis_t brhs, int crhs, nsb_basis_t bx, int cx, int field0, int nf,
                        double h1, double h2, double tol, int maxit, int *iters, double *res);

/* Dealiased convection of Nek5000's perturbation step ([UPSTREAM-RECALL] convect.f set_dealias_rx /
 * set_convect_new / convect_new: the terms advabp adds for (U.grad)u' and (u'.grad)U; SURVEY.md
 * section 8 f-3).  The fine mesh has lxd Gauss-Legendre points per direction (lxd <= 0: Nek's 3 lx1 / 2).
 * 3-D: line kernels for (lx1, lxd) = (8,12), (6,9), (5,8), (4,6); 2-D (the reference's cylinder and
 * backward-facing-step meshes): any lxd, `dim` velocity fields instead of three.
 *   dealias_setup : metrics rxm1..tzm1 interpolated to the fine mesh times the Gauss weights
 *   set_convect   : slot (0 or 1) <- contravariant fine-mesh form of the velocity in fields
 *                   field0..field0+2 of (b, col)
 *   convect       : out(field f) (+)= scale * J^T [ (c_slot . grad_rst)(J in(field f)) ],
 *                   f = field0..field0+nf-1; the mass matrix and the Jacobian are inside c */
int nsb_sem_dealias_setup(nsb_sem_t sem, int lxd);
/* Host only: lxd Gauss-Legendre nodes / weights, the GLL(N) -> GL(lxd) interpolation matrix J[lxd][N+1] and the
 * derivative matrix on the Gauss nodes Dg[lxd][lxd] (both row-major) that the kernels above use; any output
 * may be NULL. */
int nsb_dealias_matrices(int N, int lxd, double *zd, double *wd, double *J, double *Dg);
int nsb_sem_set_convect(nsb_sem_t sem, int slot, nsb_basis_t b, int col, int field0);
int nsb_sem_convect(nsb_sem_t sem, int slot, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout,
                    int field0, int nf, double scale, int accumulate);
/* The exact transpose of nsb_sem_convect as a matrix on the local points: sum_p v_p (C u)_p = sum_p u_p (C^T v)_p;
 * the convective term of the discrete adjoint time-stepper below. */
int nsb_sem_convect_t(nsb_sem_t sem, int slot, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout,
                      int field0, int nf, double scale, int accumulate);
/* EXT / BDF sums of the same step ([UPSTREAM-RECALL] perturb.f makextp + makebdfp) in one pass over
 * fields field0..field0+nf-1 of columns of b:
 *   ta = ab[1] e1 + ab[2] e2 ; e2 <- e1 ; e1 <- bf ; bf <- ab[0] bf + ta ;
 *   bf += rho_over_dt * bm1 * sum_{i<nbd} bd[i+1] * v(col_vlag[i])     (col_vlag[0] = current velocity) */
int nsb_sem_bdf_ext(nsb_sem_t sem, nsb_basis_t b, int col_bf, int col_e1, int col_e2, const int *col_vlag,
                    int nbd, int field0, int nf, const double *ab, const double *bd, double rho_over_dt);

/* ---------------------------------------------------------------------------------------------
 * Linear operator: the abstract_linop%matvec(vec_in, vec_out) boundary
 * (core/linear_operators.f90:17-23, 39-44) / legacy matvec(f, q) (core/matvec.f90:56).
 * ------------------------------------------------------------------------------------------- */
/* Built-in device operator on every velocity-like field f in [0, nfields_apply):
 *   out = alpha * in + beta * binvm1 * mask * dssum( h1 * A in + h2 * B in  [+ B (U . grad) in] )
 * convection velocity (cx,cy,cz element-local arrays, or NULL for none) makes it non-symmetric
 * (SURVEY.md section 8d operators M1 / M2). */
int nsb_op_create_sem(nsb_sem_t sem, int nfields_apply, double alpha, double beta, double h1,
                      double h2, const double *cx, const double *cy, const double *cz,
                      nsb_op_t *op);
/* Host operator: the reference's time-stepper (nek_advance on host arrays vxp, vyp ...).  The
 * library downloads vec_in into pinned staging buffers, calls the callback, and uploads vec_out.
 * out_fields[i] arrives pointing at a pinned staging buffer the callback may fill; it may instead
 * overwrite out_fields[i] with a pointer to its own array (zero-copy from Nek's vxp ...). */
typedef int (*nsb_host_matvec_fn)(void *user, const double *const *in_fields, double in_time,
                                  double **out_fields, double *out_time);
int nsb_op_create_host(nsb_layout_t layout, nsb_host_matvec_fn fn, void *user, nsb_op_t *op);
/* Declare a host operator LINEAR (M(a x) = a M(x) for every component, %time included): true for the linearised
 * time-steppers of the eigenvalue / transient-growth analyses (core/linear_operators.f90:39-103), false for the
 * nonlinear forward map of newton_krylov.  nsb_arnoldi then starts the download of q_m+1 while the last sweep
 * of step m is still running: the callback receives the UN-NORMALISED vector beta q_m+1 and the library divides
 * the returned vector by beta on the device.  H and the basis are the same to rounding. */
int nsb_op_set_linear(nsb_op_t op, int linear);
/* out = outer(inner(in)): the reference's composite maps are built this way from the basic solvers,
 * e.g. transient_growth_map = adjoint_linearized_map(forward_linearized_map(q))
 * (core/matvec.f90:478-495).  The component operators stay owned by the caller. */
int nsb_op_create_compose(nsb_layout_t layout, nsb_op_t outer, nsb_op_t inner, nsb_op_t *op);
/* out = alpha A(in) + beta B(in); A or B = NULL stands for the identity.  LightKrylov's axpby_linop / identity_linop
 * as the reference combines them for the resolvent's S = I - exp(TL) (core/linear_operators.f90:364-403), and the
 * legacy maps made from the basic solvers with k_sub2 / k_cmult: newton_linearized_map = exp(TL) - I
 * (core/matvec.f90:520-541; what ts_gmres / newton_krylov iterate on) and ts_force_sensitivity_map = I - exp(TL)^+
 * (core/matvec.f90:499-516).  All fields and %time take part, like k_sub2.  A and B stay owned by the caller. */
int nsb_op_create_axpby(nsb_layout_t layout, nsb_op_t A, nsb_op_t B, double alpha, double beta, nsb_op_t *op);
/* forward_finite_difference_map (core/matvec.f90:246-379, iffindiff): the linearised forward map as finite differences
 * of a NONLINEAR map F (any operator handle; in the reference the nonlinear Nek stepper, i.e. a host callback) about
 * the base state X = column col_base of base:  f = (1/eps0) sum_i coef_i F(X + amp_i eps0 q),  eps0 = 1e-6 |X|;
 * order (findiff_order) 2: amp (1,-1), coef (1,-1)/2;  4: amp (1,-1,2,-2), coef (8,-8,-1,1)/12.  X is read at every
 * application and stays owned by the caller; combine with nsb_op_create_axpby for newton_linearized_map. */
int nsb_op_create_frechet_fd(nsb_layout_t layout, nsb_op_t F, nsb_basis_t base, int col_base, int order, nsb_op_t *op);
/* epsilon_base (core/main.f90:16, default 1e-6; the new API's eps0 = epsilon_base |X|, core/linear_operators.f90:192). */
int nsb_op_frechet_set_epsilon(nsb_op_t op, double epsilon_base);
/* Device time-stepper operator, the structure of exponential_prop%matvec
 * (core/linear_operators.f90:225-274: integrate over tau from a cold start, return the final state) for
 * Nek's scalar step cdscal [UPSTREAM-RECALL]: nsteps BDF/EXT steps (order ramp 1, 2, 3) of
 *     rho dT/dt + rho (U.grad) T = kappa lap T,   T = 0 where the mesh mask is 0,
 * applied independently to the first nfields_apply fields; U = the convecting field in `slot`
 * (nsb_sem_set_convect; -1: no convection).  Each step = nsb_sem_convect, nsb_sem_bdf_ext,
 * nsb_sem_dssum, nsb_sem_hmholtz(kappa, rho bd1/dt, tol, maxit).  The pressure-coupled velocity
 * step is nsb_op_create_ns_stepper below. */
int nsb_op_create_stepper(nsb_sem_t sem, nsb_layout_t layout, int nfields_apply, int slot, double kappa,
                          double rho, double dt, int nsteps, double tol, int maxit, nsb_op_t *op);
/* exponential_prop%rmatvec (core/linear_operators.f90:84-103) for the same step sequence: the DISCRETE adjoint of
 * the operator nsb_op_create_stepper builds with respect to the BM1 inner product,
 *     <A u, v>_B = <u, A^+ v>_B   for continuous, masked u, v  (to rounding and the Helmholtz tolerance),
 * obtained by running the transposed BDF/EXT recurrence backwards in time (transposed dealiased convection,
 * the same symmetric Helmholtz solves).  Together they make transient_growth_map = A^+ A (core/matvec.f90:478-495,
 * nsb_op_create_compose) and nsb_svds device-resident.  Same arguments as nsb_op_create_stepper. */
int nsb_op_create_stepper_adjoint(nsb_sem_t sem, nsb_layout_t layout, int nfields_apply, int slot, double kappa,
                                  double rho, double dt, int nsteps, double tol, int maxit, nsb_op_t *op);
/* ---- pressure-coupled perturbation step (the body of nek_advance reached from exponential_prop%matvec,
 * core/linear_operators.f90:225-274, :247) for Nek5000's P_N - P_N-2 formulation [UPSTREAM-RECALL perturb.f
 * perturbv / incomprp, navier1.f opdiv / opgradt / cdabdtp / uzawa, coef.f geom2; parity unpinned] -------------
 * Velocity on the lx1 = N+1 GLL mesh (fields 0..dim-1 of a column, equally long), pressure on lx2 = lx1-2
 * Gauss-Legendre points per direction and element (field dim of the column, nsb_sem_npres() values; nekStab keeps it
 * in the Krylov vector and out of the inner product, core/nek_vectors.f90:20-31).
 *   pressure_matrices : host-only; z2, w2 [lx2], I12, D12 [lx2][lx1] row-major (ixm12, dxm12)
 *   pressure_setup    : metrics on the pressure mesh (rxm2 = w3m2 map12(rxm1), bm2) -- needs N >= 3
 *   pressure_get      : which 0: rx2 [dim*dim][n2], 1: 1 / bm2 [n2]
 *   opdiv             : out(pressure) = D in(velocity),  (D u)_q = w_q sum_ab (J dr_a/dx_b)_q (du_b/dr_a)_q
 *   opgradt           : out(velocity, element-local, no dssum) = D^T in(pressure)
 *   cdabdtp           : out = E in,  E = D (binvm1 mask QQ^T) D^T  (consistent Poisson operator)
 *   esolve            : E x = rhs by CG preconditioned with 1 / bm2 (uzawa without the Schwarz part); stops when
 *                       sqrt(r.z) <= tol sqrt(r0.z0); mean_free != 0 removes the mean of the right-hand side and of
 *                       every preconditioned residual (Nek's ortho for all-Dirichlet velocity; exact on affine
 *                       elements, where E 1 = 0 -- on deformed elements E is regular and mean_free = 0 solves it as
 *                       it stands).  CG scalars stay on the device; the host polls a flag every 8 iterations.
 *                       precond 0: 1 / bm2 (uzprec without the Schwarz part; O(1000) iterations on large meshes).
 *                       precond 1: two levels -- element-wise fast-diagonalisation solves (the local solves of Nek's
 *                       Schwarz preconditioner, fast.f / hsmg.f, without overlap: every element replaced by the box
 *                       with its mean edge lengths) plus a coarse correction on one constant per element
 *                       (E_c = R E R^T, sparse, Jacobi-CG on the device); iteration count independent of the number
 *                       of elements (about 60 at tol 1e-8).  On a multi-rank context the coarse level couples the
 *                       elements of a rank only. */
int nsb_pressure_matrices(int N, double *z2, double *w2, double *I12, double *D12);
/* host-only: 1-D generalised eigenpairs of the element-wise pressure solves (precond 1): E^ S = M^ S Lambda,
 * S^T M^ S = 1, S [lx2][lx2] row-major (column m = eigenvector m), lam [lx2] */
int nsb_fdm_matrices(int N, double *S, double *lam);
int nsb_sem_pressure_setup(nsb_sem_t sem);
int64_t nsb_sem_npres(nsb_sem_t sem);
int nsb_sem_pressure_get(nsb_sem_t sem, int which, double *out);
int nsb_sem_opdiv(nsb_sem_t sem, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);
int nsb_sem_opgradt(nsb_sem_t sem, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);
int nsb_sem_cdabdtp(nsb_sem_t sem, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);
int nsb_sem_esolve(nsb_sem_t sem, nsb_basis_t brhs, int crhs, nsb_basis_t bx, int cx, double tol, int maxit,
                   int mean_free, int precond, int *iters, double *res);
/* norm_grad (core/utils.f90:446-486): sum_c sum_b glsc3(du_c/dx_b, bm1s, du_c/dx_b) of the velocity fields 0..dim-1 of
 * (b, col), gradm1 collocation derivatives, the layout's weight, summed over all ranks; NOT a square root -- the
 * number outpost_ks compares with 1.1 to skip spurious Ritz vectors (core/eigensolvers.f90:587-594). */
int nsb_sem_norm_grad(nsb_sem_t sem, nsb_basis_t b, int col, double *norma);
/* compute_cfl(cfl, vx, vy, vz, dt) ([UPSTREAM-RECALL] Nek5000; call sites core/linear_stab.f90:222,231) of the velocity
 * fields of (b, col): max over the GLL points of dt (|u.grad r| / dr_i + |u.grad s| / ds_j [+ |u.grad t| / dt_k]), over
 * all ranks.  set_linear_solver derives the stepper's dt and nsteps from it: dt = ctarg / cfl(dt = 1),
 * nsteps = ceiling(T / dt), dt = T / nsteps (core/linear_stab.f90:220-236). */
int nsb_sem_cfl(nsb_sem_t sem, nsb_basis_t b, int col, double dt, double *cfl);
/* exponential_prop%matvec for the linearised incompressible Navier-Stokes equations, device-resident:
 *     dv/dt + (U.grad) v + (v.grad) U = -grad p + nu lap v,   div v = 0,   v = 0 where the mesh mask is 0.
 * The input vector's velocity and pressure start nsteps BDF/EXT steps (order ramp 1, 2, 3 -- the reference restarts
 * the time-stepper for every matvec); per step: advabp (dealiased, both convection slots of the mesh are used),
 * makextp / makebdfp, H v* = QQ^T (bf + D^T p*), E dp = -(bd0/dt) D v*, v = v* + (dt/bd0) B^-1 D^T dp, p = p* + dp
 * with p* = p^(n-1) (third step on: 2 p^(n-1) - p^(n-2)).  base / col_base: the base flow U (velocity fields of that
 * column, same layout; nsb_sem_dealias_setup first), NULL = Stokes.  The output vector receives the final velocity
 * and pressure; other fields and %time are carried through.  nsb_op_ns_iterations: Helmholtz / pressure iterations
 * spent so far. */
int nsb_op_create_ns_stepper(nsb_sem_t sem, nsb_layout_t layout, nsb_basis_t base, int col_base, double nu, double dt,
                             int nsteps, double tol_v, double tol_p, int maxit, int mean_free, int precond,
                             nsb_op_t *op);
/* exponential_prop%rmatvec (core/linear_operators.f90:84-103) for the same equations: Nek's stepper in adjoint mode --
 * the same splitting on the continuous adjoint equations, explicit term +(U.grad) w - sum_c w_c grad U_c (transport
 * term dealiased, base-flow-gradient term pointwise with gradm1 of U).  Dual to the forward operator up to the
 * discretisation error (first order in dt), as in the reference; with nsb_op_create_compose it gives the
 * transient-growth map of core/matvec.f90:478-495 for the Navier-Stokes equations, device-resident.  Same arguments. */
int nsb_op_create_ns_stepper_adjoint(nsb_sem_t sem, nsb_layout_t layout, nsb_basis_t base, int col_base, double nu,
                                     double dt, int nsteps, double tol_v, double tol_p, int maxit, int mean_free,
                                     int precond, nsb_op_t *op);
/* Time-periodic base flow (Floquet analysis / Newton for periodic orbits): the stored orbit uor / vor / wor(:, istep) of
 * core/linear_operators.f90:254-275 and core/matvec.f90:347-362 as a run of columns of a device basis with the
 * operator's layout -- step n of every application linearises about column col0 + (n-1) stride (the reference's step
 * istep runs with uor(:, istep-1), uor(:, 0) = ubase; stride -1 walks the orbit backwards).  For operators created
 * with a base flow, forward or adjoint; orbit = NULL returns to the steady base flow.  The basis stays the caller's. */
int nsb_op_ns_set_orbit(nsb_op_t op, nsb_basis_t orbit, int col0, int stride);
int nsb_op_ns_iterations(nsb_op_t op, int64_t *helmholtz, int64_t *pressure);
int nsb_op_destroy(nsb_op_t op);
int nsb_op_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);
int nsb_op_count(nsb_op_t op, int64_t *napply);

/* ---------------------------------------------------------------------------------------------
 * Krylov drivers (host logic in C++, dense k x k step through the injected LAPACK).
 * ------------------------------------------------------------------------------------------- */
/* arnoldi_factorization(Q, H, mstart, mend, ksize) (core/krylov_decomposition.f90:2-99).
 * Steps mstart..mend (0-based, inclusive): f = op(Q[m]); orthonormalise against Q[0..m];
 * Q[m+1] = f; column m of H (ldh >= mend+2) is written.  H is host memory; the loop itself is
 * device-resident (no host synchronisation until the end). */
int nsb_arnoldi(nsb_basis_t Q, nsb_op_t op, int mstart, int mend, int orth_mode, double *H,
                int ldh);
/* Projection passes the orthogonalisation took in steps mstart..mend of the last device-resident nsb_arnoldi
 * on this basis' context: 1 or 2 per step for NSB_ORTH_DGKS (the decision is taken on the device), 2 otherwise. */
int nsb_arnoldi_passes(nsb_basis_t Q, int mstart, int mend, int orth_mode, int *passes);

/* LAPACK provider: raw Fortran-ABI entry points, as core/lapack_wrapper.f90 links them
 * (dgeev :158, dgees :49, dtrsen :108, dgels :288).  The Fortran host passes c_funloc(dgeev)...;
 * Python passes scipy's cython_lapack pointers. */
int nsb_set_lapack(void *dgeev, void *dgees, void *dtrsen, void *dgels);
/* dgesvd for nsb_svd / nsb_svds (LightKrylov's svd wrapper calls the same routine). */
int nsb_set_lapack_svd(void *dgesvd);
/* Thin SVD A(m x n) = U diag(S) V^T; U is m x min(m,n), V is n x min(m,n) (untransposed). */
int nsb_svd(const double *A, int lda, int m, int n, double *U, double *S, double *V);
/* lapack_wrapper mirrors (host, column-major):
 *   eig   : dgeev + complexification + sort by decreasing |lambda| (:114-228);
 *           vals/vecs interleaved (re,im) complex*16
 *   schur : dgees('V','S', |lambda|>0.9) (:3-55), A overwritten by T
 *   ordschur : dtrsen (:59-111);  lstsq : dgels (:248-300) */
int nsb_eig(const double *A, int lda, int n, double *vecs_c16, double *vals_c16);
int nsb_schur(double *A, int lda, int n, double *vecs, double *vals_c16);
int nsb_ordschur(double *T, int ldt, double *Q, int ldq, const int *selected, int n);
int nsb_lstsq(const double *A, int lda, int m, int n, const double *b, double *x);
/* select_eigenvalues (core/eigensolvers.f90:688-754) */
int nsb_select_eigenvalues(int *selected, int *cnt, const double *vals_c16, double delta, int nev,
                           int n);
/* schur_condensation(mstart, H, Q, ksize) (core/eigensolvers.f90:363-468); mstart in/out, 0-based
 * index of the next Arnoldi step. */
int nsb_schur_condensation(nsb_basis_t Q, int *mstart, double *H, int ldh, int ksize,
                           double schur_del, int schur_tgt);
/* krylov_schur (core/eigensolvers.f90:120-359): Q[0] must hold the unit-norm seed.
 * Outputs: vals/vecs of H(1:k,1:k) sorted as eig() does, residual(k), number converged,
 * number of Schur condensations, H ((k+1) x k, ldh). */
int nsb_krylov_schur(nsb_basis_t Q, nsb_op_t op, int k_dim, int schur_tgt, double eigen_tol,
                     double schur_del, int orth_mode, int max_restarts, double *H, int ldh,
                     double *vals_c16, double *vecs_c16, double *residual, int *cnt,
                     int *schur_cnt);
/* Step-wise eigensolver of the LightKrylov path, the call linear_stability_analysis makes
 * (core/linear_stab.f90:66: eigs(A, X, eigvecs, eigvals, residuals, info, nev, tolerance)):
 * one Arnoldi step at a time, eig(H(1:k,1:k)) and residuals |H(k+1,k) y_k| after every step, stop
 * when nev Ritz pairs are below tol.  [UPSTREAM-RECALL: LightKrylov is not vendored.]  vecs_c16 has
 * leading dimension k_dim; *kused = Krylov dimension reached. */
int nsb_eigs(nsb_basis_t Q, nsb_op_t op, int k_dim, int nev, double tol, int orth_mode, double *H,
             int ldh, double *vals_c16, double *vecs_c16, double *residual, int *kused, int *nconv);
/* newton_krylov (core/newton_krylov.f90:1-168): f = F(q); residual = |f|^2; stop when residual < tol;
 * dq = ts_gmres(J, f); q -= dq.  fop applies the (nonlinear) forward map F, jop the linearisation about the
 * current q -- in the reference both are the host time-stepper, i.e. host-callback operators whose owner
 * re-linearises when F is called.  (bw, cf), (bw, cdq): two work vectors; Q: GMRES basis with >= ksize+2
 * columns; residual_hist: maxiter_newton entries; *iters = Newton iterations performed. */
int nsb_newton_krylov(nsb_basis_t Q, nsb_op_t fop, nsb_op_t jop, nsb_basis_t bq, int cq, nsb_basis_t bw, int cf,
                      int cdq, int maxiter_newton, int maxiter_gmres, int ksize, double tol, int orth_mode,
                      int *iters, double *residual_hist, int *calls);
/* Ritz-vector assembly (core/eigensolvers.f90:565-585, 609-615; get_vec, core/linear_stab.f90:362):
 * fp = Q(:,1:k) y for complex y (interleaved re,im): column cre <- Q Re(y), column cim <- Q Im(y) of bout;
 * alpha_re / alpha_im = their BM1 norms (before scaling); normalize != 0 scales both parts by
 * 1 / sqrt(alpha_re^2 + alpha_im^2) as the reference does before writing the mode. */
int nsb_ritz_vector(nsb_basis_t Q, int k, const double *y_c16, nsb_basis_t bout, int cre, int cim,
                    int normalize, double *alpha_re, double *alpha_im);
/* Step-wise singular-value solver, the call transient_growth_analysis makes
 * (core/linear_stab.f90:112: svds(A, U, V, uvecs, vvecs, sigma, residuals, info, nev, tolerance)):
 * Golub-Kahan bidiagonalisation with full re-orthogonalisation, v_k = A^T u_k, u_k+1 = A v_k,
 * B(k,k) = |v_k|, B(k+1,k) = |u_k+1|, svd(B(1:k,1:k)) and residuals |B(k+1,k) vvecs(k,:)| after
 * every step.  [UPSTREAM-RECALL: LightKrylov is not vendored.]  U[0] must hold the unit-norm seed;
 * U needs k_dim+1 columns, V k_dim.  op_adj applies the adjoint (A%rmatvec).  B is (k_dim+1) x k_dim
 * (ldb); uvecs / vvecs have leading dimension k_dim; *kused = Krylov dimension reached. */
int nsb_svds(nsb_basis_t U, nsb_basis_t V, nsb_op_t op, nsb_op_t op_adj, int k_dim, int nev, double tol,
             int orth_mode, double *B, int ldb, double *sigma, double *uvecs, double *vvecs,
             double *residual, int *kused, int *nconv);
/* ts_gmres(rhs, sol, maxiter, ksize, calls) (core/newton_krylov.f90:170-299).  rhs and sol are
 * (basis, col) vectors; Q is the caller's Krylov basis with >= ksize+2 columns (last = work). */
int nsb_ts_gmres(nsb_basis_t Q, nsb_op_t op, nsb_basis_t brhs, int crhs, nsb_basis_t bsol, int csol,
                 int maxiter, int ksize, double tol, int orth_mode, int *calls,
                 double *residual_hist, int *nhist);

/* ---------------------------------------------------------------------------------------------
 * Krylov checkpoint / restart wire formats of the reference (host files <-> device basis).
 * ------------------------------------------------------------------------------------------- */
/* HES<session>%04d: H(1:k+1,1:k) row by row as list-directed text (core/eigensolvers.f90:837-843). */
int nsb_hessenberg_write(const char *path, const double *H, int ldh, int k);
/* The restart read of core/eigensolvers.f90:246-266: the file holds (mstart+1) x mstart values; the leading
 * block of the (k_dim+1) x k_dim matrix H is filled (subsampled to k_dim columns when k_dim < mstart). */
int nsb_hessenberg_read(const char *path, int k_dim, int mstart, double *H, int ldh);
/* One Nek5000 field file (KRY<session>0.f%05d written by outpost2, core/eigensolvers.f90:803-809) into
 * column col: velocity -> layout fields ufield0.., pressure -> pfield, temperature -> tfield (-1: skip).
 * lglel = 1-based global ids of this rank's elements (Nek's LGLEL; NULL: the file's first nel_local elements).
 * %time of the column is left zero like load_files does (core/IO.f90:60-68); the header's time -> *time_out. */
int nsb_fld_read_into(nsb_basis_t b, int col, const char *path, const int64_t *lglel, int64_t nel_local,
                      int ufield0, int pfield, int tfield, double *time_out);
/* The restart branch of krylov_schur (core/eigensolvers.f90:240-285): H from HES<session><mstart>, Krylov
 * vectors 1..mstart+1 from KRY<session>0.f00001.. into columns 0..mstart; *mstart_next = 0-based index of the
 * next Arnoldi step (pass it to nsb_arnoldi / continue nsb_krylov_schur's loop from there). */
int nsb_restart_load(nsb_basis_t Q, const char *dir, const char *session, int mstart, int k_dim,
                     const int64_t *lglel, int64_t nel_local, int ufield0, int pfield, int tfield, double *H,
                     int ldh, int *mstart_next);

#ifdef __cplusplus
}
#endif
#endif /* NEKSTAB_B200_H */
