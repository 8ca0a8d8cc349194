// Host-side Krylov drivers over the device-resident basis: Arnoldi, Krylov-Schur, GMRES, and the
// lapack_wrapper mirror.  The k x k dense step stays on the host through LAPACK, as in the
// reference (core/lapack_wrapper.f90); the LAPACK entry points are injected by the host program
// (nsb_set_lapack) because this library must not depend on a particular BLAS/LAPACK build.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <numeric>
#include <tuple>
#include <vector>

#include "nsb_internal.h"

using namespace nsb;

// ------------------------------------------------------------------------------------------------
// LAPACK provider (Fortran ABI; trailing hidden character lengths are passed and ignored by C
// wrappers such as scipy's cython_lapack)
// ------------------------------------------------------------------------------------------------
namespace {
typedef int (*select2_fn)(double *, double *);
typedef void (*dgeev_fn)(char *, char *, int *, double *, int *, double *, double *, double *, int *,
                         double *, int *, double *, int *, int *, size_t, size_t);
typedef void (*dgees_fn)(char *, char *, select2_fn, int *, double *, int *, int *, double *, double *,
                         double *, int *, double *, int *, int *, int *, size_t, size_t);
typedef void (*dtrsen_fn)(char *, char *, int *, int *, double *, int *, double *, int *, double *,
                          double *, int *, double *, double *, double *, int *, int *, int *, int *,
                          size_t, size_t);
typedef void (*dgels_fn)(char *, int *, int *, int *, double *, int *, double *, int *, double *, int *,
                         int *, size_t);
dgeev_fn p_dgeev = nullptr;
dgees_fn p_dgees = nullptr;
dtrsen_fn p_dtrsen = nullptr;
typedef void (*dgesvd_fn)(char *, char *, int *, int *, double *, int *, double *, double *, int *, double *,
                          int *, double *, int *, int *, size_t, size_t);
dgels_fn p_dgels = nullptr;
dgesvd_fn p_dgesvd = nullptr;

int need_lapack(void *p, const char *name) {
  if (!p) {
    set_error("%s: no LAPACK provider registered; call nsb_set_lapack first", name);
    return NSB_ELAPACK;
  }
  return NSB_OK;
}

// core/lapack_wrapper.f90:232-244
int select_eigvals(double *wr, double *wi) { return std::sqrt(*wr * *wr + *wi * *wi) > 0.9; }
}  // namespace

extern "C" int nsb_set_lapack(void *dgeev, void *dgees, void *dtrsen, void *dgels) {
  p_dgeev = (dgeev_fn)dgeev;
  p_dgees = (dgees_fn)dgees;
  p_dtrsen = (dtrsen_fn)dtrsen;
  p_dgels = (dgels_fn)dgels;
  return NSB_OK;
}

extern "C" int nsb_set_lapack_svd(void *dgesvd) {
  p_dgesvd = (dgesvd_fn)dgesvd;
  return NSB_OK;
}

// Thin SVD A = U diag(S) V^T (dgesvd, 'S','S'); V is returned untransposed, as LightKrylov's svd
// wrapper hands it to svds.  U: m x min(m,n) (ldu = m), V: n x min(m,n) (ldv = n).
extern "C" int nsb_svd(const double *A, int lda, int m, int n, double *U, double *S, double *V) {
  NSB_REQUIRE(A && U && S && V && m >= 1 && n >= 1 && lda >= m, "nsb_svd: bad argument");
  NSB_CHECK(need_lapack((void *)p_dgesvd, "nsb_svd"));
  const int r = std::min(m, n);
  std::vector<double> At((size_t)m * n), vt((size_t)r * n);
  for (int j = 0; j < n; ++j) memcpy(&At[(size_t)j * m], A + (size_t)j * lda, sizeof(double) * m);
  char jobu = 'S', jobvt = 'S';
  int mm = m, nn = n, ldu = m, ldvt = r, lwork = -1, info = 0;
  double wq = 0;
  p_dgesvd(&jobu, &jobvt, &mm, &nn, At.data(), &mm, S, U, &ldu, vt.data(), &ldvt, &wq, &lwork, &info, 1, 1);
  lwork = std::max(std::max(3 * r + std::max(m, n), 5 * r), (int)wq);
  std::vector<double> work(lwork);
  p_dgesvd(&jobu, &jobvt, &mm, &nn, At.data(), &mm, S, U, &ldu, vt.data(), &ldvt, work.data(), &lwork, &info, 1,
           1);
  if (info != 0) {
    set_error("nsb_svd: dgesvd info=%d", info);
    return NSB_ELAPACK;
  }
  for (int j = 0; j < r; ++j)
    for (int i = 0; i < n; ++i) V[(size_t)j * n + i] = vt[(size_t)i * r + j];
  return NSB_OK;
}

// core/lapack_wrapper.f90:114-228
extern "C" int nsb_eig(const double *A, int lda, int n, double *vecs_c16, double *vals_c16) {
  NSB_REQUIRE(A && vecs_c16 && vals_c16 && n >= 1 && lda >= n, "nsb_eig: bad argument");
  NSB_CHECK(need_lapack((void *)p_dgeev, "nsb_eig"));
  std::vector<double> At((size_t)n * n), wr(n), wi(n), vr((size_t)n * n), vl(1);
  for (int j = 0; j < n; ++j) memcpy(&At[(size_t)j * n], A + (size_t)j * lda, sizeof(double) * n);
  char jobvl = 'N', jobvr = 'V';
  int ldvl = 1, ldvr = n, lwork = -1, info = 0, nn = n;
  double wq = 0;
  p_dgeev(&jobvl, &jobvr, &nn, At.data(), &nn, wr.data(), wi.data(), vl.data(), &ldvl, vr.data(), &ldvr,
          &wq, &lwork, &info, 1, 1);
  lwork = std::max(4 * n, (int)wq);
  std::vector<double> work(lwork);
  p_dgeev(&jobvl, &jobvr, &nn, At.data(), &nn, wr.data(), wi.data(), vl.data(), &ldvl, vr.data(), &ldvr,
          work.data(), &lwork, &info, 1, 1);
  if (info != 0) {
    set_error("nsb_eig: dgeev info=%d", info);
    return NSB_ELAPACK;
  }
  typedef std::complex<double> cd;
  std::vector<cd> vals(n), vecs((size_t)n * n);
  for (int i = 0; i < n; ++i) vals[i] = cd(wr[i], wi[i]);
  for (size_t t = 0; t < (size_t)n * n; ++t) vecs[t] = cd(vr[t], 0.0);
  for (int i = 0; i < n - 1; ++i) {  // :167-173
    if (wi[i] > 0) {
      for (int r = 0; r < n; ++r) {
        vecs[(size_t)i * n + r] = cd(vr[(size_t)i * n + r], vr[(size_t)(i + 1) * n + r]);
        vecs[(size_t)(i + 1) * n + r] = cd(vr[(size_t)i * n + r], -vr[(size_t)(i + 1) * n + r]);
      }
    } else if (wi[i] == 0) {
      for (int r = 0; r < n; ++r) vecs[(size_t)i * n + r] = cd(vr[(size_t)i * n + r], 0.0);
    }
  }
  // sort_eigendecomp :181-228 (exchange sort, decreasing magnitude)
  std::vector<double> nrm(n);
  for (int i = 0; i < n; ++i) nrm[i] = std::sqrt(vals[i].real() * vals[i].real() + vals[i].imag() * vals[i].imag());
  for (int k = 0; k < n - 1; ++k)
    for (int l = k + 1; l < n; ++l)
      if (nrm[k] < nrm[l]) {
        std::swap(nrm[k], nrm[l]);
        std::swap(vals[k], vals[l]);
        for (int r = 0; r < n; ++r) std::swap(vecs[(size_t)k * n + r], vecs[(size_t)l * n + r]);
      }
  memcpy(vals_c16, vals.data(), sizeof(cd) * n);
  memcpy(vecs_c16, vecs.data(), sizeof(cd) * n * n);
  return NSB_OK;
}

// core/lapack_wrapper.f90:3-55
extern "C" int nsb_schur(double *A, int lda, int n, double *vecs, double *vals_c16) {
  NSB_REQUIRE(A && vecs && vals_c16 && n >= 1 && lda >= n, "nsb_schur: bad argument");
  NSB_CHECK(need_lapack((void *)p_dgees, "nsb_schur"));
  char jobvs = 'V', sort = 'S';
  int nn = n, ld = lda, sdim = 0, ldvs = n, lwork = -1, info = 0;
  std::vector<double> wr(n), wi(n);
  std::vector<int> bwork(n);
  double wq = 0;
  p_dgees(&jobvs, &sort, select_eigvals, &nn, A, &ld, &sdim, wr.data(), wi.data(), vecs, &ldvs, &wq,
          &lwork, bwork.data(), &info, 1, 1);
  lwork = std::max(3 * n, (int)wq);
  std::vector<double> work(lwork);
  p_dgees(&jobvs, &sort, select_eigvals, &nn, A, &ld, &sdim, wr.data(), wi.data(), vecs, &ldvs,
          work.data(), &lwork, bwork.data(), &info, 1, 1);
  // info = n+2 only says the reordered eigenvalues changed their selection status after
  // roundoff; the factorisation is still valid (the reference ignores info altogether).
  if (info != 0 && info != n + 2) {
    set_error("nsb_schur: dgees info=%d", info);
    return NSB_ELAPACK;
  }
  for (int i = 0; i < n; ++i) {
    vals_c16[2 * i] = wr[i];
    vals_c16[2 * i + 1] = wi[i];
  }
  return NSB_OK;
}

// core/lapack_wrapper.f90:59-111
extern "C" int nsb_ordschur(double *T, int ldt, double *Q, int ldq, const int *selected, int n) {
  NSB_REQUIRE(T && Q && selected && n >= 1 && ldt >= n && ldq >= n, "nsb_ordschur: bad argument");
  NSB_CHECK(need_lapack((void *)p_dtrsen, "nsb_ordschur"));
  char job = 'N', compq = 'V';
  int nn = n, lt = ldt, lq = ldq, m = 0, lwork = std::max(1, n), liwork = 1, info = 0;
  double s = 0, sep = 0;
  std::vector<double> wr(n), wi(n), work(lwork);
  std::vector<int> sel(selected, selected + n), iwork(1);
  p_dtrsen(&job, &compq, sel.data(), &nn, T, &lt, Q, &lq, wr.data(), wi.data(), &m, &s, &sep,
           work.data(), &lwork, iwork.data(), &liwork, &info, 1, 1);
  if (info != 0) {
    set_error("nsb_ordschur: dtrsen info=%d", info);
    return NSB_ELAPACK;
  }
  return NSB_OK;
}

// core/lapack_wrapper.f90:248-300
extern "C" int nsb_lstsq(const double *A, int lda, int m, int n, const double *b, double *x) {
  NSB_REQUIRE(A && b && x && m >= n && n >= 1 && lda >= m, "nsb_lstsq: bad argument");
  NSB_CHECK(need_lapack((void *)p_dgels, "nsb_lstsq"));
  std::vector<double> At((size_t)m * n), bt(b, b + m);
  for (int j = 0; j < n; ++j) memcpy(&At[(size_t)j * m], A + (size_t)j * lda, sizeof(double) * m);
  char trans = 'N';
  int mm = m, nn = n, nrhs = 1, ldb = m, lwork = -1, info = 0;
  double wq = 0;
  p_dgels(&trans, &mm, &nn, &nrhs, At.data(), &mm, bt.data(), &ldb, &wq, &lwork, &info, 1);
  lwork = std::max(2 * m * n, (int)wq);
  std::vector<double> work(lwork);
  p_dgels(&trans, &mm, &nn, &nrhs, At.data(), &mm, bt.data(), &ldb, work.data(), &lwork, &info, 1);
  if (info != 0) {
    set_error("nsb_lstsq: dgels info=%d", info);
    return NSB_ELAPACK;
  }
  memcpy(x, bt.data(), sizeof(double) * n);
  return NSB_OK;
}

// core/eigensolvers.f90:688-754
extern "C" int nsb_select_eigenvalues(int *selected, int *cnt, const double *vals_c16, double delta,
                                      int nev, int n) {
  NSB_REQUIRE(selected && cnt && vals_c16, "nsb_select_eigenvalues: NULL argument");
  NSB_REQUIRE(nev >= 0 && n >= nev + 5, "nsb_select_eigenvalues: need n >= nev+5 (n=%d, nev=%d)", n, nev);
  std::vector<double> mag(n);
  for (int i = 0; i < n; ++i) mag[i] = std::hypot(vals_c16[2 * i], vals_c16[2 * i + 1]);
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return mag[a] < mag[b]; });  // quicksort2, ascending
  for (int i = 0; i < n; ++i) selected[i] = mag[i] >= (1.0 - delta);                         // :743
  for (int q = n - (nev + 3) - 1; q < n; ++q) selected[idx[q]] = 1;                          // :746
  const double a = vals_c16[2 * idx[n - (nev + 3) - 1] + 1], b = vals_c16[2 * idx[n - (nev + 4) - 1] + 1];
  if (a == -b) selected[idx[n - (nev + 4) - 1]] = 1;                                         // :747-749
  int c = 0;
  for (int i = 0; i < n; ++i) c += selected[i] ? 1 : 0;
  *cnt = c;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Arnoldi  (core/krylov_decomposition.f90:2-99)
// ------------------------------------------------------------------------------------------------
namespace {

// One Arnoldi step {f = M q_m ; orthonormalise ; H column -> pinned staging} enqueued on the context stream.
int enqueue_step(nsb_basis_t Q, nsb_op_t op, int m, int orth_mode, double *h_dst) {
  NSB_CHECK(nsb_op_apply(op, Q, m, Q, m + 1));   // f written straight into column m+1 (saves the k_copy of :81)
  return nsb_orthonormalize_async(Q, m + 1, m + 1, orth_mode, h_dst);
}

// The same step as a CUDA graph, captured the first time (basis, operator, m, mode) is seen and replayed
// afterwards: everything a step needs at run time -- all-reduce sequence numbers, the DGKS decision, the
// halo-exchange slots -- lives in device memory, so the captured kernels are valid for every replay.
int step_graph(nsb_basis_t Q, nsb_op_t op, int m, int orth_mode, double *h_dst) {
  nsb_context_t ctx = Q->lay->ctx;
  const auto key = std::make_tuple((const void *)Q, (const void *)op, m, orth_mode);
  auto it = ctx->step_graphs.find(key);
  if (it == ctx->step_graphs.end()) {
    cudaSetDevice(ctx->device);
    const int64_t l0 = ctx->launches;
    const int64_t n0 = op->napply;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      ctx->use_graph = false;
      return enqueue_step(Q, op, m, orth_mode, h_dst);
    }
    const int rc = enqueue_step(Q, op, m, orth_mode, h_dst);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    const int64_t nl = ctx->launches - l0;
    ctx->launches = l0;
    op->napply = n0;
    cudaGraphExec_t exec = nullptr;
    if (rc == NSB_OK && e == cudaSuccess && g && cudaGraphInstantiate(&exec, g, 0) == cudaSuccess) {
      nsb_context_s::StepGraph sg;
      sg.exec = exec;
      sg.launches = nl;
      it = ctx->step_graphs.emplace(key, sg).first;
    }
    if (g) cudaGraphDestroy(g);
    if (it == ctx->step_graphs.end()) {   // capture not possible here: plain launches from now on
      cudaGetLastError();
      if (rc != NSB_OK) return rc;
      ctx->use_graph = false;
      return enqueue_step(Q, op, m, orth_mode, h_dst);
    }
  }
  NSB_CUDA(cudaGraphLaunch(it->second.exec, ctx->stream));
  ctx->launches += it->second.launches;
  op->napply++;
  return NSB_OK;
}

}  // namespace

extern "C" int nsb_arnoldi(nsb_basis_t Q, nsb_op_t op, int mstart, int mend, int orth_mode, double *H,
                           int ldh) {
  NSB_REQUIRE(Q && op && H, "nsb_arnoldi: NULL argument");
  NSB_REQUIRE(mstart >= 0 && mend >= mstart && mend + 1 < Q->ncols,
              "nsb_arnoldi: steps %d..%d need %d columns, basis has %d", mstart, mend, mend + 2, Q->ncols);
  NSB_REQUIRE(ldh >= mend + 2, "nsb_arnoldi: ldh=%d < %d", ldh, mend + 2);
  NSB_REQUIRE(mend + 1 <= kMaxK, "nsb_arnoldi: Krylov dimension above %d", kMaxK);
  NSB_REQUIRE(orth_mode >= 0 && orth_mode <= 2, "nsb_arnoldi: unknown orthogonalisation mode %d", orth_mode);
  nsb_context_t ctx = Q->lay->ctx;
  // device-resident operators (SEM, or compositions of them) allow a factorisation without host syncs: the
  // DGKS decision is taken on the device as well, so all three orthogonalisation modes qualify
  bool dev_op = op->kind == 0;
  if (op->kind == 2) dev_op = op->outer->kind == 0 && op->inner->kind == 0;
  const bool async = dev_op;
  const size_t stride = kMaxK + 8;
  if (async && !ctx->hstage) {   // pinned staging of the H columns, one fixed slot per step index
    NSB_CHECK(nsb_host_alloc((void **)&ctx->hstage, (int64_t)(sizeof(double) * stride * stride)));
    ctx->hstage_elems = stride * stride;
  }
  double *hbuf = ctx->hstage;
  // whole steps as CUDA graphs: single rank, or every collective of the step on the peer-memory path
  bool graphs = async && ctx->use_graph && !ctx->prof && orth_mode != NSB_ORTH_MGS2_REF;
  if (graphs && ctx->nranks > 1) {
    const nsb_op_t parts[2] = {op->kind == 2 ? op->outer : op, op->kind == 2 ? op->inner : op};
    for (nsb_op_t part : parts) graphs = graphs && ctx->p2p && part->sem && part->sem->p2p_halo;
  }
  int rc = NSB_OK;
  for (int m = mstart; m <= mend && rc == NSB_OK; ++m) {
    if (op->kind == 1 && orth_mode == NSB_ORTH_CGS2 && ctx->pipeline_upload && !Q->lay->c0_sem) {
      // host operator: q_m goes to the host, the host matvec runs, f comes back in row chunks with the first
      // projection running on every chunk as it lands.  For a LINEAR operator (nsb_op_set_linear) the download
      // of q_m+1 already started during the third sweep of this step: the host then sees the un-normalised
      // vector beta q and 1/beta is applied to the returned chunks on the device.
      nsb_layout_t L = Q->lay;
      NSB_REQUIRE(L == op->lay, "nsb_arnoldi: host operator built for another layout");
      std::vector<const double *> pin(L->nfields);
      std::vector<double *> pout(L->nfields), pdl(L->nfields);
      for (int f = 0; f < L->nfields; ++f) {
        pin[f] = op->hin[f];
        pdl[f] = op->hin[f];
        pout[f] = op->hout[f];
      }
      double tin = 0.0, tout = 0.0;
      const bool streamed = op->linear && m > mstart;          // hin is being filled by the previous step
      if (streamed) {
        cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
        if (e != cudaSuccess) {
          set_error("nsb_arnoldi: %s", cudaGetErrorString(e));
          rc = NSB_ECUDA;
          break;
        }
        tin = ctx->hpin[3 * (kMaxK + 8) + 9];
      } else {
        rc = nsb_vec_download(Q, m, pdl.data(), &tin);
        if (rc != NSB_OK) break;
      }
      op->napply++;
      if (op->fn(op->user, pin.data(), tin, pout.data(), &tout) != 0) {
        set_error("nsb_arnoldi: host matvec callback failed at step %d", m);
        rc = NSB_EINVAL;
        break;
      }
      std::vector<const double *> pup(pout.begin(), pout.end());
      rc = upload_multidot_pipelined(Q, m + 1, pup.data(), tout, m + 1,
                                     streamed ? ctx->hvec_d + 3 * (kMaxK + 8) + 3 : nullptr);
      if (rc == NSB_OK) {
        if (op->linear && m < mend) {
          StreamOut so{pdl.data(), ctx->hpin + 3 * (kMaxK + 8) + 9};
          rc = orthonormalize_stream_out(Q, m + 1, m + 1, orth_mode, H + (size_t)m * ldh, &so);
        } else {
          rc = nsb_orthonormalize(Q, m + 1, m + 1, orth_mode, H + (size_t)m * ldh, nullptr);
        }
      }
    } else if (async) {
      rc = (graphs && ctx->use_graph) ? step_graph(Q, op, m, orth_mode, hbuf + stride * m)
                                      : enqueue_step(Q, op, m, orth_mode, hbuf + stride * m);
      continue;
    } else {
      rc = nsb_op_apply(op, Q, m, Q, m + 1);
      if (rc == NSB_OK) rc = nsb_orthonormalize(Q, m + 1, m + 1, orth_mode, H + (size_t)m * ldh, nullptr);
    }
    if (rc == NSB_OK && !(H[(size_t)m * ldh + m + 1] > 0.0)) {
      set_error("nsb_arnoldi: breakdown at step %d (residual norm %g)", m, H[(size_t)m * ldh + m + 1]);
      rc = NSB_EBREAKDOWN;
    }
  }
  if (async) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == NSB_OK && e != cudaSuccess) {
      set_error("nsb_arnoldi: %s", cudaGetErrorString(e));
      rc = NSB_ECUDA;
    }
    if (rc == NSB_OK) rc = check_dev_err(ctx);
    if (rc == NSB_OK)
      for (int m = mstart; m <= mend; ++m) {
        const double *h = hbuf + stride * m;
        memcpy(H + (size_t)m * ldh, h, sizeof(double) * (m + 2));
        for (int i = 0; i < m + 2; ++i)
          if (std::isnan(h[i])) {
            set_error("NaN detected in dot product (Arnoldi step %d)", m);
            rc = NSB_ENAN;
          }
        if (rc == NSB_OK && !(h[m + 1] > 0.0)) {
          set_error("nsb_arnoldi: breakdown at step %d (residual norm %g)", m, h[m + 1]);
          rc = NSB_EBREAKDOWN;
        }
      }
  }
  return rc;
}

// Projection passes the DGKS orthogonalisation took in the steps mstart..mend of the LAST device-resident
// nsb_arnoldi call on this basis' context (1 or 2 per step; 2 for every step in the other modes).
extern "C" int nsb_arnoldi_passes(nsb_basis_t Q, int mstart, int mend, int orth_mode, int *passes) {
  NSB_REQUIRE(Q && passes && mstart >= 0 && mend >= mstart && mend + 1 <= kMaxK, "nsb_arnoldi_passes: bad argument");
  nsb_context_t ctx = Q->lay->ctx;
  for (int m = mstart; m <= mend; ++m)
    passes[m - mstart] = (orth_mode == NSB_ORTH_DGKS && ctx->hstage) ? (int)ctx->hstage[(kMaxK + 8) * (size_t)m + m + 2] : 2;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// BM1-weighted QR of the first k columns, in place: qr_dec of BoostConv (core/fixedp.f90:331-385),
// the same weighted Gram-Schmidt kernels as the Arnoldi orthogonalisation.  X = Q R, R upper
// triangular (k x k, ldr).  A column whose residual norm^2 is below 1e-60 is zeroed and gets
// R(j,j) = 1, as in the reference (:371-374).
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_basis_qr(nsb_basis_t B, int k, int orth_mode, double *R, int ldr) {
  NSB_REQUIRE(B && R && k >= 1 && k <= B->ncols && ldr >= k, "nsb_basis_qr: bad argument");
  std::vector<double> h(k + 1);
  for (int j = 0; j < k; ++j) {
    for (int i = 0; i < k; ++i) R[(size_t)j * ldr + i] = 0.0;
    int rc = nsb_orthonormalize(B, j, j, orth_mode, h.data(), nullptr);
    if (rc == NSB_ENAN && !(h[j] * h[j] >= 1e-60)) rc = NSB_OK;   // 0/0 from a vanishing column
    NSB_CHECK(rc);
    for (int i = 0; i < j; ++i) R[(size_t)j * ldr + i] = h[i];
    if (!(h[j] * h[j] >= 1e-60)) {
      NSB_CHECK(nsb_vec_zero(B, j));
      R[(size_t)j * ldr + j] = 1.0;
    } else {
      R[(size_t)j * ldr + j] = h[j];
    }
  }
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Krylov-Schur  (core/eigensolvers.f90:120-359, 363-468)
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_schur_condensation(nsb_basis_t Q, int *mstart, double *H, int ldh, int ksize,
                                      double schur_del, int schur_tgt) {
  NSB_REQUIRE(Q && mstart && H, "nsb_schur_condensation: NULL argument");
  NSB_REQUIRE(ksize >= 1 && ksize + 1 <= Q->ncols && ldh >= ksize + 1, "nsb_schur_condensation: bad sizes");
  const int k = ksize;
  std::vector<double> b_vec(k, 0.0), vecs((size_t)k * k, 0.0), vals(2 * (size_t)k);
  b_vec[k - 1] = H[(size_t)(k - 1) * ldh + k];                                  // :403
  NSB_CHECK(nsb_schur(H, ldh, k, vecs.data(), vals.data()));                    // :407
  std::vector<int> selected(k);
  int m = 0;
  NSB_CHECK(nsb_select_eigenvalues(selected.data(), &m, vals.data(), schur_del, schur_tgt, k));  // :410
  NSB_CHECK(nsb_ordschur(H, ldh, vecs.data(), k, selected.data(), k));          // :414
  for (int j = m; j < k; ++j)                                                   // :417
    for (int i = 0; i < m; ++i) H[(size_t)j * ldh + i] = 0.0;
  for (int j = 0; j < k; ++j)                                                   // :418
    for (int i = m; i < k + 1; ++i) H[(size_t)j * ldh + i] = 0.0;
  NSB_CHECK(nsb_basis_rotate(Q, k, vecs.data(), k, 0));                         // :421-442
  for (int j = 0; j < k; ++j) {                                                 // :446-447
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += b_vec[i] * vecs[(size_t)j * k + i];
    H[(size_t)j * ldh + m] = s;
  }
  NSB_CHECK(nsb_vec_copy(Q, m, Q, k));                                          // :450-453
  *mstart = m;
  return NSB_OK;
}

extern "C" int nsb_krylov_schur(nsb_basis_t Q, nsb_op_t op, int k_dim, int schur_tgt, double eigen_tol,
                                double schur_del, int orth_mode, int max_restarts, double *H, int ldh,
                                double *vals_c16, double *vecs_c16, double *residual, int *cnt_out,
                                int *schur_cnt_out) {
  NSB_REQUIRE(Q && op && H && vals_c16 && vecs_c16 && residual, "nsb_krylov_schur: NULL argument");
  NSB_REQUIRE(k_dim >= 1 && k_dim + 1 <= Q->ncols && ldh >= k_dim + 1, "nsb_krylov_schur: bad sizes");
  const int k = k_dim;
  for (int j = 0; j < k; ++j) memset(H + (size_t)j * ldh, 0, sizeof(double) * (k + 1));
  int mstart = 0, schur_cnt = 0, cnt = 0;
  for (;;) {
    NSB_CHECK(nsb_arnoldi(Q, op, mstart, k - 1, orth_mode, H, ldh));            // :297
    NSB_CHECK(nsb_eig(H, ldh, k, vecs_c16, vals_c16));                          // :306
    const double hk = H[(size_t)(k - 1) * ldh + k];
    cnt = 0;
    for (int j = 0; j < k; ++j) {                                               // :309-310
      const double re = vecs_c16[2 * ((size_t)j * k + (k - 1))], im = vecs_c16[2 * ((size_t)j * k + (k - 1)) + 1];
      residual[j] = std::fabs(hk) * std::hypot(re, im);
      if (residual[j] < eigen_tol) ++cnt;
    }
    if (schur_tgt <= 0 || cnt >= schur_tgt || schur_cnt >= max_restarts) break;  // :314-331
    ++schur_cnt;
    NSB_CHECK(nsb_schur_condensation(Q, &mstart, H, ldh, k, schur_del, schur_tgt));
  }
  if (cnt_out) *cnt_out = cnt;
  if (schur_cnt_out) *schur_cnt_out = schur_cnt;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Step-wise eigensolver of the LightKrylov path: linear_stability_analysis calls
// eigs(A, X, eigvecs, eigvals, residuals, info, nev=schur_tgt, tolerance=eigen_tol)
// (core/linear_stab.f90:66).  [UPSTREAM-RECALL, LightKrylov is not vendored]: one Arnoldi step at a
// time, eig(H(1:k,1:k)) and residuals |H(k+1,k) y_k| after every step, stop once nev Ritz pairs
// are below the tolerance.  Returns the Krylov dimension reached in *kused.
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_eigs(nsb_basis_t Q, nsb_op_t op, int k_dim, int nev, double tol, int orth_mode, double *H,
                        int ldh, double *vals_c16, double *vecs_c16, double *residual, int *kused, int *nconv) {
  NSB_REQUIRE(Q && op && H && vals_c16 && vecs_c16 && residual && kused, "nsb_eigs: NULL argument");
  NSB_REQUIRE(k_dim >= 1 && k_dim + 1 <= Q->ncols && ldh >= k_dim + 1 && nev >= 1, "nsb_eigs: bad sizes");
  for (int j = 0; j < k_dim; ++j) memset(H + (size_t)j * ldh, 0, sizeof(double) * (k_dim + 1));
  int k = 0, cnt = 0;
  std::vector<double> vecs, vals;
  for (k = 1; k <= k_dim; ++k) {
    NSB_CHECK(nsb_arnoldi(Q, op, k - 1, k - 1, orth_mode, H, ldh));
    vecs.assign((size_t)2 * k * k, 0.0);
    vals.assign((size_t)2 * k, 0.0);
    NSB_CHECK(nsb_eig(H, ldh, k, vecs.data(), vals.data()));
    const double hk = std::fabs(H[(size_t)(k - 1) * ldh + k]);
    cnt = 0;
    for (int j = 0; j < k; ++j) {
      residual[j] = hk * std::hypot(vecs[2 * ((size_t)j * k + (k - 1))], vecs[2 * ((size_t)j * k + (k - 1)) + 1]);
      if (residual[j] < tol) ++cnt;
    }
    if (cnt >= nev) break;
  }
  if (k > k_dim) k = k_dim;
  memcpy(vals_c16, vals.data(), sizeof(double) * 2 * k);
  for (int j = 0; j < k; ++j)   // k x k eigenvector block into the caller's k_dim-leading-dimension array
    memcpy(vecs_c16 + 2 * (size_t)j * k_dim, vecs.data() + 2 * (size_t)j * k, sizeof(double) * 2 * k);
  *kused = k;
  if (nconv) *nconv = cnt;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Ritz-vector assembly (core/eigensolvers.f90:565-585, 609-615; get_vec core/linear_stab.f90:362):
// fp = Q(:,1:k) y with complex y -> real part = Q Re(y), imaginary part = Q Im(y) (two device GEMVs over
// the basis), alpha_r = |Re fp|, alpha_i = |Im fp| in the BM1 norm, and the reference's scaling of both
// parts by beta = 1 / sqrt(alpha_r^2 + alpha_i^2) so that the volume integral of fp conj(fp) is 1.
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_ritz_vector(nsb_basis_t Q, int k, const double *y_c16, nsb_basis_t bout, int cre, int cim,
                               int normalize, double *alpha_re, double *alpha_im) {
  NSB_REQUIRE(Q && y_c16 && bout, "nsb_ritz_vector: NULL argument");
  NSB_REQUIRE(k >= 1 && k <= Q->ncols, "nsb_ritz_vector: k=%d out of range", k);
  NSB_REQUIRE(cre != cim, "nsb_ritz_vector: real and imaginary part need different columns");
  std::vector<double> yre(k), yim(k);
  for (int i = 0; i < k; ++i) {
    yre[i] = y_c16[2 * i];
    yim[i] = y_c16[2 * i + 1];
  }
  NSB_CHECK(nsb_basis_gemv(Q, k, yre.data(), bout, cre));
  NSB_CHECK(nsb_basis_gemv(Q, k, yim.data(), bout, cim));
  double ar = 0.0, ai = 0.0;
  NSB_CHECK(nsb_vec_norm(bout, cre, &ar));
  NSB_CHECK(nsb_vec_norm(bout, cim, &ai));
  if (normalize) {
    const double a2 = ar * ar + ai * ai;
    if (!(a2 > 0.0)) {
      set_error("nsb_ritz_vector: zero vector");
      return NSB_EBREAKDOWN;
    }
    const double beta = 1.0 / std::sqrt(a2);
    NSB_CHECK(nsb_vec_scal(bout, cre, beta));
    NSB_CHECK(nsb_vec_scal(bout, cim, beta));
  }
  if (alpha_re) *alpha_re = ar;
  if (alpha_im) *alpha_im = ai;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// svds: the call transient_growth_analysis makes (core/linear_stab.f90:112:
// svds(A, U, V, uvecs, vvecs, sigma, residuals, info, nev, tolerance)).  [UPSTREAM-RECALL:
// LightKrylov's Golub-Kahan Lanczos bidiagonalisation with full re-orthogonalisation, one step at a
// time:  v_k = A^T u_k  (orthogonalised against V(1:k-1)), alpha = |v_k| = B(k,k);
//        u_k+1 = A v_k  (orthogonalised against U(1:k)),   beta  = |u_k+1| = B(k+1,k);
// then svd(B(1:k,1:k)), residual_i = |beta * vvecs(k,i)|, stop once nev triplets are below tol.]
// The re-orthogonalisation is the same fused BM1-weighted kernel sequence as the Arnoldi step.
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_svds(nsb_basis_t U, nsb_basis_t V, nsb_op_t op, nsb_op_t op_adj, int k_dim, int nev, double tol,
                        int orth_mode, double *B, int ldb, double *sigma, double *uvecs, double *vvecs,
                        double *residual, int *kused, int *nconv) {
  NSB_REQUIRE(U && V && op && op_adj && B && sigma && uvecs && vvecs && residual && kused, "nsb_svds: NULL argument");
  NSB_REQUIRE(k_dim >= 1 && k_dim + 1 <= U->ncols && k_dim <= V->ncols && ldb >= k_dim + 1 && nev >= 1,
              "nsb_svds: bad sizes (U needs k_dim+1 columns, V k_dim)");
  NSB_REQUIRE(k_dim <= kMaxK, "nsb_svds: k_dim=%d exceeds %d", k_dim, kMaxK);
  for (int j = 0; j < k_dim; ++j) memset(B + (size_t)j * ldb, 0, sizeof(double) * (k_dim + 1));
  std::vector<double> h(k_dim + 2), bu, bv, bs;
  int k = 0, kdone = 0, cnt = 0;
  for (k = 1; k <= k_dim; ++k) {
    const int c = k - 1;
    NSB_CHECK(nsb_op_apply(op_adj, U, c, V, c));
    NSB_CHECK(nsb_orthonormalize(V, c, c, orth_mode, h.data(), nullptr));
    const double alpha = h[c];
    B[(size_t)c * ldb + c] = alpha;
    if (!(alpha > tol)) break;                                   // invariant subspace
    NSB_CHECK(nsb_op_apply(op, V, c, U, k));
    NSB_CHECK(nsb_orthonormalize(U, k, k, orth_mode, h.data(), nullptr));
    const double beta = h[k];
    B[(size_t)c * ldb + k] = beta;
    kdone = k;
    bu.assign((size_t)k * k, 0.0);
    bv.assign((size_t)k * k, 0.0);
    bs.assign(k, 0.0);
    NSB_CHECK(nsb_svd(B, ldb, k, k, bu.data(), bs.data(), bv.data()));
    cnt = 0;
    for (int j = 0; j < k; ++j) {
      residual[j] = std::fabs(beta * bv[(size_t)j * k + (k - 1)]);
      if (residual[j] < tol) ++cnt;
    }
    if (cnt >= nev || !(beta > tol)) break;
  }
  k = kdone;
  for (int j = 0; j < k; ++j) {
    sigma[j] = bs[j];
    memcpy(uvecs + (size_t)j * k_dim, bu.data() + (size_t)j * k, sizeof(double) * k);
    memcpy(vvecs + (size_t)j * k_dim, bv.data() + (size_t)j * k, sizeof(double) * k);
  }
  *kused = k;
  if (nconv) *nconv = cnt;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// GMRES  (core/newton_krylov.f90:170-326)
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_ts_gmres(nsb_basis_t Q, nsb_op_t op, nsb_basis_t brhs, int crhs, nsb_basis_t bsol,
                            int csol, int maxiter, int ksize, double tol, int orth_mode, int *calls_out,
                            double *residual_hist, int *nhist) {
  NSB_REQUIRE(Q && op && brhs && bsol, "nsb_ts_gmres: NULL argument");
  NSB_REQUIRE(ksize >= 1 && Q->ncols >= ksize + 2, "nsb_ts_gmres: basis needs ksize+2 columns");
  NSB_REQUIRE(maxiter >= 1, "nsb_ts_gmres: maxiter=%d", maxiter);
  const int k1 = ksize + 1, wrk = ksize + 1, ldh = ksize + 1;
  std::vector<double> H((size_t)ldh * ksize), yvec(ksize), evec(k1);
  double beta = 0.0;
  int calls = 0, nh = 0;
  NSB_CHECK(nsb_vec_zero(bsol, csol));                                          // :236
  NSB_CHECK(nsb_vec_copy(Q, 0, brhs, crhs));                                    // :241
  NSB_CHECK(nsb_vec_normalize(Q, 0, &beta));                                    // :242
  for (int it = 0; it < maxiter; ++it) {                                        // :245
    std::fill(H.begin(), H.end(), 0.0);
    std::fill(yvec.begin(), yvec.end(), 0.0);
    std::fill(evec.begin(), evec.end(), 0.0);
    evec[0] = beta;
    int kk = ksize;
    for (int k = 1; k <= ksize; ++k) {                                          // :250
      NSB_CHECK(nsb_arnoldi(Q, op, k - 1, k - 1, orth_mode, H.data(), ldh));    // :252
      ++calls;
      NSB_CHECK(nsb_lstsq(H.data(), ldh, k + 1, k, evec.data(), yvec.data()));  // :255
      double r2 = 0.0;                                                          // :258
      for (int i = 0; i < k + 1; ++i) {
        double s = evec[i];
        for (int j = 0; j < k; ++j) s -= H[(size_t)j * ldh + i] * yvec[j];
        r2 += s * s;
      }
      beta = std::sqrt(r2);
      if (beta * beta < tol) {                                                  // :266
        kk = k;
        break;
      }
    }
    NSB_CHECK(nsb_basis_gemv(Q, kk, yvec.data(), Q, wrk));                      // :279 (min(k, ksize) columns)
    NSB_CHECK(nsb_vec_add2(bsol, csol, Q, wrk));                                // :280
    // initialize_gmres_vector :303-326
    NSB_CHECK(nsb_vec_copy(Q, 0, bsol, csol));
    NSB_CHECK(nsb_op_apply(op, Q, 0, Q, wrk));
    ++calls;
    NSB_CHECK(nsb_vec_sub2(Q, wrk, brhs, crhs));
    NSB_CHECK(nsb_vec_scal(Q, wrk, -1.0));
    NSB_CHECK(nsb_vec_normalize(Q, wrk, &beta));
    NSB_CHECK(nsb_vec_copy(Q, 0, Q, wrk));
    if (residual_hist) residual_hist[nh] = beta * beta;
    ++nh;
    if (beta * beta < tol) break;                                               // :293
  }
  if (calls_out) *calls_out = calls;
  if (nhist) *nhist = nh;
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Newton-Krylov fixed-point iteration  (core/newton_krylov.f90:1-168): f = F(q) (:102), residual = |f|^2
// (:107), stop when residual < tol (:117), dq = ts_gmres(J, f) (:125), q -= dq (:130).  F and J are operator
// handles: in the reference both are the host time-stepper (nonlinear_forward_map and the linearised
// solver set up about the current q by prepare_linearized_solver, :71), i.e. host callbacks whose owner
// re-linearises inside F; all vectors, the GMRES basis and its orthogonalisation stay on the device.
// (bw, cf) and (bw, cdq) are two work vectors.  residual_hist holds maxiter_newton entries.
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_newton_krylov(nsb_basis_t Q, nsb_op_t fop, nsb_op_t jop, nsb_basis_t bq, int cq, nsb_basis_t bw,
                                 int cf, int cdq, int maxiter_newton, int maxiter_gmres, int ksize, double tol,
                                 int orth_mode, int *iters, double *residual_hist, int *calls) {
  NSB_REQUIRE(Q && fop && jop && bq && bw, "nsb_newton_krylov: NULL argument");
  NSB_REQUIRE(maxiter_newton >= 1 && maxiter_gmres >= 1, "nsb_newton_krylov: bad iteration limits");
  NSB_REQUIRE(cf != cdq && !(bq == bw && (cq == cf || cq == cdq)), "nsb_newton_krylov: q, f and dq must be distinct");
  int ncalls = 0, it = 0;
  NSB_CHECK(nsb_vec_zero(bw, cf));                                              // :53
  NSB_CHECK(nsb_vec_zero(bw, cdq));
  for (it = 1; it <= maxiter_newton; ++it) {                                    // :59
    NSB_CHECK(nsb_op_apply(fop, bq, cq, bw, cf));                               // :102
    double residual = 0.0;
    NSB_CHECK(nsb_vec_dot(bw, cf, bw, cf, &residual));                          // :107
    if (residual_hist) residual_hist[it - 1] = residual;
    if (residual < tol) break;                                                  // :117
    int c = 0;
    NSB_CHECK(nsb_ts_gmres(Q, jop, bw, cf, bw, cdq, maxiter_gmres, ksize, tol, orth_mode, &c, nullptr, nullptr));  // :125
    ncalls += c;
    NSB_CHECK(nsb_vec_sub2(bq, cq, bw, cdq));                                   // :130
  }
  if (iters) *iters = it > maxiter_newton ? maxiter_newton : it;
  if (calls) *calls = ncalls;
  return NSB_OK;
}
