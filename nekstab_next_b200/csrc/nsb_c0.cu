// C0 (unique-node) storage of the Krylov basis.
//
// Nek's element-local layout (core/nek_vectors.f90:20-31: vx(lv), lv = lx1*ly1*lz1*lelv) stores a node shared by
// several elements once per element: at N = 7 on the 32^3 box 16.78 M local points stand for 11.39 M distinct
// nodes.  Every Krylov vector of the path is continuous (it leaves nek_advance / the operator through a dssum), so
// the copies carry no information, and the reference's inner product
//     sum over local points  a bm1s b        (core/krylov_subspace.f90:40-49, glsc3 on the unassembled bm1s)
// equals   sum over distinct nodes  a (QQ^T bm1s) b   with the ASSEMBLED weight.  A layout created with
// nsb_layout_create_c0 keeps the first n_c0 fields on the distinct nodes of a mesh:
//     [ element-interior nodes, element by element | element-boundary nodes in gather-scatter order ]
// -- 32 % fewer rows at N = 7, hence 32 % fewer bytes in every sweep of the orthogonalisation, in the basis
// rotation and in the BLAS-1 set, and no write-back to the copies in the operator's gather-scatter.  The host
// interface is unchanged: nsb_vec_upload / nsb_vec_download exchange element-local arrays (upload takes the first
// copy of a node: the caller's field must be continuous -- that is the contract of this layout; use the
// element-local layout otherwise).  Multi-rank: a node on a rank interface is stored by every rank touching it,
// each with its share of the weight, so the all-reduced inner product counts it once.
#include <algorithm>
#include <cstring>
#include <vector>

#include "nsb_internal.h"
#include "nsb_device.cuh"

using namespace nsb;

namespace {

inline unsigned nblk(int64_t n, int nt = 256) { return (unsigned)((n + nt - 1) / nt); }

// interior values: unique index q = e * NI + ((k-1) m + (j-1)) m + (i-1), m = lx - 2  <->  local point p
// DIR 0: loc <- uni ; DIR 1: uni <- loc
template <int DIR>
__global__ void __launch_bounds__(256)
c0_interior_kernel(double *__restrict__ uni, double *__restrict__ loc, int64_t nint, int NI, int m, int lx, int nloc,
                   int dim, int nf, int64_t fs_uni, int64_t fs_loc) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nint) return;
  const int64_t e = q / NI;
  const int r = (int)(q - e * NI);
  const int i = r % m, j = (r / m) % m, k = dim == 3 ? r / (m * m) : 0;
  const int64_t p = e * nloc + (dim == 3 ? ((int64_t)(k + 1) * lx + (j + 1)) * lx + (i + 1) : (int64_t)(j + 1) * lx + (i + 1));
  for (int f = 0; f < nf; ++f) {
    if (DIR == 0) loc[(int64_t)f * fs_loc + p] = uni[(int64_t)f * fs_uni + q];
    else uni[(int64_t)f * fs_uni + q] = loc[(int64_t)f * fs_loc + p];
  }
}

// boundary nodes n in [n0, n1): MODE 0 loc copies <- uni ; MODE 1 uni <- first copy ; MODE 2 uni <- sum of copies ;
// MODE 3 uni_out <- alpha uni_in + beta bnode * sum of copies (the operator's assembly)
template <int MODE>
__global__ void __launch_bounds__(256)
c0_boundary_kernel(double *__restrict__ uni, double *__restrict__ loc, const int64_t *__restrict__ off,
                   const int32_t *__restrict__ idx, int64_t n0, int64_t n1, int nf, int64_t fs_uni, int64_t fs_loc,
                   const double *__restrict__ uni_in, double alpha, double beta, const double *__restrict__ bnode) {
  const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n1) return;
  const int64_t a = off[n], b = off[n + 1];
  for (int f = 0; f < nf; ++f) {
    double *lf = loc + (int64_t)f * fs_loc;
    if (MODE == 0) {
      const double v = uni[(int64_t)f * fs_uni + n];
      for (int64_t q = a; q < b; ++q) lf[idx[q]] = v;
    } else if (MODE == 1) {
      uni[(int64_t)f * fs_uni + n] = lf[idx[a]];
    } else {
      double s = 0.0;
      for (int64_t q = a; q < b; ++q) s += lf[idx[q]];
      if (MODE == 2) uni[(int64_t)f * fs_uni + n] = s;
      else uni[(int64_t)f * fs_uni + n] = alpha * uni_in[(int64_t)f * fs_uni + n] + beta * bnode[n] * s;
    }
  }
}

// loc[p] = uni[l2u[p]]: one thread per local point, coalesced writes of whole rows, the index read once for all
// fields (the scatter form -- one thread per node writing its copies -- left every 32-byte sector of the
// element-local scratch to be completed by several threads at different times)
__global__ void __launch_bounds__(256)
c0_gather_kernel(const double *__restrict__ uni, double *__restrict__ loc, const int32_t *__restrict__ l2u, int64_t npts,
                 int nf, int64_t fs_uni, int64_t fs_loc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    const int32_t q = __ldg(l2u + p);
    for (int f = 0; f < nf; ++f) loc[(int64_t)f * fs_loc + p] = __ldg(uni + (int64_t)f * fs_uni + q);
  }
}

// local point -> unique row (interior rows by arithmetic, boundary rows from the gather-scatter lists)
int ensure_l2u(nsb_sem_t S) {
  if (S->c0_l2u_d && S->c0_l2u_nshared == S->nshared && S->c0_l2u_nlocal == S->n_local) return NSB_OK;
  cudaSetDevice(S->ctx->device);
  const int64_t nint = c0_nint(S), nloc = S->npts / S->nel;
  const int lx = S->lx, m = lx - 2, NI = (int)(S->nel ? nint / S->nel : 0);
  std::vector<int32_t> l2u((size_t)S->npts, -1);
  for (int64_t e = 0; e < S->nel; ++e)
    for (int r = 0; r < NI; ++r) {
      const int i = r % m, j = (r / m) % m, k = S->dim == 3 ? r / (m * m) : 0;
      const int64_t p = e * nloc + (S->dim == 3 ? ((int64_t)(k + 1) * lx + (j + 1)) * lx + (i + 1) : (int64_t)(j + 1) * lx + (i + 1));
      l2u[p] = (int32_t)(e * NI + r);
    }
  for (int64_t n = 0; n < S->nshared; ++n)
    for (int64_t q = S->gs_off_h[n]; q < S->gs_off_h[n + 1]; ++q) l2u[S->gs_idx_h[q]] = (int32_t)(nint + n);
  if (S->c0_l2u_d) cudaFree(S->c0_l2u_d);
  S->c0_l2u_d = nullptr;
  NSB_CUDA(cudaMalloc(&S->c0_l2u_d, sizeof(int32_t) * S->npts));
  NSB_CUDA(cudaMemcpy(S->c0_l2u_d, l2u.data(), sizeof(int32_t) * S->npts, cudaMemcpyHostToDevice));
  S->c0_l2u_nshared = S->nshared;
  S->c0_l2u_nlocal = S->n_local;
  return NSB_OK;
}

int ensure_scratch(nsb_sem_t S, int nf) {
  const size_t need = (size_t)2 * nf * S->npts;
  if (S->c0_scratch_elems >= need) return NSB_OK;
  cudaSetDevice(S->ctx->device);
  NSB_CUDA(cudaStreamSynchronize(S->ctx->stream));
  clear_step_graphs(S->ctx);
  if (S->c0_scratch_d) cudaFree(S->c0_scratch_d);
  S->c0_scratch_d = nullptr;
  NSB_CUDA(cudaMalloc(&S->c0_scratch_d, sizeof(double) * need));
  S->c0_scratch_elems = need;
  return NSB_OK;
}

}  // namespace

namespace nsb {

int64_t c0_nint(nsb_sem_t S) {
  int64_t ni = 1;
  for (int a = 0; a < S->dim; ++a) ni *= (S->lx - 2 > 0 ? S->lx - 2 : 0);
  return S->nel * ni;
}

// element-local field(s) on the device -> unique-node field(s)   (first copy of every boundary node)
int c0_compact(nsb_sem_t S, const double *loc, int64_t fs_loc, double *uni, int64_t fs_uni, int nf, int mode_sum) {
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  const int64_t nint = c0_nint(S);
  const int m = S->lx - 2, NI = (int)(nint / S->nel), nloc = (int)(S->npts / S->nel);
  if (nint > 0)
    c0_interior_kernel<1><<<nblk(nint), 256, 0, ctx->stream>>>(uni, const_cast<double *>(loc), nint, NI, m, S->lx, nloc, S->dim,
                                                              nf, fs_uni, fs_loc);
  if (S->nshared > 0) {
    if (mode_sum)
      c0_boundary_kernel<2><<<nblk(S->nshared), 256, 0, ctx->stream>>>(uni + nint, const_cast<double *>(loc), S->gs_off_d,
                                                                      S->gs_idx_d, 0, S->nshared, nf, fs_uni, fs_loc, nullptr,
                                                                      0, 0, nullptr);
    else
      c0_boundary_kernel<1><<<nblk(S->nshared), 256, 0, ctx->stream>>>(uni + nint, const_cast<double *>(loc), S->gs_off_d,
                                                                      S->gs_idx_d, 0, S->nshared, nf, fs_uni, fs_loc, nullptr,
                                                                      0, 0, nullptr);
  }
  ctx->launches += 2;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// unique-node field(s) -> element-local field(s) on the device (every copy of a node gets its value)
int c0_expand(nsb_sem_t S, const double *uni, int64_t fs_uni, double *loc, int64_t fs_loc, int nf) {
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  const int64_t nint = c0_nint(S);
  const int m = S->lx - 2, NI = (int)(nint / S->nel), nloc = (int)(S->npts / S->nel);
  ProfScope ps(ctx, PC_GS, 8.0 * nf * (double)(nint + S->nshared + S->npts) + 4.0 * (double)S->npts);
  (void)m; (void)NI; (void)nloc;
  NSB_CHECK(ensure_l2u(S));
  c0_gather_kernel<<<ctx->num_sms * 16, 256, 0, ctx->stream>>>(uni, loc, S->c0_l2u_d, S->npts, nf, fs_uni, fs_loc);
  ctx->launches += 1;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// host element-local arrays -> fields [0, n_c0) of column col
int c0_upload(nsb_basis_t B, int col, const double *const *fields) {
  nsb_layout_t L = B->lay;
  nsb_sem_t S = L->c0_sem;
  NSB_CHECK(ensure_scratch(S, 1));
  cudaStream_t st = L->ctx->stream;
  for (int f = 0; f < L->c0_nfields; ++f) {
    double *dst = B->col(col) + L->off[f];
    if (!fields[f]) {
      NSB_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * L->len[f], st));
      continue;
    }
    NSB_CUDA(cudaMemcpyAsync(S->c0_scratch_d, fields[f], sizeof(double) * S->npts, cudaMemcpyHostToDevice, st));
    NSB_CHECK(c0_compact(S, S->c0_scratch_d, 0, dst, 0, 1, 0));
  }
  return NSB_OK;
}

int c0_download(nsb_basis_t B, int col, double *const *fields) {
  nsb_layout_t L = B->lay;
  nsb_sem_t S = L->c0_sem;
  NSB_CHECK(ensure_scratch(S, 1));
  cudaStream_t st = L->ctx->stream;
  for (int f = 0; f < L->c0_nfields; ++f) {
    if (!fields[f]) continue;
    NSB_CHECK(c0_expand(S, B->col(col) + L->off[f], 0, S->c0_scratch_d, 0, 1));
    NSB_CUDA(cudaMemcpyAsync(fields[f], S->c0_scratch_d, sizeof(double) * S->npts, cudaMemcpyDeviceToHost, st));
    NSB_CUDA(cudaStreamSynchronize(st));   // the scratch is reused by the next field
  }
  return NSB_OK;
}

// weight of a C0 field: the assembled QQ^T bm1s (this rank's copies only -- the all-reduce adds the other ranks')
int c0_set_weight(nsb_layout_t L, int f, const double *w_host) {
  nsb_sem_t S = L->c0_sem;
  NSB_CHECK(ensure_scratch(S, 1));
  cudaStream_t st = L->ctx->stream;
  NSB_CUDA(cudaMemcpyAsync(S->c0_scratch_d, w_host, sizeof(double) * S->npts, cudaMemcpyHostToDevice, st));
  NSB_CHECK(c0_compact(S, S->c0_scratch_d, 0, L->w_d + L->off[f], 0, 1, 1));
  return NSB_OK;
}

// The fused SEM operator on a C0 layout:  out = alpha in + beta bmask QQ^T (h1 A + h2 B [+ C.grad]) in
//   1. expand `in` to the element-local scratch (every element needs its own copy of a shared node)
//   2. axhelm with the interior epilogue (interior points final, boundary points raw)
//   3. assembly: interior values are copied, a boundary node gets alpha in + beta bnode * (sum of its copies);
//      interface nodes go through the peer-memory exchange and land in the unique layout directly.
// Nothing is written back to the copies and `in` is not re-read per copy.
int c0_apply_sem(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  nsb_sem_t S = op->sem;
  nsb_layout_t L = bin->lay;
  nsb_context_t ctx = L->ctx;
  const int nf = op->nfields_apply;
  NSB_REQUIRE(nf <= L->c0_nfields, "nsb_op_apply: operator covers %d fields, %d are stored on unique nodes", nf, L->c0_nfields);
  NSB_REQUIRE(nf <= 3 || ctx->nranks == 1, "nsb_op_apply: more than 3 fields on a multi-rank C0 layout");
  NSB_CHECK(ensure_scratch(S, nf));
  const int64_t fs_u = nf > 1 ? L->off[1] - L->off[0] : 0;
  for (int f = 1; f < nf; ++f) NSB_REQUIRE(L->off[f] - L->off[f - 1] == fs_u, "nsb_op_apply: C0 fields are not equally spaced");
  const int64_t nint = c0_nint(S), npts = S->npts;
  double *uloc = S->c0_scratch_d, *wloc = S->c0_scratch_d + (size_t)nf * npts;
  const double *uin = bin->col(cin) + L->off[0];
  double *wout = bout->col(cout) + L->off[0];
  cudaSetDevice(ctx->device);
  NSB_CHECK(c0_expand(S, uin, fs_u, uloc, npts, nf));
  NSB_CHECK(launch_axhelm_ext(S, uloc, wloc, nf, npts, op->h1, op->h2, op->c_d, 1, op->alpha, op->beta, S->bmask_d));
  const int m = S->lx - 2, NI = (int)(nint / S->nel), nloc = (int)(npts / S->nel);
  const bool multi = ctx->nranks > 1 && (S->nshared - S->n_local) > 0 && !S->peers.empty();
  if (multi) {
    NSB_REQUIRE(S->p2p_halo, "nsb_op_apply: the C0 layout needs the peer-memory halo exchange on multi-rank runs");
    NSB_CUDA(cudaEventRecord(S->ev_a, ctx->stream));
    NSB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, S->ev_a, 0));
    NSB_CHECK(halo_exchange_fused_c0(S, wloc, nf, npts, wout + nint, uin + nint, fs_u, op->alpha, op->beta, ctx->copy_stream));
    NSB_CUDA(cudaEventRecord(S->ev_b, ctx->copy_stream));
  }
  {
    ProfScope ps(ctx, PC_GS, 8.0 * nf * (double)(2 * nint + 2 * S->n_local + S->npts) + 4.0 * (double)S->gs_nnz);
    if (nint > 0)
      c0_interior_kernel<1><<<nblk(nint), 256, 0, ctx->stream>>>(wout, wloc, nint, NI, m, S->lx, nloc, S->dim, nf, fs_u, npts);
    if (S->n_local > 0)
      c0_boundary_kernel<3><<<nblk(S->n_local), 256, 0, ctx->stream>>>(wout + nint, wloc, S->gs_off_d, S->gs_idx_d, 0, S->n_local,
                                                                      nf, fs_u, npts, uin + nint, op->alpha, op->beta, S->bnode_d);
    ctx->launches += 2;
  }
  if (multi) NSB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_b, 0));
  NSB_CUDA(cudaGetLastError());
  // rows outside the operator are carried through: %time and the remaining fields
  cudaStream_t s = ctx->stream;
  NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->time_row, bin->col(cin) + L->time_row, sizeof(double), cudaMemcpyDeviceToDevice, s));
  for (int f = nf; f < L->nfields; ++f)
    if (L->len[f] > 0)
      NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->off[f], bin->col(cin) + L->off[f], sizeof(double) * L->len[f],
                               cudaMemcpyDeviceToDevice, s));
  return NSB_OK;
}

}  // namespace nsb

// field_len[f] are the HOST lengths (element-local points for the first n_c0 fields, which must equal the mesh's
// point count); those fields are stored on the mesh's distinct nodes.
extern "C" int nsb_layout_create_c0(nsb_context_t ctx, nsb_sem_t sem, int nfields, const int64_t *field_len,
                                    const int *field_in_dot, int time_in_dot, int n_c0, nsb_layout_t *layout) {
  NSB_REQUIRE(ctx && sem && field_len && field_in_dot && layout, "nsb_layout_create_c0: NULL argument");
  NSB_REQUIRE(n_c0 >= 1 && n_c0 <= nfields, "nsb_layout_create_c0: n_c0=%d of %d fields", n_c0, nfields);
  NSB_REQUIRE(sem->ctx == ctx, "nsb_layout_create_c0: mesh and context differ");
  NSB_REQUIRE(sem->exchange_ready, "nsb_layout_create_c0: call nsb_sem_setup_exchange first");
  std::vector<int64_t> stored(field_len, field_len + nfields);
  const int64_t nuni = c0_nint(sem) + sem->nshared;
  for (int f = 0; f < n_c0; ++f) {
    NSB_REQUIRE(field_len[f] == sem->npts, "nsb_layout_create_c0: field %d has %lld points, the mesh has %lld", f,
                (long long)field_len[f], (long long)sem->npts);
    stored[f] = nuni;
  }
  NSB_CHECK(ensure_scratch(sem, std::min(n_c0, 3)));   // not inside a captured Arnoldi step later on
  NSB_CHECK(ensure_l2u(sem));
  NSB_CHECK(nsb_layout_create(ctx, nfields, stored.data(), field_in_dot, time_in_dot, layout));
  (*layout)->c0_sem = sem;
  (*layout)->c0_nfields = n_c0;
  (*layout)->hlen.assign(field_len, field_len + nfields);
  // the reference's dot counts every local point: ndof_dot keeps that count for the algorithmic-byte reports of
  // the element-local layout; stored rows are what the kernels move
  return NSB_OK;
}

extern "C" int nsb_layout_is_c0(nsb_layout_t layout, int *n_c0, int64_t *stored_rows_per_field) {
  NSB_REQUIRE(layout, "nsb_layout_is_c0: NULL layout");
  if (n_c0) *n_c0 = layout->c0_nfields;
  if (stored_rows_per_field) *stored_rows_per_field = layout->c0_nfields ? layout->len[0] : 0;
  return NSB_OK;
}
