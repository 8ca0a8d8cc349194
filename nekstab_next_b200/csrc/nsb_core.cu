// Context, layout, basis storage and the BLAS-1 set of the nek_dvector type.
//
// Reference semantics: core/nek_vectors.f90:70-139, 209-362 (real_zero/dot/scal/axpby, nop*),
// core/krylov_subspace.f90:26-161 (k_dot, k_norm, k_normalize, k_cmult, k_add2, k_sub2, k_sub3,
// k_zero, k_copy).  All kernels are HBM-streaming: 128-bit loads, 4 independent loads per
// thread in flight, grid a multiple of the SM count.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

#include <cuda_profiler_api.h>

#include "nsb_internal.h"
#include "nsb_device.cuh"

namespace nsb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_partial(nsb_context_t ctx, int64_t rows) {
  if (rows <= ctx->partial_rows) return NSB_OK;
  clear_step_graphs(ctx);   // captured steps hold the old address
  if (ctx->partial_d) NSB_CUDA(cudaFree(ctx->partial_d));
  ctx->partial_d = nullptr;
  NSB_CUDA(cudaMalloc(&ctx->partial_d, sizeof(double) * rows * (kMaxK + 8)));
  ctx->partial_rows = rows;
  return NSB_OK;
}

int check_dev_err(nsb_context_t ctx) {
  if (!ctx->dev_err) return NSB_OK;
  const int e = *(volatile int *)ctx->dev_err;
  if (e == 0) return NSB_OK;
  *(volatile int *)ctx->dev_err = 0;
  set_error("device-side wait timed out (%s): a peer rank never published its contribution",
            e == 1 ? "peer-memory all-reduce" : "halo exchange");
  return NSB_ECUDA;
}

void clear_step_graphs(nsb_context_t ctx) {
  for (auto &kv : ctx->step_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  ctx->step_graphs.clear();
}

ProfScope::ProfScope(nsb_context_t ctx, int cls, double bytes) : c(ctx) {
  if (!ctx->prof) return;
  cudaEvent_t e[2];
  for (int i = 0; i < 2; ++i) {
    if (!ctx->prof_pool.empty()) {
      e[i] = ctx->prof_pool.back();
      ctx->prof_pool.pop_back();
    } else if (cudaEventCreate(&e[i]) != cudaSuccess) {
      return;
    }
  }
  cudaEventRecord(e[0], ctx->stream);
  ctx->prof_recs.push_back({cls, e[0], e[1], bytes});
  slot = (int)ctx->prof_recs.size() - 1;
}

ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(c->prof_recs[slot].e1, c->stream);
}

}  // namespace nsb

using namespace nsb;

extern "C" int nsb_prof_enable(nsb_context_t ctx, int on) {
  NSB_REQUIRE(ctx, "nsb_prof_enable: NULL context");
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &r : ctx->prof_recs) {
    ctx->prof_pool.push_back(r.e0);
    ctx->prof_pool.push_back(r.e1);
  }
  ctx->prof_recs.clear();
  ctx->prof = on != 0;
  return NSB_OK;
}

extern "C" int nsb_prof_get(nsb_context_t ctx, int cls, double *ms, int64_t *launches, double *bytes) {
  NSB_REQUIRE(ctx && cls >= 0 && cls < PC_COUNT, "nsb_prof_get: bad argument");
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  double t = 0, b = 0;
  int64_t n = 0;
  for (auto &r : ctx->prof_recs) {
    if (r.cls != cls) continue;
    float f = 0.f;
    NSB_CUDA(cudaEventElapsedTime(&f, r.e0, r.e1));
    t += f;
    b += r.bytes;
    ++n;
  }
  if (ms) *ms = t;
  if (launches) *launches = n;
  if (bytes) *bytes = b;
  return NSB_OK;
}

// Delimit the region a profiler captures (ncu --profile-from-start off).
extern "C" int nsb_profiler_start(void) {
  NSB_CUDA(cudaProfilerStart());
  return NSB_OK;
}
extern "C" int nsb_profiler_stop(void) {
  NSB_CUDA(cudaProfilerStop());
  return NSB_OK;
}

extern "C" const char *nsb_last_error(void) { return g_err; }
extern "C" int nsb_version(void) { return NSB_VERSION; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_init(int device, int rank, int nranks, const void *unique_id,
                        nsb_context_t *out) {
  NSB_REQUIRE(out != nullptr, "nsb_init: ctx is NULL");
  NSB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "nsb_init: bad rank %d/%d", rank, nranks);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("nsb_init: no CUDA device (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return NSB_ENODEVICE;
  }
  NSB_REQUIRE(device >= 0 && device < ndev, "nsb_init: device %d out of range (%d)", device, ndev);
  NSB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  NSB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("nsb_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
              prop.major, prop.minor);
    return NSB_ENODEVICE;
  }
  nsb_context_t ctx = new nsb_context_s();
  ctx->device = device;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->num_sms = prop.multiProcessorCount;
  if (const char *e = getenv("NSB_NO_FUSED")) ctx->no_fused = e[0] == '1';
  if (const char *e = getenv("NSB_FOLD_NORM")) ctx->fold_norm = e[0] != '0';
  if (const char *e = getenv("NSB_NS_NO_COARSE")) ctx->ns_no_coarse = e[0] == '1';
  if (const char *e = getenv("NSB_NS_GENERIC")) ctx->ns_generic = e[0] == '1';
  if (const char *e = getenv("NSB_FUSED_LOADER")) ctx->fused_loader = atoi(e);
  if (const char *e = getenv("NSB_FUSED_RC")) ctx->fused_rc = atoi(e);
  if (const char *e = getenv("NSB_FUSED_REG_MIN_K")) ctx->fused_reg_min_k = atoi(e);
  if (const char *e = getenv("NSB_AX_GENERIC")) ctx->ax_generic = e[0] == '1';
  if (const char *e = getenv("NSB_AX_RING")) ctx->ax_ring = e[0] != '0';
  if (const char *e = getenv("NSB_AX_STAGES")) ctx->ax_stages = atoi(e);
  if (const char *e = getenv("NSB_FUSED_PRIV")) ctx->fused_priv = e[0] != '0';
  if (const char *e = getenv("NSB_AX_DMMA")) ctx->ax_dmma = e[0] != '0';
  if (const char *e = getenv("NSB_PIPELINE_UPLOAD")) ctx->pipeline_upload = e[0] != '0';
  if (const char *e = getenv("NSB_ROTATE_SIMPLE")) ctx->rotate_simple = e[0] == '1';
  if (const char *e = getenv("NSB_ROTATE_DMMA")) ctx->rotate_dmma = e[0] != '0';
  if (const char *e = getenv("NSB_FUSED_ALLWARPS")) ctx->fused_allwarps = e[0] != '0';
  if (const char *e = getenv("NSB_TAIL")) ctx->tail = e[0] != '0';
  if (const char *e = getenv("NSB_HALO_FUSED")) ctx->halo_fused = e[0] != '0';
  if (const char *e = getenv("NSB_GRAPH")) ctx->use_graph = e[0] != '0';
  NSB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  NSB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  NSB_CUDA(cudaStreamCreateWithFlags(&ctx->gs_stream, cudaStreamNonBlocking));
  if (const char *e = getenv("NSB_AX_SLAB_MB")) ctx->ax_slab_mb = atof(e);
  NSB_CUDA(cudaEventCreate(&ctx->ev0));
  NSB_CUDA(cudaEventCreate(&ctx->ev1));
  NSB_CUDA(cudaMalloc(&ctx->hvec_d, sizeof(double) * 4 * (kMaxK + 8)));
  NSB_CUDA(cudaMemset(ctx->hvec_d, 0, sizeof(double) * 4 * (kMaxK + 8)));
  NSB_CUDA(cudaMallocHost(&ctx->hpin, sizeof(double) * 4 * (kMaxK + 8)));
  NSB_CUDA(cudaMalloc(&ctx->ticket_d, sizeof(unsigned int) * 8));
  NSB_CUDA(cudaMemset(ctx->ticket_d, 0, sizeof(unsigned int) * 8));
  NSB_CUDA(cudaMalloc(&ctx->flag_d, sizeof(int) * 4));
  NSB_CUDA(cudaMemset(ctx->flag_d, 0, sizeof(int) * 4));
  NSB_CUDA(cudaMalloc(&ctx->seq_d, sizeof(unsigned long long) * 2));
  NSB_CUDA(cudaMemset(ctx->seq_d, 0, sizeof(unsigned long long) * 2));
  NSB_CUDA(cudaHostAlloc((void **)&ctx->dev_err, sizeof(int) * 4, cudaHostAllocMapped));
  ctx->dev_err[0] = 0;
  NSB_CUDA(cudaHostGetDevicePointer((void **)&ctx->dev_err_d, ctx->dev_err, 0));
  NSB_CHECK(ensure_partial(ctx, (int64_t)ctx->num_sms * 8));
  if (nranks > 1) {
    NSB_REQUIRE(unique_id != nullptr, "nsb_init: unique_id required for nranks > 1");
    int r = comm_init(ctx, unique_id);
    if (r != NSB_OK) return r;
  }
  *out = ctx;
  return NSB_OK;
}

extern "C" int nsb_finalize(nsb_context_t ctx) {
  if (!ctx) return NSB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ctx->nranks && r < nsb_context_s::kMaxPeers; ++r)
    if (r != ctx->rank && ctx->peer_mail[r]) cudaIpcCloseMemHandle(ctx->peer_mail[r]);
  if (ctx->mail_d) cudaFree(ctx->mail_d);
  comm_destroy(ctx);
  clear_step_graphs(ctx);
  if (ctx->hstage) cudaFreeHost(ctx->hstage);
  if (ctx->rot_d) cudaFree(ctx->rot_d);
  if (ctx->gram_d) cudaFree(ctx->gram_d);
  if (ctx->ticket_d) cudaFree(ctx->ticket_d);
  if (ctx->flag_d) cudaFree(ctx->flag_d);
  if (ctx->seq_d) cudaFree(ctx->seq_d);
  if (ctx->dev_err) cudaFreeHost(ctx->dev_err);
  if (ctx->partial_d) cudaFree(ctx->partial_d);
  if (ctx->hvec_d) cudaFree(ctx->hvec_d);
  if (ctx->hpin) cudaFreeHost(ctx->hpin);
  if (ctx->flush_d) cudaFree(ctx->flush_d);
  for (auto &r : ctx->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->gs_stream);
  delete ctx;
  return NSB_OK;
}

extern "C" int nsb_sync(nsb_context_t ctx) {
  NSB_REQUIRE(ctx, "nsb_sync: NULL context");
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return check_dev_err(ctx);
}

extern "C" int nsb_rank(nsb_context_t ctx, int *rank, int *nranks) {
  NSB_REQUIRE(ctx, "nsb_rank: NULL context");
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return NSB_OK;
}

extern "C" int nsb_stream(nsb_context_t ctx, uint64_t *stream) {
  NSB_REQUIRE(ctx && stream, "nsb_stream: NULL argument");
  *stream = (uint64_t)(uintptr_t)ctx->stream;
  return NSB_OK;
}

extern "C" int nsb_launch_count(nsb_context_t ctx, int64_t *count) {
  NSB_REQUIRE(ctx && count, "nsb_launch_count: NULL argument");
  *count = ctx->launches;
  return NSB_OK;
}

extern "C" int nsb_timer_start(nsb_context_t ctx) {
  NSB_REQUIRE(ctx, "nsb_timer_start: NULL context");
  NSB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return NSB_OK;
}

extern "C" int nsb_timer_stop(nsb_context_t ctx, double *ms) {
  NSB_REQUIRE(ctx && ms, "nsb_timer_stop: NULL argument");
  NSB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  NSB_CUDA(cudaEventSynchronize(ctx->ev1));
  float f = 0.f;
  NSB_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = (double)f;
  return NSB_OK;
}

extern "C" int nsb_allreduce_host(nsb_context_t ctx, double *x, int n) {
  NSB_REQUIRE(ctx && x && n >= 0 && n <= 4 * (kMaxK + 8), "nsb_allreduce_host: bad argument");
  if (ctx->nranks == 1 || n == 0) return NSB_OK;
  NSB_CUDA(cudaMemcpyAsync(ctx->hvec_d, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  NSB_CHECK(allreduce_sum_d(ctx, ctx->hvec_d, n));
  NSB_CUDA(cudaMemcpyAsync(x, ctx->hvec_d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return NSB_OK;
}

extern "C" int nsb_flush_l2(nsb_context_t ctx) {
  NSB_REQUIRE(ctx, "nsb_flush_l2: NULL context");
  if (!ctx->flush_d) {
    ctx->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    NSB_CUDA(cudaMalloc(&ctx->flush_d, ctx->flush_bytes));
  }
  NSB_CUDA(cudaMemsetAsync(ctx->flush_d, 0, ctx->flush_bytes, ctx->stream));
  return NSB_OK;
}

extern "C" int nsb_host_alloc(void **ptr, int64_t bytes) {
  NSB_REQUIRE(ptr && bytes > 0, "nsb_host_alloc: bad argument");
  NSB_CUDA(cudaMallocHost(ptr, (size_t)bytes));
  return NSB_OK;
}

extern "C" int nsb_host_free(void *ptr) {
  if (ptr) NSB_CUDA(cudaFreeHost(ptr));
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

extern "C" int nsb_layout_create(nsb_context_t ctx, int nfields, const int64_t *field_len,
                                 const int *field_in_dot, int time_in_dot, nsb_layout_t *out) {
  NSB_REQUIRE(ctx && out && field_len && field_in_dot, "nsb_layout_create: NULL argument");
  NSB_REQUIRE(nfields >= 1 && nfields <= 64, "nsb_layout_create: nfields=%d", nfields);
  nsb_layout_t L = new nsb_layout_s();
  L->ctx = ctx;
  L->nfields = nfields;
  L->len.assign(field_len, field_len + nfields);
  L->hlen = L->len;
  L->in_dot.assign(field_in_dot, field_in_dot + nfields);
  L->off.assign(nfields, 0);
  L->time_in_dot = time_in_dot ? 1 : 0;
  // Column = [ in-dot fields | time row | pad ] [ other fields | pad ]; field starts are 128-byte
  // aligned; pad rows are zero in every column and carry zero weight.
  int64_t row = 0;
  for (int f = 0; f < nfields; ++f) {
    if (field_len[f] < 0) {
      delete L;
      set_error("nsb_layout_create: negative field length");
      return NSB_EINVAL;
    }
    if (!L->in_dot[f]) continue;
    L->off[f] = row;
    row = round_up(row + field_len[f], 16);
    L->ndof_dot += field_len[f];
  }
  L->time_row = row;
  row += 1;
  L->ndot = round_up(row, kRowPad);
  row = L->ndot;
  for (int f = 0; f < nfields; ++f) {
    if (L->in_dot[f]) continue;
    L->off[f] = row;
    row = round_up(row + field_len[f], 16);
  }
  L->ld = round_up(row, kRowPad);
  L->nact = 1;
  for (int f = 0; f < nfields; ++f) L->nact += field_len[f];
  cudaSetDevice(ctx->device);
  NSB_CUDA(cudaMalloc(&L->w_d, sizeof(double) * L->ndot));
  NSB_CUDA(cudaMemsetAsync(L->w_d, 0, sizeof(double) * L->ndot, ctx->stream));
  *out = L;
  return NSB_OK;
}

extern "C" int nsb_layout_destroy(nsb_layout_t L) {
  if (!L) return NSB_OK;
  cudaSetDevice(L->ctx->device);
  cudaStreamSynchronize(L->ctx->stream);
  clear_step_graphs(L->ctx);
  if (L->w_d) cudaFree(L->w_d);
  delete L;
  return NSB_OK;
}

extern "C" int nsb_layout_info(nsb_layout_t L, int64_t *ld, int64_t *ndot, int64_t *ndof_dot) {
  NSB_REQUIRE(L, "nsb_layout_info: NULL layout");
  if (ld) *ld = L->ld;
  if (ndot) *ndot = L->ndot;
  if (ndof_dot) *ndof_dot = L->ndof_dot;
  return NSB_OK;
}

extern "C" int nsb_layout_set_weight(nsb_layout_t L, const double *const *w) {
  NSB_REQUIRE(L && w, "nsb_layout_set_weight: NULL argument");
  nsb_context_t ctx = L->ctx;
  cudaSetDevice(ctx->device);
  int j = 0;
  for (int f = 0; f < L->nfields; ++f) {
    if (!L->in_dot[f]) continue;
    NSB_REQUIRE(w[j] != nullptr, "nsb_layout_set_weight: weight %d is NULL", j);
    if (f < L->c0_nfields) {
      NSB_CHECK(c0_set_weight(L, f, w[j]));          // assembled weight on the distinct nodes
    } else {
      NSB_CUDA(cudaMemcpyAsync(L->w_d + L->off[f], w[j], sizeof(double) * L->len[f],
                               cudaMemcpyHostToDevice, ctx->stream));
    }
    ++j;
  }
  // %time enters the dot once globally: weight 1 on rank 0 only (the reference adds
  // time*time after the gop, core/nek_vectors.f90:105-107).
  double tw = (L->time_in_dot && ctx->rank == 0) ? 1.0 : 0.0;
  NSB_CUDA(cudaMemcpyAsync(L->w_d + L->time_row, &tw, sizeof(double), cudaMemcpyHostToDevice,
                           ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// basis
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_basis_create(nsb_layout_t L, int ncols, nsb_basis_t *out) {
  NSB_REQUIRE(L && out && ncols >= 1, "nsb_basis_create: bad argument");
  cudaSetDevice(L->ctx->device);
  nsb_basis_t B = new nsb_basis_s();
  B->lay = L;
  B->ncols = ncols;
  size_t bytes = sizeof(double) * (size_t)L->ld * (size_t)ncols;
  cudaError_t e = cudaMalloc(&B->v_d, bytes);
  if (e != cudaSuccess) {
    set_error("nsb_basis_create: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    delete B;
    return NSB_ECUDA;
  }
  NSB_CUDA(cudaMemsetAsync(B->v_d, 0, bytes, L->ctx->stream));
  *out = B;
  return NSB_OK;
}

extern "C" int nsb_basis_destroy(nsb_basis_t B) {
  if (!B) return NSB_OK;
  cudaSetDevice(B->lay->ctx->device);
  cudaStreamSynchronize(B->lay->ctx->stream);
  clear_step_graphs(B->lay->ctx);
  if (B->v_d) cudaFree(B->v_d);
  delete B;
  return NSB_OK;
}

extern "C" int nsb_basis_ncols(nsb_basis_t B, int *ncols) {
  NSB_REQUIRE(B && ncols, "nsb_basis_ncols: NULL argument");
  *ncols = B->ncols;
  return NSB_OK;
}

#define NSB_COL_OK(B, c, name)                                                              \
  NSB_REQUIRE((B) != nullptr && (c) >= 0 && (c) < (B)->ncols, name ": column %d out of range", \
              (int)(c))

extern "C" int nsb_basis_col_ptr(nsb_basis_t B, int col, uint64_t *ptr) {
  NSB_COL_OK(B, col, "nsb_basis_col_ptr");
  NSB_REQUIRE(ptr, "nsb_basis_col_ptr: NULL");
  *ptr = (uint64_t)(uintptr_t)B->col(col);
  return NSB_OK;
}

extern "C" int nsb_vec_upload(nsb_basis_t B, int col, const double *const *fields, double time) {
  NSB_COL_OK(B, col, "nsb_vec_upload");
  NSB_REQUIRE(fields, "nsb_vec_upload: NULL fields");
  nsb_layout_t L = B->lay;
  cudaStream_t s = L->ctx->stream;
  cudaSetDevice(L->ctx->device);
  double *c = B->col(col);
  NSB_CUDA(cudaMemsetAsync(c, 0, sizeof(double) * L->ld, s));
  if (L->c0_nfields) NSB_CHECK(c0_upload(B, col, fields));
  for (int f = L->c0_nfields; f < L->nfields; ++f) {
    if (!fields[f] || L->len[f] == 0) continue;
    NSB_CUDA(cudaMemcpyAsync(c + L->off[f], fields[f], sizeof(double) * L->len[f],
                             cudaMemcpyHostToDevice, s));
  }
  NSB_CUDA(cudaMemcpyAsync(c + L->time_row, &time, sizeof(double), cudaMemcpyHostToDevice, s));
  NSB_CUDA(cudaStreamSynchronize(s));  // host buffers may be reused by the caller
  return NSB_OK;
}

extern "C" int nsb_vec_download(nsb_basis_t B, int col, double *const *fields, double *time) {
  NSB_COL_OK(B, col, "nsb_vec_download");
  nsb_layout_t L = B->lay;
  cudaStream_t s = L->ctx->stream;
  cudaSetDevice(L->ctx->device);
  const double *c = B->col(col);
  if (fields && L->c0_nfields) NSB_CHECK(c0_download(B, col, fields));
  if (fields)
    for (int f = L->c0_nfields; f < L->nfields; ++f) {
      if (!fields[f] || L->len[f] == 0) continue;
      NSB_CUDA(cudaMemcpyAsync(fields[f], c + L->off[f], sizeof(double) * L->len[f],
                               cudaMemcpyDeviceToHost, s));
    }
  if (time)
    NSB_CUDA(cudaMemcpyAsync(time, c + L->time_row, sizeof(double), cudaMemcpyDeviceToHost, s));
  NSB_CUDA(cudaStreamSynchronize(s));
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// streaming BLAS-1 kernels
// ------------------------------------------------------------------------------------------------
namespace {

enum { OP_SCAL, OP_AXPBY, OP_ADD2, OP_SUB2, OP_SUB3, OP_COPY };

// One column is n2 double2's (ld is a multiple of 1024 rows).  Each thread moves 4 x 16 B per
// operand per iteration, all loads issued before the first use.
template <int OP>
__global__ void __launch_bounds__(256) blas1_kernel(double2 *__restrict__ x,
                                                    const double2 *__restrict__ y,
                                                    const double2 *__restrict__ z, double a,
                                                    double b, int64_t n2, int64_t skip2, int skip_lane) {
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n2; base += stride) {
    double2 xv[U], yv[U], zv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n2) {
        if (OP != OP_COPY && OP != OP_SUB3) xv[u] = x[i];
        if (OP != OP_SCAL) yv[u] = ld_stream(y + i);
        if (OP == OP_SUB3) zv[u] = ld_stream(z + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n2) {
        double2 r;
        if (OP == OP_SCAL) { r.x = xv[u].x * a; r.y = xv[u].y * a; }
        if (OP == OP_AXPBY) { r.x = xv[u].x * a + yv[u].x * b; r.y = xv[u].y * a + yv[u].y * b; }
        if (OP == OP_ADD2) { r.x = xv[u].x + yv[u].x; r.y = xv[u].y + yv[u].y; }
        if (OP == OP_SUB2) { r.x = xv[u].x - yv[u].x; r.y = xv[u].y - yv[u].y; }
        if (OP == OP_SUB3) { r.x = yv[u].x - zv[u].x; r.y = yv[u].y - zv[u].y; }
        if (OP == OP_COPY) r = yv[u];
        if (OP == OP_AXPBY && i == skip2) {  // real_axpby leaves %time untouched (quirk)
          if (skip_lane == 0) r.x = xv[u].x; else r.y = xv[u].y;
        }
        x[i] = r;
      }
    }
  }
}

template <int OP>
int launch_blas1(nsb_context_t ctx, double *x, const double *y, const double *z, double a, double b,
                 int64_t n, int64_t skip_row = -1) {
  int64_t n2 = n / 2;
  int64_t per_cta = 256 * 4;
  int64_t want = (n2 + per_cta - 1) / per_cta;
  int64_t cap = (int64_t)ctx->num_sms * 16;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  cudaSetDevice(ctx->device);
  ProfScope ps(ctx, PC_BLAS1, 8.0 * n * (OP == OP_SCAL ? 2 : 3));
  blas1_kernel<OP><<<grid, 256, 0, ctx->stream>>>(
      reinterpret_cast<double2 *>(x), reinterpret_cast<const double2 *>(y),
      reinterpret_cast<const double2 *>(z), a, b, n2, skip_row >= 0 ? skip_row / 2 : -1,
      skip_row >= 0 ? (int)(skip_row & 1) : 0);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// weighted dot partials: partial[cta] = sum_i a_i W_i b_i over the dot prefix
__global__ void __launch_bounds__(256) wdot_kernel(const double2 *__restrict__ a,
                                                   const double2 *__restrict__ b,
                                                   const double2 *__restrict__ w, int64_t n2,
                                                   double *__restrict__ partial) {
  constexpr int U = 4;
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n2; base += stride) {
    double2 av[U], bv[U], wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n2) { av[u] = ld_stream(a + i); bv[u] = ld_stream(b + i); wv[u] = ld_stream(w + i); }
      else { av[u] = bv[u] = wv[u] = make_double2(0.0, 0.0); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc = fma(av[u].x * wv[u].x, bv[u].x, acc);
      acc = fma(av[u].y * wv[u].y, bv[u].y, acc);
    }
  }
  acc = block_reduce_sum<256>(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

}  // namespace

namespace nsb {
// out[j] = sum_b partial[b * pstride + j], fixed order -> deterministic.  One warp per column.
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nblk, int pstride,
                                       int k, double *__restrict__ out, int accumulate_into,
                                       double *__restrict__ out2) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= k) return;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += partial[(size_t)b * pstride + warp];
  s = warp_reduce_sum(s);
  if (lane == 0) {
    out[warp] = s;
    if (accumulate_into) out2[warp] += s;
  }
}
}  // namespace nsb

static int dot_device(nsb_basis_t ba, int ca, nsb_basis_t bb, int cb, double *out_d) {
  nsb_layout_t L = ba->lay;
  nsb_context_t ctx = L->ctx;
  int64_t n2 = L->ndot / 2;
  int64_t want = (n2 + 1023) / 1024;
  int64_t cap = (int64_t)ctx->num_sms * 8;
  int grid = (int)(want < cap ? want : cap);
  cudaSetDevice(ctx->device);
  ProfScope ps(ctx, PC_DOT, 24.0 * (L->ndof_dot + 1));
  wdot_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const double2 *>(ba->col(ca)),
                                             reinterpret_cast<const double2 *>(bb->col(cb)),
                                             reinterpret_cast<const double2 *>(L->w_d), n2,
                                             ctx->partial_d);
  reduce_partials_kernel<<<1, 32, 0, ctx->stream>>>(ctx->partial_d, grid, 1, 1, out_d, 0, nullptr);
  ctx->launches += 2;
  NSB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, out_d, 1));
  return NSB_OK;
}

#define NSB_SAME_LAYOUT(a, b, name) \
  NSB_REQUIRE((a)->lay == (b)->lay, name ": vectors belong to different layouts")

extern "C" int nsb_vec_zero(nsb_basis_t B, int col) {
  NSB_COL_OK(B, col, "nsb_vec_zero");
  cudaSetDevice(B->lay->ctx->device);
  NSB_CUDA(cudaMemsetAsync(B->col(col), 0, sizeof(double) * B->lay->ld, B->lay->ctx->stream));
  return NSB_OK;
}

extern "C" int nsb_vec_copy(nsb_basis_t bd, int cd, nsb_basis_t bs, int cs) {
  NSB_COL_OK(bd, cd, "nsb_vec_copy");
  NSB_COL_OK(bs, cs, "nsb_vec_copy");
  NSB_SAME_LAYOUT(bd, bs, "nsb_vec_copy");
  if (bd == bs && cd == cs) return NSB_OK;
  return launch_blas1<OP_COPY>(bd->lay->ctx, bd->col(cd), bs->col(cs), nullptr, 0, 0, bd->lay->ld);
}

extern "C" int nsb_vec_scal(nsb_basis_t B, int col, double alpha) {
  NSB_COL_OK(B, col, "nsb_vec_scal");
  return launch_blas1<OP_SCAL>(B->lay->ctx, B->col(col), nullptr, nullptr, alpha, 0, B->lay->ld);
}

extern "C" int nsb_vec_axpby(nsb_basis_t bx, int cx, double alpha, nsb_basis_t by, int cy,
                             double beta, int flags) {
  NSB_COL_OK(bx, cx, "nsb_vec_axpby");
  NSB_COL_OK(by, cy, "nsb_vec_axpby");
  NSB_SAME_LAYOUT(bx, by, "nsb_vec_axpby");
  int64_t skip = (flags & NSB_AXPBY_SKIP_TIME) ? bx->lay->time_row : -1;
  return launch_blas1<OP_AXPBY>(bx->lay->ctx, bx->col(cx), by->col(cy), nullptr, alpha, beta,
                                bx->lay->ld, skip);
}

extern "C" int nsb_vec_add2(nsb_basis_t bp, int cp, nsb_basis_t bq, int cq) {
  NSB_COL_OK(bp, cp, "nsb_vec_add2");
  NSB_COL_OK(bq, cq, "nsb_vec_add2");
  NSB_SAME_LAYOUT(bp, bq, "nsb_vec_add2");
  return launch_blas1<OP_ADD2>(bp->lay->ctx, bp->col(cp), bq->col(cq), nullptr, 0, 0, bp->lay->ld);
}

extern "C" int nsb_vec_sub2(nsb_basis_t bp, int cp, nsb_basis_t bq, int cq) {
  NSB_COL_OK(bp, cp, "nsb_vec_sub2");
  NSB_COL_OK(bq, cq, "nsb_vec_sub2");
  NSB_SAME_LAYOUT(bp, bq, "nsb_vec_sub2");
  return launch_blas1<OP_SUB2>(bp->lay->ctx, bp->col(cp), bq->col(cq), nullptr, 0, 0, bp->lay->ld);
}

extern "C" int nsb_vec_sub3(nsb_basis_t bp, int cp, nsb_basis_t bq, int cq, nsb_basis_t br, int cr) {
  NSB_COL_OK(bp, cp, "nsb_vec_sub3");
  NSB_COL_OK(bq, cq, "nsb_vec_sub3");
  NSB_COL_OK(br, cr, "nsb_vec_sub3");
  NSB_SAME_LAYOUT(bp, bq, "nsb_vec_sub3");
  NSB_SAME_LAYOUT(bp, br, "nsb_vec_sub3");
  return launch_blas1<OP_SUB3>(bp->lay->ctx, bp->col(cp), bq->col(cq), br->col(cr), 0, 0,
                               bp->lay->ld);
}

extern "C" int nsb_vec_dot(nsb_basis_t ba, int ca, nsb_basis_t bb, int cb, double *alpha) {
  NSB_COL_OK(ba, ca, "nsb_vec_dot");
  NSB_COL_OK(bb, cb, "nsb_vec_dot");
  NSB_SAME_LAYOUT(ba, bb, "nsb_vec_dot");
  NSB_REQUIRE(alpha, "nsb_vec_dot: NULL result");
  nsb_context_t ctx = ba->lay->ctx;
  double *out_d = ctx->hvec_d + 3 * (kMaxK + 8);
  NSB_CHECK(dot_device(ba, ca, bb, cb, out_d));
  NSB_CUDA(cudaMemcpyAsync(ctx->hpin, out_d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  *alpha = ctx->hpin[0];
  if (std::isnan(*alpha)) {
    set_error("NaN detected in dot product");  // core/nek_vectors.f90:108-111
    return NSB_ENAN;
  }
  return NSB_OK;
}

extern "C" int nsb_vec_norm(nsb_basis_t B, int col, double *alpha) {
  NSB_CHECK(nsb_vec_dot(B, col, B, col, alpha));
  *alpha = std::sqrt(*alpha);
  return NSB_OK;
}

extern "C" int nsb_vec_normalize(nsb_basis_t B, int col, double *alpha) {
  double a = 0.0;
  NSB_CHECK(nsb_vec_norm(B, col, &a));
  if (alpha) *alpha = a;
  return nsb_vec_scal(B, col, 1.0 / a);  // k_normalize: inv_alpha = 1/alpha; k_cmult
}
