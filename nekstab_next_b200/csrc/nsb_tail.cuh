// Kernel tails: what used to be three extra launches per inner product (reduce_partials -> peer-memory
// all-reduce -> add into H) now runs inside the sweep kernel itself.  Every CTA writes its partial row,
// takes a ticket, and the CTA that draws the last ticket
//   1. sums the partial rows in a fixed order (bitwise reproducible whichever CTA is last),
//   2. for nranks > 1 exchanges the vector through the NVLink mailboxes (stores into every peer's slot,
//      release flag, acquire spin, sum in rank order -> identical bits on every rank),
//   3. applies the bookkeeping of update_hessenberg_matrix (core/krylov_decomposition.f90:155-186): H column
//      accumulation, the norm for k_normalize, and the DGKS re-orthogonalisation decision -- on the device,
//      so the Arnoldi loop has no host round trip and a whole step can be replayed as a CUDA graph.
// The all-reduce sequence number lives in device memory for the same reason (a kernel skipped on the device
// must not consume a number: the two mailbox slots rely on consecutive collectives alternating).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nsb_internal.h"
#include "nsb_device.cuh"

namespace nsb {

constexpr int ARN = kMaxK + 8;                    // doubles per all-reduce vector slot
struct PeerPtrs { double *p[nsb_context_s::kMaxPeers]; };

// mailbox layout (in doubles): [ar data 2*P*ARN][ar flags 2*P][reserved 2*P][halo area ...]
__host__ __device__ inline size_t mb_ar_data(int P, int slot, int r) { return ((size_t)slot * P + r) * ARN; }
__host__ __device__ inline size_t mb_ar_flag(int P, int slot, int r) { return (size_t)2 * P * ARN + (size_t)slot * P + r; }
__host__ __device__ inline size_t mb_halo_base(int P) { return (((size_t)2 * P * ARN + 4 * P) + 31) & ~(size_t)31; }

__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// peer-written data: a coherent load (never the read-only / non-coherent path)
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Spin until *flag == seq.  Bounded: after kSpinLimitNs the device error word is set and the wait is
// abandoned, so a peer that never arrives (lost rank, failed mapping) surfaces as an error at the next
// synchronising call instead of a hung GPU.
constexpr unsigned long long kSpinLimitNs = 120ull * 1000000000ull;
enum DevErr { DEVERR_NONE = 0, DEVERR_ALLREDUCE_TIMEOUT = 1, DEVERR_HALO_TIMEOUT = 2 };
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool spin_until(const uint64_t *flag, uint64_t seq, int *err, int code) {
  unsigned long long t0 = 0;
  for (unsigned it = 0;; ++it) {
    if (ld_acquire_sys(flag) == seq) return true;
    if ((it & 1023u) == 1023u) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kSpinLimitNs) break;
    }
  }
  if (err) atomicExch_system(err, code);
  return false;
}

// One CTA sums buf[0..n) over all ranks in place.  Called by every thread of the CTA.
struct PeerComm {
  int P = 1, rank = 0;
  unsigned long long *seq = nullptr;   // device counter of the all-reduces issued on this context
  int *err = nullptr;                  // device-visible error word
  PeerPtrs mail;
};

__device__ __forceinline__ void p2p_allreduce_block(double *buf, int n, const PeerComm &c) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = ++(*c.seq);   // a single CTA of a single kernel at a time touches the counter
  __syncthreads();
  const uint64_t seq = s_seq;
  const int slot = (int)(seq & 1), P = c.P;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = buf[j];
    for (int r = 0; r < P; ++r) c.mail.p[r][mb_ar_data(P, slot, c.rank) + j] = v;   // peer stores over NVLink
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < P) {
    st_release_sys(reinterpret_cast<uint64_t *>(c.mail.p[threadIdx.x] + mb_ar_flag(P, slot, c.rank)), seq);
    spin_until(reinterpret_cast<const uint64_t *>(c.mail.p[c.rank] + mb_ar_flag(P, slot, threadIdx.x)), seq, c.err,
               DEVERR_ALLREDUCE_TIMEOUT);
  }
  __syncthreads();
  const double *mine = c.mail.p[c.rank];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < P; ++r) s += ld_relaxed_sys(mine + mb_ar_data(P, slot, r) + j);
    buf[j] = s;
  }
  __syncthreads();
}

// What the last CTA does with the reduced vector out[0..kout):  out[0..k) are projection coefficients, out[k]
// (if kout == k + 1) a squared norm.
//   hsum_op : 0 none, 1 hsum[0..k) = out, 2 hsum[0..k) += out
//   norm_op : 0 none
//             1 scal[0] = out[k]                      (norm^2 that normalize_kernel divides by)
//             2 scal[1] = out[k]                      (|w|^2 before the first projection, kept for the DGKS test)
//             3 DGKS decision: second = !(out[k] >= eta^2 scal[1])  (|w'| < eta |w|, eta = 1/sqrt 2 by default;
//               also taken on NaN);
//               *flag = second; if (!second) scal[0] = out[k] and hsum_op is NOT applied (the second
//               projection is dropped as a whole); *passes_out = number of passes (read back by the host)
//             4 folded normalisation (CGS2): out[k] = |w'|^2 rode along with the second projection h2 = out[0..k);
//               because V is B-orthonormal, |w' - V h2|^2 = |w'|^2 - |h2|^2, so the norm of the FINAL vector is
//               known before the third sweep: scal[0] = beta^2 = out[k] - sum_j out[j]^2, hsum[k] = beta
//               (H(k+1,k)), scal[3] = beta.  The third sweep then writes (w' - V h2) / beta directly: no norm
//               reduction, no third all-reduce and no separate normalisation pass.
//             5 norm_op 3 and 4 together (DGKS with the folded normalisation): decision as in 3; beta^2 as in 4 when the
//               second projection is kept, |w'|^2 when it is dropped (then normalize_kernel runs instead of the sweep)
struct OrthTail {
  double *partial = nullptr;   // [gridDim.x][pstride]
  int pstride = 0;
  unsigned int *ticket = nullptr;   // nullptr: no tail (the caller reduces the partial rows itself)
  double *out = nullptr;
  double *hsum = nullptr;
  double *scal = nullptr;
  int *flag = nullptr;
  double *passes_out = nullptr;     // norm_op 3: number of projection passes taken (1.0 or 2.0), for the host
  const int *skip_flag = nullptr;   // kernel (and tail) do nothing when *skip_flag == 0 (DGKS: pass not needed)
  int k = 0, kout = 0;
  int hsum_op = 0, norm_op = 0;
  int exchange = 0;                 // 1: all-reduce through the peer mailboxes inside the tail
  double eta2 = 0.5;                // DGKS threshold eta^2 (|w'|^2 >= eta^2 |w|^2: one pass is enough)
  PeerComm comm;
};

__device__ __forceinline__ void orth_post_ops(const OrthTail &t) {
  __shared__ int s_second;
  if (threadIdx.x == 0) {
    int second = 1;
    if (t.norm_op == 1) t.scal[0] = t.out[t.k];
    if (t.norm_op == 2) t.scal[1] = t.out[t.k];
    if (t.norm_op == 3 || t.norm_op == 5) {
      const double n1 = t.out[t.k], n0 = t.scal[1];
      second = !(n1 >= t.eta2 * n0);
      *t.flag = second;
      if (!second && t.norm_op == 3) t.scal[0] = n1;
      if (t.passes_out) *t.passes_out = second ? 2.0 : 1.0;
    }
    s_second = second;
  }
  if (t.norm_op == 4 || t.norm_op == 5) {
    // |h2|^2 with a fixed summation order (thread-strided, shuffle tree, warp partials in turn)
    __shared__ double s_red[32];
    double s = 0.0;
    for (int j = threadIdx.x; j < t.k; j += blockDim.x) s = fma(t.out[j], t.out[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double h2n = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) h2n += s_red[i];
      // The identity loses accuracy only as beta^2 / |w'|^2 -> 0 (relative error ~ eps |w'|^2 / beta^2), i.e. when the
      // second projection removed w' altogether: w' was rounding noise (f in span(V) to working precision).  The
      // reference measures the norm of that noise and carries on; here beta is kept at the rounding level of |w'|
      // (never below eps |w'|), so the step neither divides by zero nor reports a breakdown the reference would
      // not see.  |w'| = 0 exactly still gives beta = 0 (NSB_EBREAKDOWN); NaN propagates.
      const double n1 = t.out[t.k];
      // norm_op 5 (DGKS): the second projection is dropped when |w'| >= eta |w| -- then beta = |w'|
      double b2 = (t.norm_op == 4 || s_second) ? n1 - h2n : n1;
      const double floor2 = 4.930380657631324e-32 * n1;   // (2^-52)^2 |w'|^2
      if (b2 == b2 && !(b2 > floor2)) b2 = floor2;
      const double beta = sqrt(b2);
      t.scal[0] = b2;
      t.scal[3] = beta;
      t.hsum[t.k] = beta;
    }
  }
  __syncthreads();
  if (t.hsum_op == 1)
    for (int j = threadIdx.x; j < t.k; j += blockDim.x) t.hsum[j] = t.out[j];
  if (t.hsum_op == 2 && s_second)
    for (int j = threadIdx.x; j < t.k; j += blockDim.x) t.hsum[j] += t.out[j];
}

// Called by ALL threads of EVERY CTA after the CTA's partial row has been written with plain stores.
__device__ __forceinline__ void orth_tail(const OrthTail &t) {
  if (!t.ticket) return;
  __shared__ unsigned int s_last;
  __threadfence();                       // this CTA's partial row is visible device-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(t.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // Same summation order as reduce_partials_kernel (per lane the rows b = lane, lane + 32, ... in turn, then the
  // shuffle tree), but a warp takes FOUR columns per round and the row loop is unrolled, so 16 independent L2
  // loads are in flight per lane: one warp walking one column with dependent loads cost ~50 us per k-column
  // tail (8 GPUs, round 2: multidot / fused 12-16 % over 1/8 of their single-GPU time, update -- a one-column
  // tail -- exactly on it).  Columns past kout are read (inside the partial buffer) and dropped.
  for (int j0 = 4 * warp; j0 < t.kout; j0 += 4 * nw) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
    for (unsigned b = lane; b < gridDim.x; b += 32) {
      const double *p = t.partial + (size_t)b * t.pstride + j0;
      s0 += __ldcg(p);
      s1 += __ldcg(p + 1);
      s2 += __ldcg(p + 2);
      s3 += __ldcg(p + 3);
    }
    s0 = warp_reduce_sum(s0);
    s1 = warp_reduce_sum(s1);
    s2 = warp_reduce_sum(s2);
    s3 = warp_reduce_sum(s3);
    if (lane == 0) {
      t.out[j0] = s0;
      if (j0 + 1 < t.kout) t.out[j0 + 1] = s1;
      if (j0 + 2 < t.kout) t.out[j0 + 2] = s2;
      if (j0 + 3 < t.kout) t.out[j0 + 3] = s3;
    }
  }
  if (threadIdx.x == 0) *t.ticket = 0;   // ready for the next launch (stream order)
  __syncthreads();
  if (t.exchange) p2p_allreduce_block(t.out, t.kout, t.comm);
  if (t.comm.P > 1 && !t.exchange) return;   // NCCL transport: the host enqueues the all-reduce, then orth_post_kernel
  orth_post_ops(t);
}

}  // namespace nsb
