// Fused BM1-weighted orthogonalisation on the device-resident Krylov basis.
//
// Replaces update_hessenberg_matrix (core/krylov_decomposition.f90:103-189): the reference runs
// 2k {k_copy, k_dot, k_cmult, k_sub2} sweeps (10 n words each, one MPI_Allreduce per component
// per dot).  Here one orthogonalisation pass is two kernels that each read the basis exactly once:
//
//   multidot : h = V_k^T (W o w)       8 n (k+2) algorithmic bytes      (V, w, W read once)
//   update   : w -= V_k h (+ ||w||_W^2) 8 n (k+3) algorithmic bytes      (V, W read; w read+written)
//
// V is column-major [ld, ncols], ld a multiple of 1024 rows with zero pads, so no tail handling.
// Both kernels are persistent (grid = 2 CTAs per SM), 256 threads, each thread owning 4 rows of a
// 1024-row chunk as two double2; 8 columns (16 x 128-bit loads per thread) are in flight at once.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "nsb_internal.h"
#include "nsb_device.cuh"
#include "nsb_tail.cuh"

using namespace nsb;

namespace nsb {
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nblk, int pstride,
                                       int k, double *__restrict__ out, int accumulate_into,
                                       double *__restrict__ out2);
}

namespace {

constexpr int KT = 8;          // columns per tile
constexpr int CHUNK = 1024;    // rows per chunk = 256 threads x 4 rows
constexpr int NT = 256;

// Reduce 8 per-lane values over the 32 lanes of a warp with 9 shuffles (recursive halving):
// on return lane L with (L & 3) == 0 holds the warp sum of column (L >> 2).
__device__ __forceinline__ double warp_reduce8(const double (&a)[KT], int lane) {
  double b[4], c[2], d;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double send = h16 ? a[i] : a[i + 4];
    double keep = h16 ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    double send = h8 ? b[i] : b[i + 2];
    double keep = h8 ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    double send = h4 ? c[0] : c[1];
    double keep = h4 ? c[1] : c[0];
    d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// h_partial[cta][j] = sum over the CTA's chunks of sum_r V[r,j] W[r] w[r];
// column k of the partial (if WITH_NORM) = sum_r W[r] w[r]^2.
template <bool WITH_NORM>
__global__ void __launch_bounds__(NT, 2)
multidot_kernel(const double *__restrict__ V, int64_t ld, int k, const double *__restrict__ w,
                const double *__restrict__ W, int64_t nchunks, double *__restrict__ partial,
                int pstride, const __grid_constant__ OrthTail tail) {
  extern __shared__ double accS[];  // [NT/32][kpad]
  if (tail.skip_flag && *tail.skip_flag == 0) return;   // DGKS: second projection not needed
  const int kpad = (k + KT) & ~(KT - 1);  // room for the norm slot too
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (NT / 32) * kpad; i += NT) accS[i] = 0.0;
  __syncthreads();
  double *myacc = accS + warp * kpad;
  double nrm = 0.0;

  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t r0 = chunk * CHUNK + 2 * threadIdx.x;  // rows r0, r0+1, r0+512, r0+513
    const double2 wa = ld_stream(reinterpret_cast<const double2 *>(w + r0));
    const double2 wb = ld_stream(reinterpret_cast<const double2 *>(w + r0 + 512));
    const double2 Wa = ld_stream(reinterpret_cast<const double2 *>(W + r0));
    const double2 Wb = ld_stream(reinterpret_cast<const double2 *>(W + r0 + 512));
    const double ww0 = wa.x * Wa.x, ww1 = wa.y * Wa.y, ww2 = wb.x * Wb.x, ww3 = wb.y * Wb.y;
    if (WITH_NORM) nrm += ww0 * wa.x + ww1 * wa.y + ww2 * wb.x + ww3 * wb.y;
    const double *vp = V + r0;
    int j0 = 0;
    for (; j0 + KT <= k; j0 += KT) {
      double2 va[KT], vb[KT];
#pragma unroll
      for (int jj = 0; jj < KT; ++jj) {
        const double *p = vp + (int64_t)(j0 + jj) * ld;
        va[jj] = ld_stream(reinterpret_cast<const double2 *>(p));
        vb[jj] = ld_stream(reinterpret_cast<const double2 *>(p + 512));
      }
      double acc[KT];
#pragma unroll
      for (int jj = 0; jj < KT; ++jj)
        acc[jj] = fma(va[jj].x, ww0, fma(va[jj].y, ww1, fma(vb[jj].x, ww2, vb[jj].y * ww3)));
      double s = warp_reduce8(acc, lane);
      if ((lane & 3) == 0) myacc[j0 + (lane >> 2)] += s;
    }
    if (j0 < k) {  // tail tile, guarded loads
      double acc[KT];
#pragma unroll
      for (int jj = 0; jj < KT; ++jj) {
        if (j0 + jj < k) {
          const double *p = vp + (int64_t)(j0 + jj) * ld;
          double2 a = ld_stream(reinterpret_cast<const double2 *>(p));
          double2 b = ld_stream(reinterpret_cast<const double2 *>(p + 512));
          acc[jj] = fma(a.x, ww0, fma(a.y, ww1, fma(b.x, ww2, b.y * ww3)));
        } else {
          acc[jj] = 0.0;
        }
      }
      double s = warp_reduce8(acc, lane);
      if ((lane & 3) == 0 && j0 + (lane >> 2) < k) myacc[j0 + (lane >> 2)] += s;
    }
  }
  if (WITH_NORM) {
    nrm = warp_reduce_sum(nrm);
    if (lane == 0) myacc[k] = nrm;
  }
  __syncthreads();
  const int kout = WITH_NORM ? k + 1 : k;
  for (int j = threadIdx.x; j < kout; j += NT) {
    double s = 0.0;
#pragma unroll
    for (int wp = 0; wp < NT / 32; ++wp) s += accS[wp * kpad + j];
    partial[(size_t)blockIdx.x * pstride + j] = s;
  }
  orth_tail(tail);
}

// MODE 0: w -= V h          (rows [0, nrows))        [+ partial norm over rows < ndot if WITH_NORM]
// MODE 1: out = V y         (k_matmul)
// MODE 2: w = (w - V h) / sqrt(*scale2)   (third sweep of CGS2 with the folded normalisation: scale2 = beta^2 is
//         already known from the second projection, nsb_tail.cuh norm_op 4)
template <int MODE, bool WITH_NORM>
__global__ void __launch_bounds__(NT, 2)
update_kernel(const double *__restrict__ V, int64_t ld, int k, const double *__restrict__ h,
              double *__restrict__ w, const double *__restrict__ W, int64_t nchunks,
              int64_t ndot_chunks, double *__restrict__ partial, const __grid_constant__ OrthTail tail,
              const double *__restrict__ scale2 = nullptr) {
  extern __shared__ double hS[];
  if (tail.skip_flag && *tail.skip_flag == 0) return;   // DGKS: second projection not needed
  for (int j = threadIdx.x; j < k; j += NT) hS[j] = h[j];
  __syncthreads();
  double nrm = 0.0;
  const double inv = (MODE == 2) ? 1.0 / sqrt(*scale2) : 1.0;
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t r0 = chunk * CHUNK + 2 * threadIdx.x;
    const double *vp = V + r0;
    double2 *wpa = reinterpret_cast<double2 *>(w + r0);
    double2 *wpb = reinterpret_cast<double2 *>(w + r0 + 512);
    // w (and W for the norm) are requested before the sweep over the columns, so the
    // read-modify-write at the end of the chunk does not wait on a fresh load
    double2 wa = make_double2(0.0, 0.0), wb = wa, Wa = wa, Wb = wa;
    if (MODE != 1) {
      wa = *wpa;
      wb = *wpb;
      if (WITH_NORM && chunk < ndot_chunks) {
        Wa = ld_stream(reinterpret_cast<const double2 *>(W + r0));
        Wb = ld_stream(reinterpret_cast<const double2 *>(W + r0 + 512));
      }
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j0 = 0;
    for (; j0 + KT <= k; j0 += KT) {
      double2 va[KT], vb[KT];
#pragma unroll
      for (int jj = 0; jj < KT; ++jj) {
        const double *p = vp + (int64_t)(j0 + jj) * ld;
        va[jj] = ld_stream(reinterpret_cast<const double2 *>(p));
        vb[jj] = ld_stream(reinterpret_cast<const double2 *>(p + 512));
      }
#pragma unroll
      for (int jj = 0; jj < KT; ++jj) {
        const double hj = hS[j0 + jj];
        a0 = fma(va[jj].x, hj, a0);
        a1 = fma(va[jj].y, hj, a1);
        a2 = fma(vb[jj].x, hj, a2);
        a3 = fma(vb[jj].y, hj, a3);
      }
    }
    if (j0 < k) {  // tail tile: all remaining columns requested at once (guarded), like the full tiles
      double2 va[KT], vb[KT];
#pragma unroll
      for (int jj = 0; jj < KT; ++jj) {
        va[jj] = vb[jj] = make_double2(0.0, 0.0);
        if (j0 + jj < k) {
          const double *p = vp + (int64_t)(j0 + jj) * ld;
          va[jj] = ld_stream(reinterpret_cast<const double2 *>(p));
          vb[jj] = ld_stream(reinterpret_cast<const double2 *>(p + 512));
        }
      }
#pragma unroll
      for (int jj = 0; jj < KT; ++jj) {
        const double hj = (j0 + jj < k) ? hS[j0 + jj] : 0.0;
        a0 = fma(va[jj].x, hj, a0);
        a1 = fma(va[jj].y, hj, a1);
        a2 = fma(vb[jj].x, hj, a2);
        a3 = fma(vb[jj].y, hj, a3);
      }
    }
    if (MODE == 0) {
      wa.x -= a0; wa.y -= a1; wb.x -= a2; wb.y -= a3;
      *wpa = wa;
      *wpb = wb;
      if (WITH_NORM && chunk < ndot_chunks)
        nrm += Wa.x * wa.x * wa.x + Wa.y * wa.y * wa.y + Wb.x * wb.x * wb.x + Wb.y * wb.y * wb.y;
    } else if (MODE == 2) {
      *wpa = make_double2((wa.x - a0) * inv, (wa.y - a1) * inv);
      *wpb = make_double2((wb.x - a2) * inv, (wb.y - a3) * inv);
    } else {
      *wpa = make_double2(a0, a1);
      *wpb = make_double2(a2, a3);
    }
  }
  if (WITH_NORM) {
    nrm = block_reduce_sum<NT>(nrm);
    if (threadIdx.x == 0) partial[blockIdx.x] = nrm;
    orth_tail(tail);
  }
}

// tail bookkeeping as a kernel of its own: NCCL transport (the all-reduce is a host-enqueued NCCL call
// between the sweep and this) and the NSB_TAIL=0 launch structure
__global__ void __launch_bounds__(NT) orth_post_kernel(const __grid_constant__ OrthTail tail) {
  if (tail.skip_flag && *tail.skip_flag == 0) return;
  orth_post_ops(tail);
}

// w *= 1/sqrt(nrm2[0]);  h_out[k] = sqrt(nrm2[0])   (k_normalize, core/krylov_subspace.f90:75-92)
__global__ void __launch_bounds__(NT)
normalize_kernel(double2 *__restrict__ w, int64_t n2, const double *__restrict__ nrm2,
                 double *__restrict__ hk, double *__restrict__ keep, const int *__restrict__ skip_if_set) {
  if (skip_if_set && *skip_if_set) return;   // DGKS, second projection kept: the third sweep normalised already
  const double beta = sqrt(nrm2[0]);
  const double inv = 1.0 / beta;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (hk) *hk = beta;
    if (keep) *keep = beta;    // survives the next step's H bookkeeping (host-operator pipeline)
  }
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n2; base += stride) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n2) v[u] = w[i];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n2) { v[u].x *= inv; v[u].y *= inv; w[i] = v[u]; }
    }
  }
}

// rows *= 1 / *beta  (the vector a LINEAR host operator returned for an un-normalised input)
__global__ void __launch_bounds__(NT)
scale_rows_kernel(double2 *__restrict__ w, int64_t n2, const double *__restrict__ beta) {
  const double inv = 1.0 / *beta;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 v = w[i];
    v.x *= inv;
    v.y *= inv;
    w[i] = v;
  }
}

__global__ void add_vec_kernel(double *__restrict__ dst, const double *__restrict__ a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += a[i];
}

// ---- panel rotation  V(:,0:k) <- V(:,0:k) Z --------------------------------------------------
// One CTA stages RB rows x k columns in shared memory (column-major, like V), then every thread
// produces a 2-row x 4-column register tile per step; the in-place write is safe because the whole
// row panel is staged before the first store.
template <int RB>
__global__ void __launch_bounds__(NT)
rotate_kernel(double *__restrict__ V, int64_t ld, int k, const double *__restrict__ Z, int ldz,
              int64_t npanels) {
  extern __shared__ double tile[];  // [k][RB]
  constexpr int RP = RB / 2;        // row pairs
  constexpr int CG = NT / RP;       // column groups
  const int rp = threadIdx.x % RP, cg = threadIdx.x / RP;
  for (int64_t panel = blockIdx.x; panel < npanels; panel += gridDim.x) {
    const int64_t r0 = panel * RB;
    __syncthreads();
    for (int idx = threadIdx.x; idx < k * RP; idx += NT) {
      int j = idx / RP, r = idx % RP;
      // plain (coherent) loads: the same kernel overwrites V in place
      reinterpret_cast<double2 *>(tile)[j * RP + r] =
          reinterpret_cast<const double2 *>(V + (int64_t)j * ld + r0)[r];
    }
    __syncthreads();
    for (int c0 = 4 * cg; c0 < k; c0 += 4 * CG) {
      double acc[4][2];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c][0] = acc[c][1] = 0.0;
      const double *z0 = Z + (int64_t)c0 * ldz;
      const int nc = (k - c0) < 4 ? (k - c0) : 4;
      for (int j = 0; j < k; ++j) {
        const double2 v = reinterpret_cast<const double2 *>(tile)[j * RP + rp];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nc) {
            const double z = __ldg(z0 + (int64_t)c * ldz + j);
            acc[c][0] = fma(v.x, z, acc[c][0]);
            acc[c][1] = fma(v.y, z, acc[c][1]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nc)
          reinterpret_cast<double2 *>(V + (int64_t)(c0 + c) * ld + r0)[rp] =
              make_double2(acc[c][0], acc[c][1]);
    }
  }
}

// ---- fused  w -= V h1 ; h2 = V^T (W o w)  (middle step of CGS2) ------------------------------
// The second projection needs the fully updated w, but only row-locally: for a block of RC rows
// the CTA stages V[r0:r0+RC, 0:k] in shared memory ONCE (TMA bulk copies, one per column,
// completing on an mbarrier), forms w' for those rows (pass A) and immediately accumulates the
// block's contribution to h2 from the staged copy (pass B).  V therefore crosses HBM once for the
// two operations: 8 n (k+3) algorithmic bytes instead of 8 n (2k+5).
// Register-tiled variant (used when the panel and at least 64 columns of Z fit in shared memory):
// warp = group of 8 output columns (Z values are warp-uniform broadcasts), lane = 4 rows
// {2l, 2l+1, 64+2l, 64+2l+1} of a 128-row panel (2 rows for 64-row panels), i.e. a 4 x 8 register
// tile per thread: 10 shared loads per 32 FMAs instead of 5 loads per 8.  Z is resident in shared
// memory (all of it when it fits, otherwise 64 columns at a time).
template <int RB>
__global__ void __launch_bounds__(NT, 1)
rotate_tiled_kernel(double *__restrict__ V, int64_t ld, int k, const double *__restrict__ Z, int ldz,
                    int64_t npanels, int ct) {
  extern __shared__ __align__(16) double rsm[];
  constexpr int RP = RB / 2, NR = RB / 64;     // row pairs per panel, row-pair slots per lane
  double *tile = rsm;                          // [k][RB]
  double *Zs = rsm + (size_t)k * RB;           // [ct][k], column-major like Z
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const bool whole = ct >= k;                  // all of Z resident: load it once
  if (whole) {
    for (int t = tid; t < k * k; t += NT) Zs[t] = Z[(int64_t)(t / k) * ldz + (t % k)];
  }
  for (int64_t panel = blockIdx.x; panel < npanels; panel += gridDim.x) {
    const int64_t r0 = panel * RB;
    __syncthreads();
    for (int t = tid; t < k * RP; t += NT) {
      const int j = t / RP, r = t % RP;
      reinterpret_cast<double2 *>(tile)[j * RP + r] =
          reinterpret_cast<const double2 *>(V + (int64_t)j * ld + r0)[r];
    }
    for (int c0 = 0; c0 < k; c0 += ct) {
      const int nc = (k - c0) < ct ? (k - c0) : ct;
      if (!whole) {
        __syncthreads();
        for (int t = tid; t < nc * k; t += NT) Zs[t] = Z[(int64_t)(c0 + t / k) * ldz + (t % k)];
      }
      __syncthreads();
      for (int cg = wq; cg * 8 < nc; cg += NT / 32) {
        double acc[NR][2][8];
#pragma unroll
        for (int a = 0; a < NR; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[a][0][c] = acc[a][1][c] = 0.0;
        const double *zc = Zs + (size_t)(cg * 8) * k;
        const int ncol = (nc - cg * 8) < 8 ? (nc - cg * 8) : 8;
#pragma unroll 4
        for (int j = 0; j < k; ++j) {
          double2 v[NR];
#pragma unroll
          for (int a = 0; a < NR; ++a) v[a] = reinterpret_cast<const double2 *>(tile)[j * RP + lane + 32 * a];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double z = c < ncol ? zc[(size_t)c * k + j] : 0.0;
#pragma unroll
            for (int a = 0; a < NR; ++a) {
              acc[a][0][c] = fma(v[a].x, z, acc[a][0][c]);
              acc[a][1][c] = fma(v[a].y, z, acc[a][1][c]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < ncol) {
            double *col = V + (int64_t)(c0 + cg * 8 + c) * ld + r0;
#pragma unroll
            for (int a = 0; a < NR; ++a)
              reinterpret_cast<double2 *>(col)[lane + 32 * a] = make_double2(acc[a][0][c], acc[a][1][c]);
          }
      }
    }
  }
}


// ---- basis rotation on the fp64 tensor cores ---------------------------------------------------
// V(:,0:k) <- V(:,0:k) Z is the one compute-bound contraction of the path (2 n k^2 flop for 16 n k bytes).
// mma.m8n8k4.f64 with M = rows, K = input columns j, N = output columns c: lane (g, t) supplies
//   A[g][t] = V[r0+g, j0+t]  -- loaded straight from global memory (8 consecutive rows of 4 columns: eight fully
//                               used 32-byte sectors per fragment), no shared-memory staging of V at all;
//   B[t][g] = Z[j0+t, c0+g]  -- from shared memory, column pitch = 4 (mod 16) so the 64-bit loads of a half warp
//                               hit 16 different bank pairs;
// and owns D[g][2t..2t+1] = out[r0+g, c0+2t..].  A warp holds RS = 2 slabs of 8 rows x k columns entirely in
// registers (all A fragments are loaded before the first store, so the in-place update needs no barrier and no
// second buffer), and re-uses every B fragment for both slabs.
__device__ __forceinline__ void dmma884_acc(double2 &d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(d.x), "+d"(d.y)
      : "d"(a), "d"(b));
}

template <int NT8>   // ceil(k / 8) <= NT8
__global__ void __launch_bounds__(NT, 1)
rotate_dmma_kernel(double *__restrict__ V, int64_t ld, int k, const double *__restrict__ Z, int ldz, int zp,
                   int64_t nslab2) {
  constexpr int RS = 2, NJ4 = 2 * NT8, CH = 7;
  extern __shared__ __align__(16) double zs[];   // [NT8 * 8][zp], zero padded
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < NT8 * 8 * zp; i += NT) {
    const int c = i / zp, j = i % zp;
    zs[i] = (c < k && j < k) ? Z[(int64_t)c * ldz + j] : 0.0;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * (NT / 32);
  for (int64_t sl = (int64_t)blockIdx.x * (NT / 32) + warp; sl < nslab2; sl += stride) {
    const int64_t r0 = sl * (8 * RS);
    double a[RS][NJ4];
#pragma unroll
    for (int q = 0; q < RS; ++q)
#pragma unroll
      for (int js = 0; js < NJ4; ++js) {
        const int j = 4 * js + t;
        a[q][js] = j < k ? V[(int64_t)j * ld + r0 + 8 * q + g] : 0.0;   // coherent load: the kernel rewrites V
      }
    // output column tiles in chunks of CH: the accumulators of one chunk live in registers next to the A
    // fragments of the whole slab (all of V's rows were read above, so storing a chunk early is safe)
#pragma unroll
    for (int c0 = 0; c0 < NT8; c0 += CH) {
      double2 acc[RS][CH];
#pragma unroll
      for (int q = 0; q < RS; ++q)
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) acc[q][cc] = make_double2(0.0, 0.0);
#pragma unroll
      for (int js = 0; js < NJ4; ++js) {
        if (4 * js < k) {
#pragma unroll
          for (int cc = 0; cc < CH; ++cc) {
            if (c0 + cc < NT8) {
              const double b = zs[((c0 + cc) * 8 + g) * zp + 4 * js + t];
#pragma unroll
              for (int q = 0; q < RS; ++q) dmma884_acc(acc[q][cc], a[q][js], b);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < RS; ++q)
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) {
          const int c = (c0 + cc) * 8 + 2 * t;
          double *dst = V + r0 + 8 * q + g;
          if (c0 + cc < NT8 && c < k) dst[(int64_t)c * ld] = acc[q][cc].x;
          if (c0 + cc < NT8 && c + 1 < k) dst[(int64_t)(c + 1) * ld] = acc[q][cc].y;
        }
    }
  }
}


// ---- Gram matrix G = V^T W V in ONE pass over V, on the fp64 tensor cores -----------------------
// The orthonormality check of the reference (orthonormality.dat, core/eigensolvers.f90:335-345).  Round 1 ran
// one multi-column dot per column (V crossed HBM k times: 4.1 TB at the benchmark size).  Here a block
// G[P-panel, Q-panel] of up to 104 x 104 entries is formed while the two column panels stream by once:
// mma.m8n8k4.f64 with the contraction over ROWS -- lane (g, t) loads V[r0 + t, c0 + g] straight from global
// memory (the element that is A[g][t] of (W o V)^T for column group c0 and B[t][g] of V for the same group, so
// one load serves both operands); a warp owns a fixed set of 8 x 8 output tiles (compile-time tile list per
// warp), all eight warps of a CTA walk the same 4-row slabs (seven of them hit L1).  Diagonal blocks compute
// the upper triangle only.  Per-CTA tile sums go to a partial buffer and are added in a fixed order.
constexpr int kGramNG = 13;                                   // column groups of 8 per panel (104 columns)

template <bool DIAG>
struct GramTiles {
  static constexpr int count = DIAG ? kGramNG * (kGramNG + 1) / 2 : kGramNG * kGramNG;
  __host__ __device__ static constexpr int row(int t) {
    if (!DIAG) return t / kGramNG;
    int i = 0, rem = t;
    while (rem >= kGramNG - i) { rem -= kGramNG - i; ++i; }
    return i;
  }
  __host__ __device__ static constexpr int col(int t) {
    if (!DIAG) return t % kGramNG;
    int i = 0, rem = t;
    while (rem >= kGramNG - i) { rem -= kGramNG - i; ++i; }
    return i + rem;
  }
};

// compile-time unrolled loops over this warp's tiles: the fragment indices must be constants (register arrays)
template <bool DIAG, int WARP, int Q, int NMINE>
__device__ __forceinline__ void gram_mma_tiles(double2 (&acc)[NMINE], const double (&fa)[kGramNG],
                                               const double (&fb)[kGramNG], double wv) {
  if constexpr (Q < NMINE) {
    constexpr int tile = WARP + (NT / 32) * Q;
    constexpr int ti = GramTiles<DIAG>::row(tile), tj = GramTiles<DIAG>::col(tile);
    dmma884_acc(acc[Q], wv * fa[ti], DIAG ? fa[tj] : fb[tj]);
    gram_mma_tiles<DIAG, WARP, Q + 1, NMINE>(acc, fa, fb, wv);
  }
}

template <bool DIAG, int WARP>
__device__ __forceinline__ void gram_warp(const double *__restrict__ Va, const double *__restrict__ Vb, int64_t ld, int ka,
                                          int kb, const double *__restrict__ W, int64_t nslabs, double *__restrict__ out) {
  using T = GramTiles<DIAG>;
  constexpr int NW = NT / 32, NMINE = (T::count - WARP + NW - 1) / NW;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double2 acc[NMINE];
#pragma unroll
  for (int q = 0; q < NMINE; ++q) acc[q] = make_double2(0.0, 0.0);
  for (int64_t sl = blockIdx.x; sl < nslabs; sl += gridDim.x) {
    const int64_t r = sl * 4 + t;
    const double wv = W[r];
    double fa[kGramNG], fb[kGramNG];
#pragma unroll
    for (int c = 0; c < kGramNG; ++c) {
      const int ca = 8 * c + g;
      fa[c] = ca < ka ? Va[(int64_t)ca * ld + r] : 0.0;
      fb[c] = (!DIAG && ca < kb) ? Vb[(int64_t)ca * ld + r] : 0.0;
    }
    gram_mma_tiles<DIAG, WARP, 0, NMINE>(acc, fa, fb, wv);
  }
#pragma unroll
  for (int q = 0; q < NMINE; ++q) {
    const int tile = WARP + NW * q;
    double *o = out + ((size_t)blockIdx.x * T::count + tile) * 64 + g * 8 + 2 * t;
    o[0] = acc[q].x;
    o[1] = acc[q].y;
  }
}

template <bool DIAG>
__global__ void __launch_bounds__(NT, DIAG ? 2 : 1)
gram_dmma_kernel(const double *__restrict__ Va, const double *__restrict__ Vb, int64_t ld, int ka, int kb,
                 const double *__restrict__ W, int64_t nslabs, double *__restrict__ out) {
  switch (threadIdx.x >> 5) {
    case 0: gram_warp<DIAG, 0>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    case 1: gram_warp<DIAG, 1>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    case 2: gram_warp<DIAG, 2>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    case 3: gram_warp<DIAG, 3>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    case 4: gram_warp<DIAG, 4>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    case 5: gram_warp<DIAG, 5>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    case 6: gram_warp<DIAG, 6>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
    default: gram_warp<DIAG, 7>(Va, Vb, ld, ka, kb, W, nslabs, out); break;
  }
}

// out[e] = sum over CTAs of part[cta * n + e], fixed order
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double *__restrict__ part, int nblk, int64_t n,
                                                          double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[(size_t)b * n + e];
  out[e] = s;
}

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Two shared-memory stages per CTA: while block b is processed, the loads of block b + gridDim.x
// are already in flight.  LOADER 1: cp.async (LDGSTS) 16 B per thread;  LOADER 3: one 2-D TMA tensor
// load per <= 256 columns on an mbarrier per stage (per-column bulk copies and a registers-only variant
// were measured and dropped, profiles/tune_orth_r01.md)
// (box = RC rows x kbox columns, dense [column][row] in shared memory, no swizzle).
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

template <int RC, bool WITH_NORM, int LOADER>
__global__ void __launch_bounds__(NT)
fused_update_dot_kernel(const double *__restrict__ V, int64_t ld, int k, const double *__restrict__ h1,
                        double *__restrict__ w, const double *__restrict__ W, int64_t nblocks,
                        int64_t ndot_blocks, double *__restrict__ partial, int pstride,
                        const __grid_constant__ CUtensorMap tmap, int kbox, int nbox,
                        const __grid_constant__ OrthTail tail) {
  constexpr int RP = RC / 2;      // row pairs per block
  constexpr int CG = NT / RP;     // column groups in pass A
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int kpad = (k + KT) & ~(KT - 1);
  const int kst = LOADER == 3 ? kbox * nbox : k;                // columns held per stage
  double *sV0 = reinterpret_cast<double *>(smem_raw);           // 2 x [kst][RC]
  double *accS = sV0 + 2 * (size_t)kst * RC;                    // [NT/32][kpad]
  double *hS = accS + (NT / 32) * kpad;                         // [k]
  double2 *sP = reinterpret_cast<double2 *>(hS + ((k + 1) & ~1));  // [CG][RP] pass-A partials
  double *sWW = reinterpret_cast<double *>(sP + CG * RP);       // [RC]  W o w'
  uint64_t *bar = reinterpret_cast<uint64_t *>(sWW + RC);       // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (NT / 32) * kpad; i += NT) accS[i] = 0.0;
  for (int j = tid; j < k; j += NT) hS[j] = h1[j];
  if (LOADER == 3 && tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
  }
  __syncthreads();
  const int rp = tid % RP, cg = tid / RP;
  double *myacc = accS + warp * kpad;
  double nrm = 0.0;
  uint32_t phase_bits = 0;  // bit s = parity the next wait on stage s expects

  auto issue = [&](int64_t blk, int stage) {
    double *dst = sV0 + (size_t)stage * kst * RC;
    const double *src = V + blk * RC;
    if (LOADER == 3) {
      if (tid == 0) {
        mbar_expect_tx(bar + stage, (uint32_t)(kst * RC * sizeof(double)));
        for (int b = 0; b < nbox; ++b)
          tma_load_2d(dst + (size_t)b * kbox * RC, &tmap, (int)(blk * RC), b * kbox, bar + stage);
      }
    } else {
      const int npieces = k * RP;   // 16-byte pieces
      for (int p = tid; p < npieces; p += NT) {
        const int j = p / RP, r = p % RP;
        cp_async16(dst + (size_t)j * RC + 2 * r, src + (int64_t)j * ld + 2 * r);
      }
      cp_async_commit();
    }
  };

  int stage = 0;
  int64_t blk = blockIdx.x;
  if (blk < nblocks) issue(blk, 0);
  for (; blk < nblocks; blk += gridDim.x, stage ^= 1) {
    const int64_t r0 = blk * RC;
    const int64_t nxt = blk + gridDim.x;
    if (nxt < nblocks) issue(nxt, stage ^ 1);
    else if (LOADER == 1) cp_async_commit();   // keep the group count uniform
    const double *sV = sV0 + (size_t)stage * kst * RC;
    const bool in_dot = blk < ndot_blocks;
    // every thread loads w / W for its own row pair (the CG copies of a row pair hit L1)
    double2 wv = *reinterpret_cast<const double2 *>(w + r0 + 2 * rp);
    double2 Wv = in_dot ? *reinterpret_cast<const double2 *>(W + r0 + 2 * rp) : make_double2(0.0, 0.0);
    if (LOADER == 3) {
      mbar_wait(bar + stage, (phase_bits >> stage) & 1u);
      phase_bits ^= 1u << stage;
    } else {
      cp_async_wait<1>();
      __syncthreads();
    }
    // pass A: partial row sums over this thread's column group (columns j = cg mod CG)
    double2 acc = make_double2(0.0, 0.0);
    for (int j = cg; j < k; j += CG) {
      const double2 v = reinterpret_cast<const double2 *>(sV + (size_t)j * RC)[rp];
      const double hj = hS[j];
      acc.x = fma(v.x, hj, acc.x);
      acc.y = fma(v.y, hj, acc.y);
    }
    sP[cg * RP + rp] = acc;
    __syncthreads();
    // every thread completes the row sum of its row pair: w' and W o w' stay in registers
    double2 s = sP[rp];
#pragma unroll
    for (int c = 1; c < CG; ++c) {
      const double2 t = sP[c * RP + rp];
      s.x += t.x;
      s.y += t.y;
    }
    wv.x -= s.x;
    wv.y -= s.y;
    if (cg == 0) *reinterpret_cast<double2 *>(w + r0 + 2 * rp) = wv;
    const double2 ww = make_double2(Wv.x * wv.x, Wv.y * wv.y);
    if (WITH_NORM && cg == 0) nrm += ww.x * wv.x + ww.y * wv.y;
    if (in_dot) {
      if (RP >= 32) {
        // pass B, balanced: the same (row pair, column group) mapping as pass A; 8 of the thread's
        // columns at a time are reduced over the warp's 32 row pairs
        for (int jb = 0; cg + CG * KT * jb < k; ++jb) {
          double a8[KT];
#pragma unroll
          for (int c = 0; c < KT; ++c) {
            const int j = cg + CG * (KT * jb + c);
            a8[c] = 0.0;
            if (j < k) {
              const double2 v = reinterpret_cast<const double2 *>(sV + (size_t)j * RC)[rp];
              a8[c] = fma(v.x, ww.x, v.y * ww.y);
            }
          }
          const double red = warp_reduce8(a8, lane);
          const int j = cg + CG * (KT * jb + (lane >> 2));
          if ((lane & 3) == 0 && j < k) myacc[j] += red;
        }
      } else {
        // RC = 32: a warp spans two column groups; go through shared memory for W o w'
        if (cg == 0) reinterpret_cast<double2 *>(sWW)[rp] = ww;
        __syncthreads();
        const double wl = sWW[lane];
        for (int j0 = warp * KT; j0 < k; j0 += (NT / 32) * KT) {
          double a8[KT];
#pragma unroll
          for (int c = 0; c < KT; ++c) a8[c] = (j0 + c < k) ? sV[(size_t)(j0 + c) * RC + lane] * wl : 0.0;
          const double red = warp_reduce8(a8, lane);
          if ((lane & 3) == 0 && j0 + (lane >> 2) < k) myacc[j0 + (lane >> 2)] += red;
        }
      }
    }
    __syncthreads();  // this stage, sP and sWW are reused two iterations / one iteration from now
  }
  if (LOADER == 1) cp_async_wait<0>();
  if (WITH_NORM) {
    nrm = warp_reduce_sum(nrm);
    if (lane == 0) myacc[k] = nrm;   // only warps holding tid < RP contribute non-zero
  }
  __syncthreads();
  const int kout = WITH_NORM ? k + 1 : k;
  for (int j = tid; j < kout; j += NT) {
    double s = 0.0;
#pragma unroll
    for (int wp = 0; wp < NT / 32; ++wp) s += accS[wp * kpad + j];
    partial[(size_t)blockIdx.x * pstride + j] = s;
  }
  orth_tail(tail);
}

// ---- fused update + multidot: 2-D TMA prefetch + register retention ---------------------------
// ncu on the shared-memory variant above showed the L1/shared pipe at 80 %: the staged block is
// written once by TMA and read twice (pass A, pass B).  Here every thread reads its part of the
// block from shared memory ONCE into registers and uses it for both passes, halving the read
// traffic, while the 2-D TMA keeps one block of prefetch in flight per CTA.
// Block = 64 rows; warp c owns the columns j = c (mod 8), lane l the row pair l; NJ = ceil(k/8).
// PRIV: the second projection accumulates in lane-private registers over the CTA's blocks and crosses
// lanes once at the end, instead of one warp_reduce8 per 8 columns and block.
// ALLW: every warp completes the row sums itself (8 shared loads of the other warps' partials, w and W of its
// lane's row pair loaded by all warps, L1 hits for seven of them) instead of waiting for warp 0 to do it:
// one CTA barrier per 64-row block instead of two, and no serial section.  The partial-sum buffer is
// double-buffered by block parity -- a warp can only write buffer b again after passing the barrier of the
// block in between, which every warp reaches after it has read buffer b.
template <int NJ, bool WITH_NORM, bool PRIV, bool ALLW>
__global__ void __launch_bounds__(NT, 2)
fused_tma_reg_kernel(int k, const double *__restrict__ h1, double *__restrict__ w,
                     const double *__restrict__ W, int64_t nblocks, int64_t ndot_blocks,
                     double *__restrict__ partial, int pstride, const __grid_constant__ CUtensorMap tmap,
                     int kbox, int nbox, const __grid_constant__ OrthTail tail) {
  constexpr int RC = 64, NW = NT / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int kst = kbox * nbox;
  double *sV0 = reinterpret_cast<double *>(smem_raw);            // 2 x [kst][RC]
  double *accS = sV0 + 2 * (size_t)kst * RC;                     // [NW * NJ] column sums (+ norm)
  double *hS = accS + NW * NJ + 8;                               // [NW * NJ]
  double2 *sP = reinterpret_cast<double2 *>(hS + NW * NJ);       // [2][NW][32]
  double2 *sWW = sP + 2 * NW * 32;                               // [32]
  uint64_t *bar = reinterpret_cast<uint64_t *>(sWW + 32);        // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < NW * NJ + 8; j += NT) accS[j] = 0.0;
  for (int j = tid; j < NW * NJ; j += NT) hS[j] = j < k ? h1[j] : 0.0;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
  }
  __syncthreads();
  double nrm = 0.0;
  double accp[PRIV ? NJ : 1];
#pragma unroll
  for (int jj = 0; jj < (PRIV ? NJ : 1); ++jj) accp[jj] = 0.0;
  uint32_t phase_bits = 0;
  auto issue = [&](int64_t blk, int stage) {
    if (tid == 0) {
      double *dst = sV0 + (size_t)stage * kst * RC;
      mbar_expect_tx(bar + stage, (uint32_t)(kst * RC * sizeof(double)));
      for (int b = 0; b < nbox; ++b)
        tma_load_2d(dst + (size_t)b * kbox * RC, &tmap, (int)(blk * RC), b * kbox, bar + stage);
    }
  };
  int stage = 0;
  int64_t blk = blockIdx.x;
  if (blk < nblocks) issue(blk, 0);
  for (; blk < nblocks; blk += gridDim.x, stage ^= 1) {
    const int64_t r = blk * RC + 2 * lane;
    const int64_t nxt = blk + gridDim.x;
    if (nxt < nblocks) issue(nxt, stage ^ 1);
    const double *sV = sV0 + (size_t)stage * kst * RC;
    const bool in_dot = blk < ndot_blocks;
    double2 wv = make_double2(0.0, 0.0), Wv = make_double2(0.0, 0.0);
    if (ALLW) {
      wv = *reinterpret_cast<const double2 *>(w + r);
      if (in_dot) Wv = __ldg(reinterpret_cast<const double2 *>(W + r));
    } else if (warp == 0) {
      wv = *reinterpret_cast<const double2 *>(w + r);
      if (in_dot) Wv = ld_stream(reinterpret_cast<const double2 *>(W + r));
    }
    mbar_wait(bar + stage, (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;
    double2 v[NJ];
    double2 a = make_double2(0.0, 0.0);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = warp + NW * jj;   // j >= k: hS is 0 and the staged column is either valid data or OOB zero fill
      v[jj] = j < kst ? reinterpret_cast<const double2 *>(sV + (size_t)j * RC)[lane] : make_double2(0.0, 0.0);
      const double hj = hS[j];
      a.x = fma(v[jj].x, hj, a.x);
      a.y = fma(v[jj].y, hj, a.y);
    }
    double2 *sPb = sP + (ALLW ? stage * NW * 32 : 0);   // stage alternates with the block parity of this CTA
    sPb[warp * 32 + lane] = a;
    __syncthreads();
    double2 ww;
    if (ALLW) {
      double2 s = sPb[lane];
#pragma unroll
      for (int c = 1; c < NW; ++c) {
        const double2 t = sPb[c * 32 + lane];
        s.x += t.x;
        s.y += t.y;
      }
      wv.x -= s.x;
      wv.y -= s.y;
      ww = make_double2(Wv.x * wv.x, Wv.y * wv.y);
      if (warp == 0) {
        *reinterpret_cast<double2 *>(w + r) = wv;
        if (WITH_NORM) nrm += ww.x * wv.x + ww.y * wv.y;
      }
    } else {
      if (warp == 0) {
        double2 s = sPb[lane];
#pragma unroll
        for (int c = 1; c < NW; ++c) {
          const double2 t = sPb[c * 32 + lane];
          s.x += t.x;
          s.y += t.y;
        }
        wv.x -= s.x;
        wv.y -= s.y;
        *reinterpret_cast<double2 *>(w + r) = wv;
        const double2 w2 = make_double2(Wv.x * wv.x, Wv.y * wv.y);
        sWW[lane] = w2;
        if (WITH_NORM) nrm += w2.x * wv.x + w2.y * wv.y;
      }
      __syncthreads();
      ww = sWW[lane];
    }
    if (in_dot) {
      if constexpr (PRIV) {
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) accp[jj] = fma(v[jj].x, ww.x, fma(v[jj].y, ww.y, accp[jj]));
      } else {
#pragma unroll
        for (int j0 = 0; j0 < NJ; j0 += KT) {
          double a8[KT];
#pragma unroll
          for (int c = 0; c < KT; ++c) {
            const int jj = (j0 + c < NJ) ? j0 + c : 0;
            a8[c] = (j0 + c < NJ) ? fma(v[jj].x, ww.x, v[jj].y * ww.y) : 0.0;
          }
          const double red = warp_reduce8(a8, lane);
          const int jj = j0 + (lane >> 2);
          if ((lane & 3) == 0 && jj < NJ) accS[warp + NW * jj] += red;   // column warp + 8 jj, owned by this warp
        }
      }
    }
  }
  if constexpr (PRIV) {
#pragma unroll
    for (int j0 = 0; j0 < NJ; j0 += KT) {
      double a8[KT];
#pragma unroll
      for (int c = 0; c < KT; ++c) a8[c] = (j0 + c < NJ) ? accp[(j0 + c < NJ) ? j0 + c : 0] : 0.0;
      const double red = warp_reduce8(a8, lane);
      const int jj = j0 + (lane >> 2);
      if ((lane & 3) == 0 && jj < NJ) accS[warp + NW * jj] = red;
    }
  }
  if (WITH_NORM) {
    nrm = warp_reduce_sum(nrm);
    if (tid == 0) accS[NW * NJ] = nrm;
  }
  __syncthreads();
  for (int j = tid; j < k; j += NT) partial[(size_t)blockIdx.x * pstride + j] = accS[j];
  if (WITH_NORM && tid == 0) partial[(size_t)blockIdx.x * pstride + k] = accS[NW * NJ];
  orth_tail(tail);
}

inline size_t fused_tma_reg_smem(int k, int nj) {
  int kbox, nbox;
  kbox = 0; nbox = (k + 255) / 256; kbox = (k + nbox - 1) / nbox;
  const size_t kst = (size_t)kbox * nbox;
  return sizeof(double) * (2 * kst * 64 + 2 * 8 * nj + 8) + sizeof(double2) * (2 * 8 * 32 + 32) + 32;
}

// Large-k variant (209 <= k <= 440): 32-row blocks, 16 warps (512 threads), lane = one row,
// warp c owns the columns j = c (mod 16); otherwise identical to fused_tma_reg_kernel.
template <int NJ, bool WITH_NORM>
__global__ void __launch_bounds__(512, 1)
fused_tma_reg32_kernel(int k, const double *__restrict__ h1, double *__restrict__ w,
                       const double *__restrict__ W, int64_t nblocks, int64_t ndot_blocks,
                       double *__restrict__ partial, int pstride, const __grid_constant__ CUtensorMap tmap,
                       int kbox, int nbox, const __grid_constant__ OrthTail tail) {
  constexpr int RC = 32, NW = 16, NTH = 512;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int kst = kbox * nbox;
  double *sV0 = reinterpret_cast<double *>(smem_raw);            // 2 x [kst][RC]
  double *accS = sV0 + 2 * (size_t)kst * RC;                     // [NW * NJ] (+ norm)
  double *hS = accS + NW * NJ + 8;                               // [NW * NJ]
  double *sP = hS + NW * NJ;                                     // [NW][32]
  double *sWW = sP + NW * 32;                                    // [32]
  uint64_t *bar = reinterpret_cast<uint64_t *>(sWW + 32);        // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < NW * NJ + 8; j += NTH) accS[j] = 0.0;
  for (int j = tid; j < NW * NJ; j += NTH) hS[j] = j < k ? h1[j] : 0.0;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
  }
  __syncthreads();
  double nrm = 0.0;
  uint32_t phase_bits = 0;
  auto issue = [&](int64_t blk, int stage) {
    if (tid == 0) {
      double *dst = sV0 + (size_t)stage * kst * RC;
      mbar_expect_tx(bar + stage, (uint32_t)(kst * RC * sizeof(double)));
      for (int b = 0; b < nbox; ++b)
        tma_load_2d(dst + (size_t)b * kbox * RC, &tmap, (int)(blk * RC), b * kbox, bar + stage);
    }
  };
  int stage = 0;
  int64_t blk = blockIdx.x;
  if (blk < nblocks) issue(blk, 0);
  for (; blk < nblocks; blk += gridDim.x, stage ^= 1) {
    const int64_t r = blk * RC + lane;
    const int64_t nxt = blk + gridDim.x;
    if (nxt < nblocks) issue(nxt, stage ^ 1);
    const double *sV = sV0 + (size_t)stage * kst * RC;
    const bool in_dot = blk < ndot_blocks;
    double wv = 0.0, Wv = 0.0;
    if (warp == 0) {
      wv = w[r];
      if (in_dot) Wv = ld_stream1(W + r);
    }
    mbar_wait(bar + stage, (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;
    double v[NJ];
    double a = 0.0;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = warp + NW * jj;
      v[jj] = j < kst ? sV[(size_t)j * RC + lane] : 0.0;
      a = fma(v[jj], hS[j], a);
    }
    sP[warp * 32 + lane] = a;
    __syncthreads();
    if (warp == 0) {
      double s = sP[lane];
#pragma unroll
      for (int c = 1; c < NW; ++c) s += sP[c * 32 + lane];
      wv -= s;
      w[r] = wv;
      const double ww = Wv * wv;
      sWW[lane] = ww;
      if (WITH_NORM) nrm += ww * wv;
    }
    __syncthreads();
    if (in_dot) {
      const double ww = sWW[lane];
#pragma unroll
      for (int j0 = 0; j0 < NJ; j0 += KT) {
        double a8[KT];
#pragma unroll
        for (int c = 0; c < KT; ++c) a8[c] = (j0 + c < NJ) ? v[(j0 + c < NJ) ? j0 + c : 0] * ww : 0.0;
        const double red = warp_reduce8(a8, lane);
        const int jj = j0 + (lane >> 2);
        if ((lane & 3) == 0 && jj < NJ) accS[warp + NW * jj] += red;
      }
    }
  }
  if (WITH_NORM) {
    nrm = warp_reduce_sum(nrm);
    if (tid == 0) accS[NW * NJ] = nrm;
  }
  __syncthreads();
  for (int j = tid; j < k; j += NTH) partial[(size_t)blockIdx.x * pstride + j] = accS[j];
  if (WITH_NORM && tid == 0) partial[(size_t)blockIdx.x * pstride + k] = accS[NW * NJ];
  orth_tail(tail);
}

inline size_t fused_tma_reg32_smem(int k, int nj) {
  const int nbox = (k + 255) / 256, kbox = (k + nbox - 1) / nbox;
  const size_t kst = (size_t)kbox * nbox;
  return sizeof(double) * (2 * kst * 32 + 2 * 16 * nj + 8 + 16 * 32 + 32) + 32;
}

inline void fused_boxes(int k, int *kbox, int *nbox) {
  *nbox = (k + 255) / 256;
  *kbox = (k + *nbox - 1) / *nbox;
}

inline size_t fused_smem_bytes(int rc, int k) {
  const int kpad = (k + KT) & ~(KT - 1);
  const int cg = NT / (rc / 2);
  int kbox, nbox;
  fused_boxes(k, &kbox, &nbox);
  const size_t kst = (size_t)kbox * nbox;   // >= k
  return sizeof(double) * (2 * kst * rc + (NT / 32) * kpad + ((k + 1) & ~1) + 2 * cg * (rc / 2) + rc) + 32;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

// 2-D tensor map over the first k columns of V: dim0 = rows (contiguous), dim1 = columns
int make_basis_tmap(CUtensorMap *tm, const double *V, int64_t ld, int k, int rc, int kbox) {
  static encode_tiled_fn enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult q;
    void *fn = nullptr;
    NSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    NSB_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    enc = (encode_tiled_fn)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)k};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)rc, (cuuint32_t)kbox};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)V, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (ld=%lld k=%d box=%dx%d)", (int)r,
              (long long)ld, k, rc, kbox);
  return NSB_OK;
}

// rows per block: the largest of 128/64/32 whose two stages let two CTAs share an SM, else the
// largest that fits alone; 0 = does not fit
inline int fused_rows(int k) {
  for (int rc : {128, 64, 32})
    if (fused_smem_bytes(rc, k) <= 113 * 1024) return rc;
  for (int rc : {128, 64, 32})
    if (fused_smem_bytes(rc, k) <= 226 * 1024) return rc;
  return 0;
}

inline int persistent_grid(nsb_context_t ctx, int64_t nchunks) {
  int64_t g = (int64_t)ctx->num_sms * 2;
  return (int)(nchunks < g ? nchunks : g);
}

// What the last CTA of a sweep does with its reduced vector (nsb_tail.cuh)
struct TailSpec {
  double *out = nullptr;
  double *hsum = nullptr;
  int hsum_op = 0, norm_op = 0;
  const int *skip_flag = nullptr;
  double *passes_out = nullptr;
};

OrthTail make_tail(nsb_context_t ctx, const TailSpec &sp, int k, int kout, int pstride) {
  OrthTail t;
  t.partial = ctx->partial_d;
  t.pstride = pstride;
  t.ticket = ctx->tail ? ctx->ticket_d : nullptr;
  t.out = sp.out;
  t.hsum = sp.hsum;
  t.scal = ctx->hvec_d + 3 * (kMaxK + 8);
  t.flag = ctx->flag_d;
  t.passes_out = sp.passes_out;
  t.skip_flag = sp.skip_flag;
  t.k = k;
  t.kout = kout;
  t.hsum_op = sp.hsum_op;
  t.norm_op = sp.norm_op;
  t.eta2 = ctx->dgks_eta2;
  t.comm.P = ctx->nranks;
  t.comm.rank = ctx->rank;
  t.comm.seq = ctx->seq_d;
  t.comm.err = ctx->dev_err_d;
  for (int r = 0; r < nsb_context_s::kMaxPeers; ++r) t.comm.mail.p[r] = r < ctx->nranks ? ctx->peer_mail[r] : nullptr;
  t.exchange = (ctx->tail && ctx->nranks > 1 && ctx->p2p && kout <= ARN) ? 1 : 0;
  return t;
}

// Whatever of {reduce, all-reduce, bookkeeping} the kernel's own tail did not do.
int finish_tail(nsb_context_t ctx, const OrthTail &t, int grid) {
  if (t.kout == 0) return NSB_OK;
  bool post = false;
  if (!t.ticket) {   // NSB_TAIL=0: the launch structure of round 1
    ProfScope ps(ctx, PC_SMALL, 8.0 * grid * t.kout);
    reduce_partials_kernel<<<(t.kout * 32 + 255) / 256, 256, 0, ctx->stream>>>(t.partial, grid, t.pstride, t.kout,
                                                                              t.out, 0, nullptr);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, t.out, t.kout));
    post = true;
  } else if (ctx->nranks > 1 && !t.exchange) {   // NCCL transport
    NSB_CHECK(allreduce_sum_d(ctx, t.out, t.kout));
    post = true;
  }
  if (post && (t.hsum_op || t.norm_op)) {
    orth_post_kernel<<<1, NT, 0, ctx->stream>>>(t);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
  }
  return NSB_OK;
}

// Row-range variant used by the pipelined upload: rows [r0, r1) only, partial sums written from
// partial row `prow` on; returns the number of partial rows produced (no second-stage reduction).
int launch_multidot_rows(nsb_context_t ctx, const double *V, int64_t ld, int k, const double *w, const double *W,
                         int64_t r0, int64_t r1, int prow, int *nrows_out) {
  const int64_t nchunks = (r1 - r0) / CHUNK;
  const int grid = persistent_grid(ctx, nchunks);
  const int kpad = (k + KT) & ~(KT - 1);
  const size_t smem = sizeof(double) * (NT / 32) * kpad;
  const int pstride = kMaxK + 8;
  cudaSetDevice(ctx->device);
  ProfScope ps(ctx, PC_MULTIDOT, 8.0 * (double)(r1 - r0) * (k + 2));
  multidot_kernel<false><<<grid, NT, smem, ctx->stream>>>(V + r0, ld, k, w + r0, W + r0, nchunks,
                                                         ctx->partial_d + (size_t)prow * pstride, pstride, OrthTail());
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  *nrows_out = grid;
  return NSB_OK;
}

int launch_multidot(nsb_context_t ctx, const double *V, int64_t ld, int k, const double *w,
                    const double *W, int64_t ndot, bool with_norm, const TailSpec &sp, int64_t nalg = -1) {
  const int64_t nchunks = ndot / CHUNK;
  const int grid = persistent_grid(ctx, nchunks);
  const int kpad = (k + KT) & ~(KT - 1);
  const size_t smem = sizeof(double) * (NT / 32) * kpad;
  const int pstride = kMaxK + 8;
  NSB_CHECK(ensure_partial(ctx, grid));
  cudaSetDevice(ctx->device);
  const int kout = with_norm ? k + 1 : k;
  const OrthTail tail = make_tail(ctx, sp, k, kout, pstride);
  {
    // algorithmic bytes: V once (k columns), w and W once
    ProfScope ps(ctx, PC_MULTIDOT, 8.0 * (double)(nalg >= 0 ? nalg : ndot) * (k + 2));
    if (with_norm)
      multidot_kernel<true><<<grid, NT, smem, ctx->stream>>>(V, ld, k, w, W, nchunks, ctx->partial_d, pstride, tail);
    else
      multidot_kernel<false><<<grid, NT, smem, ctx->stream>>>(V, ld, k, w, W, nchunks, ctx->partial_d, pstride, tail);
  }
  ctx->launches += 1;
  NSB_CUDA(cudaGetLastError());
  return finish_tail(ctx, tail, grid);
}

template <int MODE>
int launch_update(nsb_context_t ctx, const double *V, int64_t ld, int k, const double *h_d, double *w,
                  const double *W, int64_t nrows, int64_t ndot, bool with_norm, const TailSpec &sp,
                  int64_t nalg = -1, int64_t nalg_dot = -1, const double *scale2_d = nullptr) {
  const int64_t nchunks = nrows / CHUNK;
  const int grid = persistent_grid(ctx, nchunks);
  const size_t smem = sizeof(double) * (k > 0 ? k : 1);
  cudaSetDevice(ctx->device);
  // algorithmic bytes: V once, w read + written (MODE 0) or written (MODE 1), W once with the norm
  const double na = (double)(nalg >= 0 ? nalg : nrows), nd = (double)(nalg_dot >= 0 ? nalg_dot : ndot);
  const double bytes = 8.0 * (na * (k + (MODE != 1 ? 2 : 1)) + (with_norm ? nd : 0.0));
  // the norm is the only thing this kernel reduces: one partial per CTA, coefficient count 0
  const OrthTail tail = make_tail(ctx, sp, 0, with_norm ? 1 : 0, 1);
  {
    ProfScope ps(ctx, MODE != 1 ? PC_UPDATE : PC_GEMV, bytes);
    if (with_norm)
      update_kernel<MODE, true><<<grid, NT, smem, ctx->stream>>>(V, ld, k, h_d, w, W, nchunks, ndot / CHUNK,
                                                                ctx->partial_d, tail, scale2_d);
    else
      update_kernel<MODE, false><<<grid, NT, smem, ctx->stream>>>(V, ld, k, h_d, w, W, nchunks, ndot / CHUNK,
                                                                 ctx->partial_d, tail, scale2_d);
  }
  ctx->launches += 1;
  NSB_CUDA(cudaGetLastError());
  return finish_tail(ctx, tail, grid);
}

int launch_fused(nsb_context_t ctx, const double *V, int64_t ld, int k, const double *h1_d, double *w,
                 const double *W, int64_t nrows, int64_t ndot, bool with_norm, const TailSpec &sp, int64_t nalg,
                 int64_t nalg_dot) {
  const int pstride = kMaxK + 8;
  const int kout = with_norm ? k + 1 : k;
  const OrthTail tail = make_tail(ctx, sp, k, kout, pstride);
  int grid = 0;
  cudaSetDevice(ctx->device);
  if (ctx->fused_loader == 3 && k >= ctx->fused_reg_min_k && k <= 8 * 26 &&
      fused_tma_reg_smem(k, (k + 7) / 8) <= 226 * 1024) {
    // 2-D TMA prefetch + register retention, 64-row blocks
    const int64_t nblocks = nrows / 64, ndot_blocks = ndot / 64;
    const int njr = (k + 7) / 8;
    const int nj = njr <= 4 ? 4 : njr <= 7 ? 7 : njr <= 10 ? 10 : njr <= 13 ? 13 : njr <= 16 ? 16 : njr <= 20 ? 20 : 26;
    const size_t smem = fused_tma_reg_smem(k, nj);
    int kbox, nbox;
    fused_boxes(k, &kbox, &nbox);
    CUtensorMap tmap;
    NSB_CHECK(make_basis_tmap(&tmap, V, ld, k, 64, kbox));
    const int64_t g = (int64_t)ctx->num_sms * 2;
    grid = (int)(nblocks < g ? nblocks : g);
    NSB_CHECK(ensure_partial(ctx, grid));
    ProfScope ps(ctx, PC_FUSED, 8.0 * ((double)nalg * (k + 2) + (double)nalg_dot));
#define LAUNCH_TR4(NJ, NORM, PRIV, ALLW)                                                                \
  do {                                                                                                  \
    NSB_CUDA(cudaFuncSetAttribute(fused_tma_reg_kernel<NJ, NORM, PRIV, ALLW>,                           \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    fused_tma_reg_kernel<NJ, NORM, PRIV, ALLW><<<grid, NT, smem, ctx->stream>>>(                        \
        k, h1_d, w, W, nblocks, ndot_blocks, ctx->partial_d, pstride, tmap, kbox, nbox, tail);          \
  } while (0)
#define LAUNCH_TR3(NJ, NORM, PRIV)                                                                      \
  do {                                                                                                  \
    if (ctx->fused_allwarps) LAUNCH_TR4(NJ, NORM, PRIV, true);                                          \
    else LAUNCH_TR4(NJ, NORM, PRIV, false);                                                             \
  } while (0)
#define LAUNCH_TR2(NJ, NORM)                                                                            \
  do {                                                                                                  \
    if ((NJ) <= 13 && ctx->fused_priv) LAUNCH_TR3(NJ, NORM, ((NJ) <= 13));                              \
    else LAUNCH_TR3(NJ, NORM, false);                                                                   \
  } while (0)
#define LAUNCH_TR(NJ) do { if (with_norm) LAUNCH_TR2(NJ, true); else LAUNCH_TR2(NJ, false); } while (0)
    switch (nj) {
      case 4: LAUNCH_TR(4); break;
      case 7: LAUNCH_TR(7); break;
      case 10: LAUNCH_TR(10); break;
      case 13: LAUNCH_TR(13); break;
      case 16: LAUNCH_TR(16); break;
      case 20: LAUNCH_TR(20); break;
      default: LAUNCH_TR(26); break;
    }
#undef LAUNCH_TR
#undef LAUNCH_TR2
#undef LAUNCH_TR3
#undef LAUNCH_TR4
  } else if (ctx->fused_loader == 3 && k > 8 * 26 && k <= 16 * 28 &&
             fused_tma_reg32_smem(k, (k + 15) / 16) <= 226 * 1024) {
    const int64_t nblocks = nrows / 32, ndot_blocks = ndot / 32;
    const int njr = (k + 15) / 16;
    const int nj = njr <= 16 ? 16 : njr <= 20 ? 20 : njr <= 24 ? 24 : 28;
    const size_t smem = fused_tma_reg32_smem(k, nj);
    int kbox, nbox;
    fused_boxes(k, &kbox, &nbox);
    CUtensorMap tmap;
    NSB_CHECK(make_basis_tmap(&tmap, V, ld, k, 32, kbox));
    grid = (int)(nblocks < ctx->num_sms ? nblocks : ctx->num_sms);
    NSB_CHECK(ensure_partial(ctx, grid));
    ProfScope ps(ctx, PC_FUSED, 8.0 * ((double)nalg * (k + 2) + (double)nalg_dot));
#define LAUNCH_T32B(NJ, NORM)                                                                            \
  do {                                                                                                   \
    NSB_CUDA(cudaFuncSetAttribute(fused_tma_reg32_kernel<NJ, NORM>,                                      \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    fused_tma_reg32_kernel<NJ, NORM><<<grid, 512, smem, ctx->stream>>>(k, h1_d, w, W, nblocks, ndot_blocks, \
                                                                      ctx->partial_d, pstride, tmap, kbox, nbox, tail); \
  } while (0)
#define LAUNCH_T32(NJ) do { if (with_norm) LAUNCH_T32B(NJ, true); else LAUNCH_T32B(NJ, false); } while (0)
    switch (nj) {
      case 16: LAUNCH_T32(16); break;
      case 20: LAUNCH_T32(20); break;
      case 24: LAUNCH_T32(24); break;
      default: LAUNCH_T32(28); break;
    }
#undef LAUNCH_T32
#undef LAUNCH_T32B
  } else {
    int rc = fused_rows(k);
    NSB_REQUIRE(rc != 0, "fused update+dot: k=%d does not fit in shared memory", k);
    if (ctx->fused_rc && fused_smem_bytes(ctx->fused_rc, k) <= 226 * 1024) rc = ctx->fused_rc;
    const int loader = ctx->fused_loader;
    const size_t smem = fused_smem_bytes(rc, k);
    int kbox, nbox;
    fused_boxes(k, &kbox, &nbox);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (loader == 3) NSB_CHECK(make_basis_tmap(&tmap, V, ld, k, rc, kbox));
    const int64_t nblocks = nrows / rc, ndot_blocks = ndot / rc;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    const int64_t g = (int64_t)ctx->num_sms * per_sm;
    grid = (int)(nblocks < g ? nblocks : g);
    NSB_CHECK(ensure_partial(ctx, grid));
    // algorithmic bytes: V once, w read + written, W once
    ProfScope ps(ctx, PC_FUSED, 8.0 * ((double)nalg * (k + 2) + (double)nalg_dot));
#define LAUNCH_FUSED2(RC, NORM, LD)                                                                     \
  do {                                                                                                  \
    NSB_CUDA(cudaFuncSetAttribute(fused_update_dot_kernel<RC, NORM, LD>,                                \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    fused_update_dot_kernel<RC, NORM, LD><<<grid, NT, smem, ctx->stream>>>(                             \
        V, ld, k, h1_d, w, W, nblocks, ndot_blocks, ctx->partial_d, pstride, tmap, kbox, nbox, tail);   \
  } while (0)
#define LAUNCH_FUSED(RC, NORM)                                                \
  do {                                                                        \
    if (loader == 1) LAUNCH_FUSED2(RC, NORM, 1);                              \
    else LAUNCH_FUSED2(RC, NORM, 3);                                          \
  } while (0)
    if (rc == 128) { if (with_norm) LAUNCH_FUSED(128, true); else LAUNCH_FUSED(128, false); }
    else if (rc == 64) { if (with_norm) LAUNCH_FUSED(64, true); else LAUNCH_FUSED(64, false); }
    else { if (with_norm) LAUNCH_FUSED(32, true); else LAUNCH_FUSED(32, false); }
#undef LAUNCH_FUSED
#undef LAUNCH_FUSED2
  }
  ctx->launches += 1;
  NSB_CUDA(cudaGetLastError());
  return finish_tail(ctx, tail, grid);
}

int launch_normalize(nsb_context_t ctx, double *w, int64_t ld, const double *nrm2_d, double *hk_d,
                     const int *skip_if_set = nullptr) {
  int64_t n2 = ld / 2;
  int64_t want = (n2 + 1023) / 1024;
  int64_t cap = (int64_t)ctx->num_sms * 16;
  int grid = (int)(want < cap ? want : cap);
  ProfScope ps(ctx, PC_NORMALIZE, 16.0 * ld);
  normalize_kernel<<<grid, NT, 0, ctx->stream>>>(reinterpret_cast<double2 *>(w), n2, nrm2_d, hk_d,
                                                 ctx->hvec_d + 3 * (kMaxK + 8) + 3, skip_if_set);
  ctx->launches += 1;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

}  // namespace

namespace nsb {
// Host vector -> column col_w in row chunks on the copy stream; the first projection
// h1 = V_k^T (W o w) of every chunk starts as soon as that chunk has landed, so the H2D transfer
// of f (the output of the host's matvec) overlaps the first sweep over the basis.
int upload_multidot_pipelined(nsb_basis_t B, int col_w, const double *const *fields, double time, int k,
                              const double *scale_by_inv_d) {
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = L->ctx;
  cudaSetDevice(ctx->device);
  // Row chunks of decreasing size (1/2, 1/4, ... , 1/64, 1/64 of the column): the first projection of a chunk starts
  // when the chunk has landed, so what stays exposed after the transfer is the sweep over the LAST chunk only
  // (equal eighths left 1/8 of a sweep, 0.37 ms at the benchmark size, behind the H2D copy).
  const int C = 7;
  static const int64_t kUpCut[C + 1] = {0, 32, 48, 56, 60, 62, 63, 64};   // chunk boundaries in 64ths
  const int S = kMaxK + 8;
  double *w = B->col(col_w);
  double *h1 = ctx->hvec_d;
  NSB_CHECK(ensure_partial(ctx, (int64_t)(C + 1) * ctx->num_sms * 2));
  while ((int)ctx->chunk_ev.size() < C + 1) {
    cudaEvent_t e;
    NSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->chunk_ev.push_back(e);
  }
  // the copy stream starts after everything already queued on the compute stream (column reuse)
  NSB_CUDA(cudaEventRecord(ctx->chunk_ev[C], ctx->stream));
  NSB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[C], 0));
  ctx->hpin[3 * S + 8] = time;
  const int64_t nchk = L->ndot / CHUNK;              // 1024-row chunks of the inner-product prefix
  int prow = 0;
  for (int c = 0; c < C; ++c) {
    const int64_t r0 = (nchk * kUpCut[c] / 64) * CHUNK;
    const int64_t r1 = (c == C - 1) ? L->ld : (nchk * kUpCut[c + 1] / 64) * CHUNK;   // last chunk: rest of the column
    for (int f = 0; f < L->nfields; ++f) {
      const int64_t a = std::max<int64_t>(L->off[f], r0), b = std::min<int64_t>(L->off[f] + L->len[f], r1);
      if (b <= a) continue;
      if (fields[f])
        NSB_CUDA(cudaMemcpyAsync(w + a, fields[f] + (a - L->off[f]), sizeof(double) * (b - a), cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
      else
        NSB_CUDA(cudaMemsetAsync(w + a, 0, sizeof(double) * (b - a), ctx->copy_stream));
    }
    if (L->time_row >= r0 && L->time_row < r1)
      NSB_CUDA(cudaMemcpyAsync(w + L->time_row, ctx->hpin + 3 * S + 8, sizeof(double), cudaMemcpyHostToDevice,
                               ctx->copy_stream));
    NSB_CUDA(cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
    NSB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
    if (scale_by_inv_d && r1 > r0) {   // f = M(w'') / beta: the host saw the un-normalised vector
      const int64_t n2 = (r1 - r0) / 2;
      const int g = (int)std::min<int64_t>((n2 + NT - 1) / NT, (int64_t)ctx->num_sms * 8);
      ProfScope ps(ctx, PC_BLAS1, 16.0 * (double)(r1 - r0));
      scale_rows_kernel<<<g, NT, 0, ctx->stream>>>(reinterpret_cast<double2 *>(w + r0), n2, scale_by_inv_d);
      ctx->launches++;
    }
    const int64_t d1 = std::min<int64_t>(r1, L->ndot);
    if (k > 0 && d1 > r0) {
      int nr = 0;
      NSB_CHECK(launch_multidot_rows(ctx, B->v_d, L->ld, k, w, L->w_d, r0, d1, prow, &nr));
      prow += nr;
    }
  }
  if (k > 0) {
    ProfScope ps(ctx, PC_SMALL, 8.0 * prow * k);
    reduce_partials_kernel<<<(k * 32 + 255) / 256, 256, 0, ctx->stream>>>(ctx->partial_d, prow, S, k, h1, 0, nullptr);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, h1, k));
    ctx->h1_ready_k = k;      // the next CGS2 orthonormalisation of this column skips its first sweep
    ctx->h1_ready_col = w;
  }
  return NSB_OK;
}

int weighted_multidot(nsb_basis_t b, int k, const double *w_col_d, double *h_d) {
  nsb_layout_t L = b->lay;
  TailSpec sp;
  sp.out = h_d;
  return launch_multidot(L->ctx, b->v_d, L->ld, k, w_col_d, L->w_d, L->ndot, false, sp);
}

// Enqueue the whole orthonormalisation (core/krylov_decomposition.f90:150-186) on the context stream; nothing
// here waits for the device.  On completion hsum = ctx->hvec_d[2*(kMaxK+8) ...] holds h[0..k] (H(1:k+1,k)) and,
// in DGKS mode, hsum[k+1] the number of projection passes taken.
int orth_enqueue(nsb_basis_t B, int k, int col_w, int mode, const StreamOut *so) {
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = L->ctx;
  const int S = kMaxK + 8;
  double *h1 = ctx->hvec_d, *h2 = ctx->hvec_d + S, *hsum = ctx->hvec_d + 2 * S,
         *scal = ctx->hvec_d + 3 * S;
  double *w = B->col(col_w);
  const double *V = B->v_d;
  cudaSetDevice(ctx->device);
  struct ClearH1 {   // the pipelined h1 is valid for exactly one orthonormalisation
    nsb_context_t c;
    ~ClearH1() { c->h1_ready_k = -1; c->h1_ready_col = nullptr; }
  } clear_h1{ctx};
  if (k == 0 || mode == NSB_ORTH_MGS2_REF) { ctx->h1_ready_k = -1; }
  TailSpec norm_only;            // |w|^2 -> scal[0], what normalize_kernel divides by
  norm_only.out = scal + 2;
  norm_only.norm_op = 1;
  if (k == 0) {
    // nothing to project out: k_normalize only
    NSB_CHECK(launch_multidot(ctx, w, L->ld, 0, w, L->w_d, L->ndot, true, norm_only));
  } else if (mode == NSB_ORTH_MGS2_REF) {
    // literal core/krylov_decomposition.f90:155-180: one column at a time, two sweeps
    NSB_CUDA(cudaMemsetAsync(hsum, 0, sizeof(double) * (k + 1), ctx->stream));
    TailSpec none;
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < k; ++i) {
        TailSpec sp;
        sp.out = h1;
        sp.hsum = hsum + i;     // H(i,k) = alpha (:166), H(i,k) += alpha (:178)
        sp.hsum_op = 2;
        NSB_CHECK(launch_multidot(ctx, B->col(i), L->ld, 1, w, L->w_d, L->ndot, false, sp));
        NSB_CHECK(launch_update<0>(ctx, B->col(i), L->ld, 1, h1, w, L->w_d, L->ld, L->ndot, false, none));
      }
    NSB_CHECK(launch_multidot(ctx, w, L->ld, 0, w, L->w_d, L->ndot, true, norm_only));
  } else {
    const bool dgks = (mode == NSB_ORTH_DGKS);
    const bool fused = !ctx->no_fused &&
                       (fused_rows(k) != 0 || (ctx->fused_loader == 3 && k > 8 * 26 && k <= 16 * 28 &&
                                               fused_tma_reg32_smem(k, (k + 15) / 16) <= 226 * 1024));
    // pass 1 (DGKS: the norm of the incoming w rides along for the test); already done chunk by chunk
    // during the upload when the vector came from the host (upload_multidot_pipelined)
    const bool have_h1 = !dgks && ctx->h1_ready_k == k && ctx->h1_ready_col == w;
    // CGS2 with the normalisation folded into the third sweep (nsb_tail.cuh, norm_op 4): |w'|^2 rides along with
    // h2, beta^2 = |w'|^2 - |h2|^2, and the last sweep writes the normalised vector -- two reductions per step
    // instead of three and no normalize_kernel.  Not with the streamed download (it sends w'' before beta is
    // needed) and not in DGKS mode (whether a second projection happens is decided by the same reduction).
    const bool fold = fused && !dgks && !so && ctx->fold_norm;
    // DGKS: the same fold when the second projection is kept; when it is dropped beta = |w'| and only the
    // normalisation pass runs (the sweep and the pass read the decision flag and one of them exits at once)
    const bool fold_dgks = fused && dgks && !so && ctx->fold_norm;
    if (!have_h1) {
      TailSpec sp;
      sp.out = h1;
      sp.hsum = hsum;
      sp.hsum_op = 1;                 // H(1:k,k) = h1
      sp.norm_op = dgks ? 2 : 0;      // scal[1] = |w|^2
      NSB_CHECK(launch_multidot(ctx, V, L->ld, k, w, L->w_d, L->ndot, dgks, sp, L->ndof_dot + 1));
    } else {
      NSB_CUDA(cudaMemcpyAsync(hsum, h1, sizeof(double) * k, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    const int *skip = dgks ? ctx->flag_d : nullptr;   // second projection only when the device-side test asks for it
    if (fused) {
      // w -= V h1 and h2 = V^T W w in one sweep over V; its tail adds h2 into H -- in DGKS mode only if
      // |w'| < |w| / sqrt 2, the decision being taken by the kernel's last CTA from h2[k] = |w'|^2
      TailSpec sp;
      sp.out = h2;
      sp.hsum = hsum;
      sp.hsum_op = 2;
      sp.norm_op = dgks ? (fold_dgks ? 5 : 3) : fold ? 4 : 0;
      sp.passes_out = dgks ? hsum + k + 1 : nullptr;
      NSB_CHECK(launch_fused(ctx, V, L->ld, k, h1, w, L->w_d, L->ld, L->ndot, dgks || fold, sp, L->nact,
                             L->ndof_dot + 1));
      if (fold) {
        NSB_CHECK(launch_update<2>(ctx, V, L->ld, k, h2, w, L->w_d, L->ld, L->ndot, false, TailSpec(), L->nact, 0, scal));
        return NSB_OK;
      }
      if (fold_dgks) {
        TailSpec t2;
        t2.skip_flag = skip;
        NSB_CHECK(launch_update<2>(ctx, V, L->ld, k, h2, w, L->w_d, L->ld, L->ndot, false, t2, L->nact, 0, scal));
        NSB_CHECK(launch_normalize(ctx, w, L->ld, scal, hsum + k, ctx->flag_d));
        return NSB_OK;
      }
    } else {
      TailSpec su;                    // DGKS: |w'|^2 and the decision ride on the first update
      if (dgks) {
        su.out = scal + 2;
        su.norm_op = 3;
        su.passes_out = hsum + k + 1;
      }
      NSB_CHECK(launch_update<0>(ctx, V, L->ld, k, h1, w, L->w_d, L->ld, L->ndot, dgks, su, L->nact, L->ndof_dot + 1));
      TailSpec sp;
      sp.out = h2;
      sp.hsum = hsum;
      sp.hsum_op = 2;
      sp.skip_flag = skip;
      NSB_CHECK(launch_multidot(ctx, V, L->ld, k, w, L->w_d, L->ndot, false, sp, L->ndof_dot + 1));
    }
    TailSpec sn = norm_only;
    sn.skip_flag = skip;
    if (so && !dgks) {
      // Third sweep in row chunks; every finished chunk of w'' (un-normalised) starts its way to the host at
      // once on the copy stream, so the download of the next Krylov vector overlaps this sweep and the
      // normalisation instead of following them.
      // chunks of increasing size (1/64, 1/64, 1/32, ... , 1/2): the download starts after 1/64 of the sweep
      const int C = 7;
      static const int64_t kDnCut[C + 1] = {0, 1, 2, 4, 8, 16, 32, 64};
      NSB_CHECK(ensure_partial(ctx, (int64_t)(C + 1) * ctx->num_sms * 2));
      while ((int)ctx->chunk_ev.size() < C + 1) {
        cudaEvent_t e;
        NSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(e);
      }
      const int64_t nchk = L->ld / CHUNK;
      int prow = 0;
      for (int c = 0; c < C; ++c) {
        const int64_t r0 = (nchk * kDnCut[c] / 64) * CHUNK, r1 = (nchk * kDnCut[c + 1] / 64) * CHUNK;
        if (r1 <= r0) continue;
        const int64_t rows_chunks = (r1 - r0) / CHUNK;
        const int64_t ndot_rel = std::max<int64_t>(0, std::min<int64_t>(rows_chunks, (L->ndot - r0) / CHUNK));
        const int grid = persistent_grid(ctx, rows_chunks);
        {
          ProfScope ps(ctx, PC_UPDATE, 8.0 * ((double)(r1 - r0) * (k + 2) + (double)ndot_rel * CHUNK));
          update_kernel<0, true><<<grid, NT, sizeof(double) * k, ctx->stream>>>(
              V + r0, L->ld, k, h2, w + r0, L->w_d + r0, rows_chunks, ndot_rel, ctx->partial_d + prow, OrthTail());
        }
        ctx->launches++;
        prow += grid;
        NSB_CUDA(cudaEventRecord(ctx->chunk_ev[c], ctx->stream));
        NSB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[c], 0));
        for (int f = 0; f < L->nfields; ++f) {
          const int64_t a = std::max<int64_t>(L->off[f], r0), b = std::min<int64_t>(L->off[f] + L->len[f], r1);
          if (b <= a || !so->fields[f]) continue;
          NSB_CUDA(cudaMemcpyAsync(so->fields[f] + (a - L->off[f]), w + a, sizeof(double) * (b - a), cudaMemcpyDeviceToHost,
                                   ctx->copy_stream));
        }
        if (so->time && L->time_row >= r0 && L->time_row < r1)
          NSB_CUDA(cudaMemcpyAsync(so->time, w + L->time_row, sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
      }
      NSB_CUDA(cudaGetLastError());
      OrthTail t = make_tail(ctx, sn, 0, 1, 1);
      t.ticket = nullptr;                       // the chunk launches reduce nothing themselves
      t.exchange = 0;
      NSB_CHECK(finish_tail(ctx, t, prow));
      // the normalisation rescales w in place: it must not start before the last chunk has left for the host
      NSB_CUDA(cudaEventRecord(ctx->chunk_ev[C], ctx->copy_stream));
      NSB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[C], 0));
    } else {
      NSB_CHECK(launch_update<0>(ctx, V, L->ld, k, h2, w, L->w_d, L->ld, L->ndot, true, sn, L->nact, L->ndof_dot + 1));
    }
  }
  NSB_CHECK(launch_normalize(ctx, w, L->ld, scal, hsum + k));
  return NSB_OK;
}
}  // namespace nsb

static int orth_args_ok(nsb_basis_t B, int k, int col_w, int mode, const void *h) {
  NSB_REQUIRE(B && h, "nsb_orthonormalize: NULL argument");
  NSB_REQUIRE(k >= 0 && k <= kMaxK && k <= B->ncols, "nsb_orthonormalize: k=%d out of range", k);
  NSB_REQUIRE(col_w >= k && col_w < B->ncols, "nsb_orthonormalize: col_w=%d must be >= k and < ncols", col_w);
  NSB_REQUIRE(mode >= 0 && mode <= 2, "nsb_orthonormalize: unknown mode %d", mode);
  return NSB_OK;
}

// h_pinned receives h[0..k] and, in DGKS mode, the number of passes in h_pinned[k+1] (so it needs k+2 doubles).
namespace nsb {
// nsb_orthonormalize with the un-normalised result streamed to the host while the last sweep runs (host-operator
// Arnoldi loop, linear operators); synchronises the compute stream, not the copy stream.
int orthonormalize_stream_out(nsb_basis_t B, int k, int col_w, int mode, double *h, const StreamOut *so) {
  nsb_context_t ctx = B->lay->ctx;
  NSB_CHECK(orth_enqueue(B, k, col_w, mode, so));
  NSB_CUDA(cudaMemcpyAsync(ctx->hpin, ctx->hvec_d + 2 * (kMaxK + 8), sizeof(double) * (k + 1), cudaMemcpyDeviceToHost,
                           ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  NSB_CHECK(check_dev_err(ctx));
  memcpy(h, ctx->hpin, sizeof(double) * (k + 1));
  for (int i = 0; i <= k; ++i)
    if (std::isnan(h[i])) {
      set_error("NaN detected in dot product");
      return NSB_ENAN;
    }
  return NSB_OK;
}
}  // namespace nsb

extern "C" int nsb_orthonormalize_async(nsb_basis_t B, int k, int col_w, int mode, double *h_pinned) {
  NSB_CHECK(orth_args_ok(B, k, col_w, mode, h_pinned));
  nsb_context_t ctx = B->lay->ctx;
  NSB_CHECK(orth_enqueue(B, k, col_w, mode, nullptr));
  NSB_CUDA(cudaMemcpyAsync(h_pinned, ctx->hvec_d + 2 * (kMaxK + 8),
                           sizeof(double) * (k + 1 + (mode == NSB_ORTH_DGKS && k > 0 ? 1 : 0)),
                           cudaMemcpyDeviceToHost, ctx->stream));
  return NSB_OK;
}

extern "C" int nsb_orthonormalize(nsb_basis_t B, int k, int col_w, int mode, double *h, int *passes) {
  NSB_CHECK(orth_args_ok(B, k, col_w, mode, h));
  nsb_context_t ctx = B->lay->ctx;
  NSB_CHECK(nsb_orthonormalize_async(B, k, col_w, mode, ctx->hpin));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  NSB_CHECK(check_dev_err(ctx));
  memcpy(h, ctx->hpin, sizeof(double) * (k + 1));
  if (passes) *passes = k == 0 ? 0 : (mode == NSB_ORTH_DGKS ? (int)ctx->hpin[k + 1] : 2);
  for (int i = 0; i <= k; ++i)
    if (std::isnan(h[i])) {
      set_error("NaN detected in dot product");
      return NSB_ENAN;
    }
  return NSB_OK;
}

extern "C" int nsb_set_dgks_eta(nsb_context_t ctx, double eta) {
  NSB_REQUIRE(ctx && eta > 0.0 && eta < 1.0, "nsb_set_dgks_eta: eta must lie in (0, 1)");
  ctx->dgks_eta2 = eta * eta;
  clear_step_graphs(ctx);   // the threshold is a kernel argument of the captured steps
  return NSB_OK;
}

extern "C" int nsb_basis_gram(nsb_basis_t B, int k, double *G, int ldg) {
  NSB_REQUIRE(B && G && k >= 1 && k <= B->ncols && k <= kMaxK && ldg >= k, "nsb_basis_gram: bad argument");
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = L->ctx;
  cudaSetDevice(ctx->device);
  constexpr int PW = 8 * kGramNG;                               // panel width, 104 columns
  const int np = (k + PW - 1) / PW;
  const int64_t nslabs = L->ndot / 4;                           // the inner product covers rows [0, ndot)
  const int grid = (int)std::min<int64_t>(nslabs, (int64_t)ctx->num_sms * 2);
  const int64_t nfull = (int64_t)GramTiles<false>::count * 64;  // entries of one (off-diagonal) block
  // partial tile sums of every CTA + the reduced block, in a scratch buffer kept by the context
  const size_t need = (size_t)grid * nfull + nfull;
  if (ctx->gram_elems < need) {
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->gram_d) cudaFree(ctx->gram_d);
    ctx->gram_d = nullptr;
    NSB_CUDA(cudaMalloc(&ctx->gram_d, sizeof(double) * need));
    ctx->gram_elems = need;
  }
  double *part = ctx->gram_d, *blk = ctx->gram_d + (size_t)grid * nfull;
  std::vector<double> host((size_t)nfull);
  for (int P = 0; P < np; ++P)
    for (int Q = P; Q < np; ++Q) {
      const bool diag = P == Q;
      const int ka = std::min(PW, k - P * PW), kb = std::min(PW, k - Q * PW);
      const int ntiles = diag ? GramTiles<true>::count : GramTiles<false>::count;
      const int64_t n = (int64_t)ntiles * 64;
      const int gridv = diag ? grid : (int)std::min<int64_t>(nslabs, (int64_t)ctx->num_sms);   // 1 CTA / SM off the diagonal
      {
        // algorithmic bytes: the two column panels and W once
        ProfScope ps(ctx, PC_MULTIDOT, 8.0 * (double)(L->ndof_dot + 1) * (diag ? ka + 1 : ka + kb + 1));
        if (diag)
          gram_dmma_kernel<true><<<gridv, NT, 0, ctx->stream>>>(B->col(P * PW), B->col(Q * PW), L->ld, ka, kb, L->w_d, nslabs, part);
        else
          gram_dmma_kernel<false><<<gridv, NT, 0, ctx->stream>>>(B->col(P * PW), B->col(Q * PW), L->ld, ka, kb, L->w_d, nslabs, part);
      }
      gram_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(part, gridv, n, blk);
      ctx->launches += 2;
      NSB_CUDA(cudaGetLastError());
      if (ctx->nranks > 1)
        for (int64_t o = 0; o < n; o += kMaxK) NSB_CHECK(allreduce_sum_d(ctx, blk + o, (int)std::min<int64_t>(kMaxK, n - o)));
      NSB_CUDA(cudaMemcpyAsync(host.data(), blk, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      NSB_CUDA(cudaStreamSynchronize(ctx->stream));
      NSB_CHECK(check_dev_err(ctx));
      for (int tIdx = 0; tIdx < ntiles; ++tIdx) {
        const int ti = diag ? GramTiles<true>::row(tIdx) : GramTiles<false>::row(tIdx);
        const int tj = diag ? GramTiles<true>::col(tIdx) : GramTiles<false>::col(tIdx);
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b) {
            const int i = P * PW + ti * 8 + a, j = Q * PW + tj * 8 + b;
            if (i >= k || j >= k || ti * 8 + a >= ka || tj * 8 + b >= kb) continue;
            const double v = host[(size_t)tIdx * 64 + a * 8 + b];
            if (diag && ti == tj && b < a) continue;            // the strict lower part of a diagonal tile is mirrored
            G[(size_t)j * ldg + i] = v;
            G[(size_t)i * ldg + j] = v;
          }
      }
    }
  return NSB_OK;
}

extern "C" int nsb_basis_gemv(nsb_basis_t B, int k, const double *y, nsb_basis_t bout, int cout) {
  NSB_REQUIRE(B && y && bout, "nsb_basis_gemv: NULL argument");
  NSB_REQUIRE(k >= 1 && k <= B->ncols && k <= kMaxK, "nsb_basis_gemv: k=%d out of range", k);
  NSB_REQUIRE(cout >= 0 && cout < bout->ncols, "nsb_basis_gemv: column out of range");
  NSB_REQUIRE(bout->lay == B->lay, "nsb_basis_gemv: different layouts");
  NSB_REQUIRE(!(bout == B && cout < k), "nsb_basis_gemv: output column aliases an input column");
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = L->ctx;
  double *y_d = ctx->hvec_d + (kMaxK + 8);
  memcpy(ctx->hpin + (kMaxK + 8), y, sizeof(double) * k);
  NSB_CUDA(cudaMemcpyAsync(y_d, ctx->hpin + (kMaxK + 8), sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream));
  NSB_CHECK(launch_update<1>(ctx, B->v_d, L->ld, k, y_d, bout->col(cout), L->w_d, L->ld, L->ndot, false, TailSpec()));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));  // hpin is reused by the next call
  return NSB_OK;
}

extern "C" int nsb_basis_rotate(nsb_basis_t B, int k, const double *Z, int ldz, int rotate_time) {
  NSB_REQUIRE(B && Z, "nsb_basis_rotate: NULL argument");
  NSB_REQUIRE(k >= 1 && k <= B->ncols && ldz >= k, "nsb_basis_rotate: k=%d ldz=%d", k, ldz);
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = L->ctx;
  cudaSetDevice(ctx->device);
  // Z and the saved %time row live in a scratch buffer kept by the context (no allocation per restart)
  const size_t need = (size_t)k * k + (size_t)k;
  if (ctx->rot_elems < need) {
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->rot_d) cudaFree(ctx->rot_d);
    ctx->rot_d = nullptr;
    NSB_CUDA(cudaMalloc(&ctx->rot_d, sizeof(double) * need));
    ctx->rot_elems = need;
  }
  double *Z_d = ctx->rot_d, *tsave = rotate_time ? nullptr : ctx->rot_d + (size_t)k * k;
  NSB_CUDA(cudaMemcpy2DAsync(Z_d, sizeof(double) * k, Z, sizeof(double) * ldz, sizeof(double) * k, k,
                             cudaMemcpyHostToDevice, ctx->stream));
  if (!rotate_time)
    NSB_CUDA(cudaMemcpy2DAsync(tsave, sizeof(double), B->v_d + L->time_row, sizeof(double) * L->ld,
                               sizeof(double), k, cudaMemcpyDeviceToDevice, ctx->stream));
  // fp64 tensor-core kernel: k <= 104, whole slabs of 16 rows (ld is a multiple of 1024)
  if (k <= 104 && ctx->rotate_dmma && !ctx->rotate_simple) {
    const int nt8 = (k + 7) / 8 <= 7 ? 7 : 13;
    int zp = 8 * nt8;
    zp += (4 - zp % 16 + 16) % 16;                      // column pitch = 4 (mod 16): conflict-free B loads
    const size_t smem = sizeof(double) * (size_t)nt8 * 8 * zp;
    const int64_t nslab2 = L->ld / 16;
    const int64_t want = (nslab2 + NT / 32 - 1) / (NT / 32);
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * 2);
    {
      ProfScope ps(ctx, PC_ROTATE, 16.0 * (double)L->nact * k);
      if (nt8 == 7) {
        NSB_CUDA(cudaFuncSetAttribute(rotate_dmma_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_dmma_kernel<7><<<grid, NT, smem, ctx->stream>>>(B->v_d, L->ld, k, Z_d, k, zp, nslab2);
      } else {
        NSB_CUDA(cudaFuncSetAttribute(rotate_dmma_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_dmma_kernel<13><<<grid, NT, smem, ctx->stream>>>(B->v_d, L->ld, k, Z_d, k, zp, nslab2);
      }
    }
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    if (!rotate_time)
      NSB_CUDA(cudaMemcpy2DAsync(B->v_d + L->time_row, sizeof(double) * L->ld, tsave, sizeof(double),
                                 sizeof(double), k, cudaMemcpyDeviceToDevice, ctx->stream));
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));     // Z is the caller's host memory
    return NSB_OK;
  }
  // register-tiled kernel when the panel plus (part of) Z fit in shared memory
  {
    const size_t lim = 226 * 1024;
    int trb = 0, ct = 0;
    for (int rbc : {128, 64}) {
      const size_t panel = (size_t)rbc * k * sizeof(double);
      if (panel + (size_t)k * k * sizeof(double) <= lim) { trb = rbc; ct = k; break; }
      if (panel + (size_t)64 * k * sizeof(double) <= lim) { trb = rbc; ct = 64; break; }
    }
    if (trb && !ctx->rotate_simple) {
      const size_t smem = ((size_t)trb * k + (size_t)ct * k) * sizeof(double);
      const int64_t npanels = L->ld / trb;
      const int grid = (int)(npanels < ctx->num_sms ? npanels : ctx->num_sms);
      {
        ProfScope ps(ctx, PC_ROTATE, 16.0 * (double)L->nact * k);
        if (trb == 128) {
          NSB_CUDA(cudaFuncSetAttribute(rotate_tiled_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          rotate_tiled_kernel<128><<<grid, NT, smem, ctx->stream>>>(B->v_d, L->ld, k, Z_d, k, npanels, ct);
        } else {
          NSB_CUDA(cudaFuncSetAttribute(rotate_tiled_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          rotate_tiled_kernel<64><<<grid, NT, smem, ctx->stream>>>(B->v_d, L->ld, k, Z_d, k, npanels, ct);
        }
      }
      ctx->launches++;
      NSB_CUDA(cudaGetLastError());
      if (!rotate_time)
        NSB_CUDA(cudaMemcpy2DAsync(B->v_d + L->time_row, sizeof(double) * L->ld, tsave, sizeof(double),
                                   sizeof(double), k, cudaMemcpyDeviceToDevice, ctx->stream));
      NSB_CUDA(cudaStreamSynchronize(ctx->stream));
      return NSB_OK;
    }
  }
  // fallback: rows per panel chosen so that the staged panel fits in shared memory
  int rb = 128;
  while (rb > 32 && (size_t)rb * k * sizeof(double) > 200 * 1024) rb >>= 1;
  NSB_REQUIRE((size_t)rb * k * sizeof(double) <= 227 * 1024, "nsb_basis_rotate: k=%d too large", k);
  size_t smem = (size_t)rb * k * sizeof(double);
  int64_t npanels = L->ld / rb;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int64_t g = (int64_t)ctx->num_sms * per_sm;
  int grid = (int)(npanels < g ? npanels : g);
#define LAUNCH_ROT(RB)                                                                              \
  do {                                                                                              \
    NSB_CUDA(cudaFuncSetAttribute(rotate_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                  (int)smem));                                                      \
    ProfScope ps(ctx, PC_ROTATE, 16.0 * (double)L->nact * k);                                       \
    rotate_kernel<RB><<<grid, NT, smem, ctx->stream>>>(B->v_d, L->ld, k, Z_d, k, npanels);          \
  } while (0)
  if (rb == 128) LAUNCH_ROT(128);
  else if (rb == 64) LAUNCH_ROT(64);
  else LAUNCH_ROT(32);
#undef LAUNCH_ROT
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  if (!rotate_time)
    NSB_CUDA(cudaMemcpy2DAsync(B->v_d + L->time_row, sizeof(double) * L->ld, tsave, sizeof(double),
                               sizeof(double), k, cudaMemcpyDeviceToDevice, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  return NSB_OK;
}
