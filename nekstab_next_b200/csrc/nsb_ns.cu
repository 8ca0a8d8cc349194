// Pressure-coupled perturbation step on the device: what nek_advance does between the nopcopy calls of
// exponential_prop%matvec (core/linear_operators.f90:225-274), for Nek5000's P_N - P_N-2 formulation
// [UPSTREAM-RECALL perturb.f: perturbv -> advabp, makextp, makebdfp, cresvipp + ophinv, incomprp;  navier1.f: opdiv /
// multd, opgradt / cdtp, cdabdtp, opbinv, ortho, uzawa;  coef.f: geom2 / map12].  Nek5000 is not vendored with the
// reference: the kernels are checked against a CPU restatement of those routines held by the test suite, which is pinned
// by independent mathematics only (tests/test_gpu_ns.py); parity unpinned.
//
//   velocity : lx1 = N + 1 Gauss-Lobatto-Legendre points per direction, C0, fields 0..dim-1 of a column
//   pressure : lx2 = lx1 - 2 Gauss-Legendre points, element-local, field `dim` of a column (nekStab keeps the
//              pressure in the Krylov vector and out of the inner product, core/nek_vectors.f90:20-31)
//
//   D  (opdiv)   (D u)_q   = w_q sum_b sum_a (J dr_a/dx_b)_q (du_b/dr_a)_q         one CTA per element, sum-factorised
//   D^T (opgradt) exact transpose of D, element-local                              the same stages backwards
//   E  (cdabdtp)  D B^-1 D^T  with B^-1 = binvm1 mask QQ^T                          opgradt -> gather-scatter -> opdiv
//   esolve        E dp = g by CG preconditioned with 1 / bm2, CG scalars on the device, polled every 8 iterations
//   step          bf = EXT(-B C(v)) + BDF ; v* = H^-1 QQ^T (bf + D^T p*) ; E dp = -(bd0/dt) D v* ;
//                 v = v* + (dt/bd0) B^-1 D^T dp ; p = p* + dp
#include <cmath>
#include <cstring>
#include <vector>

#include "nsb_internal.h"
#include "nsb_quadrature.h"
#include "nsb_device.cuh"

#include <cooperative_groups.h>

using namespace nsb;
namespace cg = cooperative_groups;

namespace {

constexpr int NT_NS = 256;
constexpr int MAXQ = 4;   // pressure points of an element per thread (lx2^dim <= MAXQ * NT_NS)

// out[(o * no + O) * inner + x] (+)= sum_l M(O, l) in[(o * nl + l) * inner + x],  M(O, l) = TR ? M[l * no + O] : M[O * nl + l]
// -- one tensor-product stage along the middle axis of an [outer][nl][inner] array, whole CTA.
template <bool TR, bool ACC>
__device__ __forceinline__ void stage(double *__restrict__ out, const double *__restrict__ in, const double *__restrict__ M,
                                      int outer, int no, int nl, int inner) {
  const int total = outer * no * inner;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int x = t % inner, O = (t / inner) % no, o = t / (inner * no);
    const double *src = in + (size_t)o * nl * inner + x;
    double s = 0.0;
#pragma unroll 8
    for (int l = 0; l < nl; ++l) s = fma(TR ? M[l * no + O] : M[O * nl + l], src[(size_t)l * inner], s);
    if (ACC) out[t] += s;
    else out[t] = s;
  }
}

struct NsDims {
  int dim, l1, l2;
  int n1e, n2e;   // points per element on the two meshes
};

// shared-memory plan (doubles): I12 | D12 | u [n1e] | A, B [lz1 ly1 l2 each] | AA, AD, BA [lz1 l2 l2 each] | G_r, G_s, G_t [n2e each]
__host__ __device__ inline int ns_smem_doubles(int dim, int l1, int l2) {
  const int lz1 = dim == 3 ? l1 : 1;
  const int n1e = l1 * l1 * lz1, n2e = l2 * l2 * (dim == 3 ? l2 : 1);
  return 2 * l2 * l1 + n1e + 2 * lz1 * l1 * l2 + 3 * lz1 * l2 * l2 + 3 * n2e;
}

// (D u): vel = dim equally spaced fields (stride fs) -> pressure array.  out = scale * D u; optional partial sums of
// out * dotw per CTA (pressure CG: (E p, p)).
template <int TL1, int TDIM>
__global__ void __launch_bounds__(NT_NS)
opdiv_kernel(NsDims d, const double *__restrict__ vel, int64_t fs, const double *__restrict__ rx2, int64_t n2,
             const double *__restrict__ I12g, const double *__restrict__ D12g, double scale, double *__restrict__ out,
             const double *__restrict__ dotw, double *__restrict__ dot_partial, const int *__restrict__ done) {
  if (done && *done) return;
  extern __shared__ double sm[];
  const int l1 = TL1 ? TL1 : d.l1, l2 = TL1 ? TL1 - 2 : d.l2, dim = TDIM ? TDIM : d.dim, lz1 = dim == 3 ? l1 : 1;
  if (TL1) {
    d.n1e = l1 * l1 * lz1;
    d.n2e = l2 * l2 * (dim == 3 ? l2 : 1);
  }
  double *I12 = sm, *D12 = I12 + l2 * l1, *u = D12 + l2 * l1, *A = u + d.n1e, *B = A + lz1 * l1 * l2,
         *AA = B + lz1 * l1 * l2, *AD = AA + lz1 * l2 * l2, *BA = AD + lz1 * l2 * l2;
  for (int t = threadIdx.x; t < l2 * l1; t += blockDim.x) {
    I12[t] = I12g[t];
    D12[t] = D12g[t];
  }
  const int64_t e = blockIdx.x;
  double acc[MAXQ];
#pragma unroll
  for (int m = 0; m < MAXQ; ++m) acc[m] = 0.0;
  for (int b = 0; b < dim; ++b) {
    __syncthreads();
    const double *ub = vel + (int64_t)b * fs + e * d.n1e;
    for (int t = threadIdx.x; t < d.n1e; t += blockDim.x) u[t] = ub[t];
    __syncthreads();
    stage<false, false>(A, u, I12, lz1 * l1, l2, l1, 1);        // r: interpolate / differentiate
    stage<false, false>(B, u, D12, lz1 * l1, l2, l1, 1);
    __syncthreads();
    stage<false, false>(AA, A, I12, lz1, l2, l1, l2);           // s
    stage<false, false>(AD, A, D12, lz1, l2, l1, l2);
    stage<false, false>(BA, B, I12, lz1, l2, l1, l2);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MAXQ; ++m) {
      const int q = threadIdx.x + m * NT_NS;
      if (q >= d.n2e) break;
      double ur, us, ut = 0.0;
      if (dim == 3) {
        const int K = q / (l2 * l2), ji = q % (l2 * l2);
        ur = us = 0.0;
        for (int k = 0; k < l1; ++k) {
          const double ik = I12[K * l1 + k], dk = D12[K * l1 + k];
          ur = fma(ik, BA[k * l2 * l2 + ji], ur);
          us = fma(ik, AD[k * l2 * l2 + ji], us);
          ut = fma(dk, AA[k * l2 * l2 + ji], ut);
        }
      } else {
        ur = BA[q];
        us = AD[q];
      }
      const int64_t g = e * d.n2e + q;
      double s = rx2[(int64_t)(0 * dim + b) * n2 + g] * ur + rx2[(int64_t)(1 * dim + b) * n2 + g] * us;
      if (dim == 3) s += rx2[(int64_t)(2 * dim + b) * n2 + g] * ut;
      acc[m] += s;
    }
  }
  double dot = 0.0;
#pragma unroll
  for (int m = 0; m < MAXQ; ++m) {
    const int q = threadIdx.x + m * NT_NS;
    if (q >= d.n2e) break;
    const int64_t g = e * d.n2e + q;
    const double v = scale * acc[m];
    out[g] = v;
    if (dotw) dot += v * dotw[g];
  }
  if (dot_partial) {
    dot = block_reduce_sum<NT_NS>(dot);
    if (threadIdx.x == 0) dot_partial[blockIdx.x] = dot;
  }
}

// (D^T p), element-local.  Epilogue per point x of velocity field b:
//   epi 0 : out = alpha * uin + beta * w                        (all points; uin may be null)
//   epi 1 : element-interior points  out = alpha * uin + beta * bmask * w,  element-boundary points  out = w
//           (the gather-scatter with the same alpha / beta / bmask finishes those: B^-1 D^T p and v* + c B^-1 D^T dp)
template <int TL1, int TDIM>
__global__ void __launch_bounds__(NT_NS)
opgradt_kernel(NsDims d, const double *__restrict__ p, const double *__restrict__ rx2, int64_t n2,
               const double *__restrict__ I12g, const double *__restrict__ D12g, double *out, int64_t fs,
               int epi, const double *uin, double alpha, double beta, const double *__restrict__ bmask,
               const int *__restrict__ done) {
  if (done && *done) return;
  extern __shared__ double sm[];
  const int l1 = TL1 ? TL1 : d.l1, l2 = TL1 ? TL1 - 2 : d.l2, dim = TDIM ? TDIM : d.dim, lz1 = dim == 3 ? l1 : 1;
  if (TL1) {
    d.n1e = l1 * l1 * lz1;
    d.n2e = l2 * l2 * (dim == 3 ? l2 : 1);
  }
  double *I12 = sm, *D12 = I12 + l2 * l1, *w = D12 + l2 * l1, *A = w + d.n1e, *B = A + lz1 * l1 * l2,
         *AA = B + lz1 * l1 * l2, *AD = AA + lz1 * l2 * l2, *BA = AD + lz1 * l2 * l2, *Gr = BA + lz1 * l2 * l2,
         *Gs = Gr + d.n2e, *Gt = Gs + d.n2e;
  for (int t = threadIdx.x; t < l2 * l1; t += blockDim.x) {
    I12[t] = I12g[t];
    D12[t] = D12g[t];
  }
  const int64_t e = blockIdx.x;
  for (int b = 0; b < dim; ++b) {
    __syncthreads();
    for (int q = threadIdx.x; q < d.n2e; q += blockDim.x) {
      const int64_t g = e * d.n2e + q;
      const double pq = p[g];
      Gr[q] = rx2[(int64_t)(0 * dim + b) * n2 + g] * pq;
      Gs[q] = rx2[(int64_t)(1 * dim + b) * n2 + g] * pq;
      if (dim == 3) Gt[q] = rx2[(int64_t)(2 * dim + b) * n2 + g] * pq;
    }
    __syncthreads();
    if (dim == 3) {   // t, transposed: [K][J I] -> [k][J I]
      stage<true, false>(BA, Gr, I12, 1, l1, l2, l2 * l2);
      stage<true, false>(AD, Gs, I12, 1, l1, l2, l2 * l2);
      stage<true, false>(AA, Gt, D12, 1, l1, l2, l2 * l2);
      __syncthreads();
    }
    const double *ba = dim == 3 ? BA : Gr, *ad = dim == 3 ? AD : Gs;
    stage<true, false>(B, ba, I12, lz1, l1, l2, l2);            // s, transposed: [k][J][I] -> [k][j][I]
    stage<true, false>(A, ad, D12, lz1, l1, l2, l2);
    if (dim == 3) {
      __syncthreads();
      stage<true, true>(A, AA, I12, lz1, l1, l2, l2);
    }
    __syncthreads();
    stage<true, false>(w, B, D12, lz1 * l1, l1, l2, 1);         // r, transposed: [k j][I] -> [k j][i]
    __syncthreads();
    stage<true, true>(w, A, I12, lz1 * l1, l1, l2, 1);
    __syncthreads();
    double *ob = out + (int64_t)b * fs + e * d.n1e;
    const double *ui = uin ? uin + (int64_t)b * fs + e * d.n1e : nullptr;
    for (int t = threadIdx.x; t < d.n1e; t += blockDim.x) {
      const double v = w[t];
      if (epi == 0) {
        ob[t] = (ui ? alpha * ui[t] : 0.0) + beta * v;
      } else {
        const int i = t % l1, j = (t / l1) % l1, k = t / (l1 * l1);
        const bool bnd = i == 0 || i == l1 - 1 || j == 0 || j == l1 - 1 || (dim == 3 && (k == 0 || k == l1 - 1));
        ob[t] = bnd ? v : (ui ? alpha * ui[t] : 0.0) + beta * bmask[e * d.n1e + t] * v;
      }
    }
  }
}

// ---- (lx1, lx2) = (8, 6), 3-D: line-per-thread formulation ---------------------------------------------------
// The generic kernels above spend their time in shared memory (two loads per FMA, one point per thread per stage).
// Here a thread owns a whole grid line: the eight / six values of the line sit in registers, the interpolation and
// derivative matrices come from constant memory (warp-uniform index: broadcast), and shared memory only carries the
// intermediates from one direction to the next -- an order of magnitude fewer shared-memory accesses.  Two elements
// per CTA, 64 threads each: 64 (k, j) lines in r, 48 (k, I) lines in s, 36 (J, I) lines in t.
__constant__ double c_I12[48], c_D12[48];   // [6][8] row-major, N = 7

constexpr int T8_NT = 128, T8_EPC = 2;
constexpr int T8_AB = 8 * 8 * 6, T8_C = 8 * 6 * 6;
constexpr int T8_SMEM = T8_EPC * (2 * T8_AB + 3 * T8_C);   // doubles

__global__ void __launch_bounds__(T8_NT)
opdiv8_kernel(int64_t nel, const double *__restrict__ vel, int64_t fs, const double *__restrict__ rx2, int64_t n2,
              double scale, double *__restrict__ out, const double *__restrict__ dotw,
              double *__restrict__ dot_partial, const int *__restrict__ done) {
  if (done && *done) return;
  __shared__ double sm[T8_SMEM];
  const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int64_t e = (int64_t)blockIdx.x * T8_EPC + g;
  const bool live = e < nel;
  double *A = sm + g * (2 * T8_AB + 3 * T8_C), *B = A + T8_AB, *AA = B + T8_AB, *AD = AA + T8_C, *BA = AD + T8_C;
  double acc[6];
#pragma unroll
  for (int K = 0; K < 6; ++K) acc[K] = 0.0;
  for (int b = 0; b < 3; ++b) {
    if (live) {   // r: line (k, j) = t
      const double2 *up = reinterpret_cast<const double2 *>(vel + (int64_t)b * fs + e * 512 + t * 8);
      double u[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 v = up[q];
        u[2 * q] = v.x;
        u[2 * q + 1] = v.y;
      }
#pragma unroll
      for (int I = 0; I < 6; ++I) {
        double a = 0.0, d = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a = fma(c_I12[I * 8 + i], u[i], a);
          d = fma(c_D12[I * 8 + i], u[i], d);
        }
        A[t * 6 + I] = a;
        B[t * 6 + I] = d;
      }
    }
    __syncthreads();
    if (live && t < 48) {   // s: line (k, I)
      const int k = t / 6, I = t % 6;
      double a[8], d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] = A[(k * 8 + j) * 6 + I];
        d[j] = B[(k * 8 + j) * 6 + I];
      }
#pragma unroll
      for (int J = 0; J < 6; ++J) {
        double aa = 0.0, ad = 0.0, ba = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          aa = fma(c_I12[J * 8 + j], a[j], aa);
          ad = fma(c_D12[J * 8 + j], a[j], ad);
          ba = fma(c_I12[J * 8 + j], d[j], ba);
        }
        AA[(k * 6 + J) * 6 + I] = aa;
        AD[(k * 6 + J) * 6 + I] = ad;
        BA[(k * 6 + J) * 6 + I] = ba;
      }
    }
    __syncthreads();
    if (live && t < 36) {   // t: line (J, I); point (K, J, I) = K * 36 + t
      double aa[8], ad[8], ba[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        aa[k] = AA[k * 36 + t];
        ad[k] = AD[k * 36 + t];
        ba[k] = BA[k * 36 + t];
      }
      const double *r0 = rx2 + (int64_t)(0 * 3 + b) * n2 + e * 216 + t;
      const double *r1 = rx2 + (int64_t)(1 * 3 + b) * n2 + e * 216 + t;
      const double *r2 = rx2 + (int64_t)(2 * 3 + b) * n2 + e * 216 + t;
#pragma unroll
      for (int K = 0; K < 6; ++K) {
        double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          ur = fma(c_I12[K * 8 + k], ba[k], ur);
          us = fma(c_I12[K * 8 + k], ad[k], us);
          ut = fma(c_D12[K * 8 + k], aa[k], ut);
        }
        acc[K] += r0[K * 36] * ur + r1[K * 36] * us + r2[K * 36] * ut;
      }
    }
    __syncthreads();
  }
  double dot = 0.0;
  if (live && t < 36) {
#pragma unroll
    for (int K = 0; K < 6; ++K) {
      const int64_t q = e * 216 + K * 36 + t;
      const double v = scale * acc[K];
      out[q] = v;
      if (dotw) dot += v * dotw[q];
    }
  }
  if (dot_partial) {
    dot = block_reduce_sum<T8_NT>(dot);
    if (threadIdx.x == 0) dot_partial[blockIdx.x] = dot;
  }
}

__global__ void __launch_bounds__(T8_NT)
opgradt8_kernel(int64_t nel, const double *__restrict__ p, const double *__restrict__ rx2, int64_t n2, double *out,
                int64_t fs, int epi, const double *uin, double alpha, double beta, const double *__restrict__ bmask,
                const int *__restrict__ done) {
  if (done && *done) return;
  __shared__ double sm[T8_SMEM];
  const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int64_t e = (int64_t)blockIdx.x * T8_EPC + g;
  const bool live = e < nel;
  double *A = sm + g * (2 * T8_AB + 3 * T8_C), *B = A + T8_AB, *AA = B + T8_AB, *AD = AA + T8_C, *BA = AD + T8_C;
  double pq[6];
  if (live && t < 36) {
#pragma unroll
    for (int K = 0; K < 6; ++K) pq[K] = p[e * 216 + K * 36 + t];
  }
  for (int b = 0; b < 3; ++b) {
    if (live && t < 36) {   // t, transposed: line (J, I)
      const double *r0 = rx2 + (int64_t)(0 * 3 + b) * n2 + e * 216 + t;
      const double *r1 = rx2 + (int64_t)(1 * 3 + b) * n2 + e * 216 + t;
      const double *r2 = rx2 + (int64_t)(2 * 3 + b) * n2 + e * 216 + t;
      double gr[6], gs[6], gt[6];
#pragma unroll
      for (int K = 0; K < 6; ++K) {
        gr[K] = r0[K * 36] * pq[K];
        gs[K] = r1[K * 36] * pq[K];
        gt[K] = r2[K * 36] * pq[K];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double ba = 0.0, ad = 0.0, aa = 0.0;
#pragma unroll
        for (int K = 0; K < 6; ++K) {
          ba = fma(c_I12[K * 8 + k], gr[K], ba);
          ad = fma(c_I12[K * 8 + k], gs[K], ad);
          aa = fma(c_D12[K * 8 + k], gt[K], aa);
        }
        BA[k * 36 + t] = ba;
        AD[k * 36 + t] = ad;
        AA[k * 36 + t] = aa;
      }
    }
    __syncthreads();
    if (live && t < 48) {   // s, transposed: line (k, I)
      const int k = t / 6, I = t % 6;
      double ba[6], ad[6], aa[6];
#pragma unroll
      for (int J = 0; J < 6; ++J) {
        ba[J] = BA[(k * 6 + J) * 6 + I];
        ad[J] = AD[(k * 6 + J) * 6 + I];
        aa[J] = AA[(k * 6 + J) * 6 + I];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double bb = 0.0, a = 0.0;
#pragma unroll
        for (int J = 0; J < 6; ++J) {
          bb = fma(c_I12[J * 8 + j], ba[J], bb);
          a = fma(c_D12[J * 8 + j], ad[J], a);
          a = fma(c_I12[J * 8 + j], aa[J], a);
        }
        B[(k * 8 + j) * 6 + I] = bb;
        A[(k * 8 + j) * 6 + I] = a;
      }
    }
    __syncthreads();
    if (live) {   // r, transposed: line (k, j) = t, then the epilogue on its eight points
      double bb[6], a[6];
#pragma unroll
      for (int I = 0; I < 6; ++I) {
        bb[I] = B[t * 6 + I];
        a[I] = A[t * 6 + I];
      }
      const int j = t & 7, k = t >> 3;
      const bool edge_line = j == 0 || j == 7 || k == 0 || k == 7;
      double *ob = out + (int64_t)b * fs + e * 512 + t * 8;
      const double *ui = uin ? uin + (int64_t)b * fs + e * 512 + t * 8 : nullptr;
      const double *bm = bmask + e * 512 + t * 8;
      double w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double v = 0.0;
#pragma unroll
        for (int I = 0; I < 6; ++I) {
          v = fma(c_D12[I * 8 + i], bb[I], v);
          v = fma(c_I12[I * 8 + i], a[I], v);
        }
        if (epi == 0) {
          v = (ui ? alpha * ui[i] : 0.0) + beta * v;
        } else if (!(edge_line || i == 0 || i == 7)) {
          v = (ui ? alpha * ui[i] : 0.0) + beta * bm[i] * v;
        }
        w[i] = v;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<double2 *>(ob)[q] = make_double2(w[2 * q], w[2 * q + 1]);
    }
    __syncthreads();
  }
}

// map12: element-local field on the velocity mesh -> pressure mesh, times scale[q % n2e] (the Gauss weights)
__global__ void __launch_bounds__(NT_NS)
map12_kernel(NsDims d, const double *__restrict__ in, const double *__restrict__ I12g, const double *__restrict__ w3,
             double *__restrict__ out, int invert) {
  extern __shared__ double sm[];
  const int l1 = d.l1, l2 = d.l2, dim = d.dim, lz1 = dim == 3 ? l1 : 1;
  double *I12 = sm, *u = I12 + 2 * l2 * l1, *A = u + d.n1e, *AA = A + 2 * lz1 * l1 * l2;
  for (int t = threadIdx.x; t < l2 * l1; t += blockDim.x) I12[t] = I12g[t];
  const int64_t e = blockIdx.x;
  for (int t = threadIdx.x; t < d.n1e; t += blockDim.x) u[t] = in[e * d.n1e + t];
  __syncthreads();
  stage<false, false>(A, u, I12, lz1 * l1, l2, l1, 1);
  __syncthreads();
  stage<false, false>(AA, A, I12, lz1, l2, l1, l2);
  __syncthreads();
  for (int q = threadIdx.x; q < d.n2e; q += blockDim.x) {
    double v;
    if (dim == 3) {
      const int K = q / (l2 * l2), ji = q % (l2 * l2);
      v = 0.0;
      for (int k = 0; k < l1; ++k) v = fma(I12[K * l1 + k], AA[k * l2 * l2 + ji], v);
    } else {
      v = AA[q];
    }
    v *= w3[q];
    out[e * d.n2e + q] = invert ? 1.0 / v : v;
  }
}

// ---- pressure CG: device state ------------------------------------------------------------------------
struct PcgP {
  double rtz1, rtz2, r0, rn, mean_z, sums[4];   // sums: 0 (w,p) | 1 sum r z | 2 sum z | 3 sum r   (z = M^-1 r)
  double ntot;                                  // pressure points over all ranks
  int it, done, mean_free, pad;
};

__global__ void reduce_rows_kernel(const double *__restrict__ partial, int rows, int ncol, double *__restrict__ out,
                                   const int *__restrict__ done) {
  if (done && *done) return;
  // one CTA per column, fixed order (thread-strided sums, shuffle tree, warp partials in turn)
  const int c = blockIdx.x;
  double s = 0.0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) s += partial[(size_t)r * ncol + c];
  s = block_reduce_sum<NT_NS>(s);
  if (threadIdx.x == 0) out[c] = s;
}

// r = rhs - mean(rhs) (mean_free) ; x = 0 ; p = 0
__global__ void __launch_bounds__(NT_NS)
pcg_init_kernel(const double *__restrict__ rhs, int64_t n2, const double *__restrict__ sum_rhs, double ntot,
                int mean_free, double *__restrict__ r, double *__restrict__ x, double *__restrict__ p) {
  const double mr = (mean_free && sum_rhs) ? *sum_rhs / ntot : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    r[i] = rhs[i] - mr;
    x[i] = 0.0;
    p[i] = 0.0;
  }
}

__global__ void sum_kernel(const double *__restrict__ a, int64_t n, double *__restrict__ partial) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += a[i];
  s = block_reduce_sum<NT_NS>(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Last stage of either preconditioner: z = minv r (minv given: uzprec's diagonal) or z += zc[element] (the coarse
// correction on top of the element-wise solves; zc may be null), and the partial sums (r, z), sum z, sum r.
__global__ void __launch_bounds__(NT_NS)
precond_sums_kernel(const int *__restrict__ done, const double *__restrict__ r, const double *__restrict__ minv,
                    const double *__restrict__ zc, int n2e, double *__restrict__ z, int64_t n2,
                    double *__restrict__ partial) {
  if (*done) return;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i];
    double zi;
    if (minv) zi = minv[i] * ri;
    else zi = z[i] + (zc ? zc[i / n2e] : 0.0);
    z[i] = zi;
    s1 = fma(ri, zi, s1);
    s2 += zi;
    s3 += ri;
  }
  s1 = block_reduce_sum<NT_NS>(s1);
  s2 = block_reduce_sum<NT_NS>(s2);
  s3 = block_reduce_sum<NT_NS>(s3);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = s1;
    partial[blockIdx.x * 3 + 1] = s2;
    partial[blockIdx.x * 3 + 2] = s3;
  }
}

// after a residual reduction: rtz2 <- rtz1 ; rtz1 <- (r, z - mean z) ; stopping test
__global__ void pcg_after_r_kernel(PcgP *st, double tol, int first) {
  if (st->done) return;
  const double mz = st->mean_free ? st->sums[2] / st->ntot : 0.0;
  const double rtz = st->sums[1] - mz * st->sums[3];
  st->mean_z = mz;
  if (first) {
    st->rtz1 = rtz;
    st->rtz2 = 1.0;
    st->r0 = -1.0;
    st->rn = 0.0;
    st->it = 0;
    if (!(rtz > 0.0)) st->done = 1;
    return;
  }
  st->rtz2 = st->rtz1;
  st->rtz1 = rtz;
  st->rn = sqrt(fabs(rtz));
  if (st->r0 < 0.0) st->r0 = sqrt(fabs(st->rtz2));
  if (st->rn <= tol * st->r0) st->done = 1;
}

// p = (z - mean_z) + beta p
__global__ void __launch_bounds__(NT_NS)
pcg_p_kernel(const PcgP *st, const double *__restrict__ z, double *__restrict__ p, int64_t n2) {
  if (st->done) return;
  const double beta = st->it == 0 ? 0.0 : st->rtz1 / st->rtz2, mz = st->mean_z;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (z[i] - mz) + beta * p[i];
}

// rho = (w, p): not positive -> stop (before the update, like Nek's cggo / uzawa) ; else it += 1
__global__ void pcg_after_w_kernel(PcgP *st) {
  if (st->done) return;
  st->it += 1;
  if (!(st->sums[0] > 0.0)) st->done = 1;
}

// x += alpha p ; r -= alpha w
__global__ void __launch_bounds__(NT_NS)
pcg_xr_kernel(const PcgP *st, const double *__restrict__ p, const double *__restrict__ w, double *__restrict__ x,
              double *__restrict__ r, int64_t n2) {
  if (st->done) return;
  const double alpha = st->rtz1 / st->sums[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    r[i] = fma(-alpha, w[i], r[i]);
  }
}

// ---- two-level preconditioner for E ------------------------------------------------------------------------
// Fine level: element-wise fast diagonalisation -- the local solves of Nek's Schwarz preconditioner [UPSTREAM-RECALL
// fast.f / hsmg.f] without overlap.  Every element is replaced by the box with its mean edge lengths, where
//   E_e = sum_a c_a (E^ in direction a) x (M^ in the others),  E^ = D^ b^-1 D^T,  M^ = I^ b^-1 I^T   (1-D, lx2 x lx2),
//   E^ S = M^ S Lambda, S^T M^ S = 1   =>   E_e^-1 = (S x S x S) diag(1 / sum_a c_a lambda_ia) (S x S x S)^T.
// Coarse level: one constant per element, E_c = R E R^T (sparse, elements sharing a velocity node), Jacobi-CG on the
// device with its scalars in device memory (no host involvement).

// c[e][a] = (2 / L_a) prod_{o != a} (L_o / 2),  L_a = 2 / mean_e(|J grad r_a| / J)
__global__ void __launch_bounds__(NT_NS)
elem_scales_kernel(NsDims d, const double *__restrict__ rst, const double *__restrict__ jac, int64_t npts,
                   double *__restrict__ c) {
  const int64_t e = blockIdx.x;
  double L[3] = {2.0, 2.0, 2.0};
  for (int a = 0; a < d.dim; ++a) {
    double s = 0.0;
    for (int t = threadIdx.x; t < d.n1e; t += blockDim.x) {
      const int64_t p = e * d.n1e + t;
      double g = 0.0;
      for (int b = 0; b < d.dim; ++b) {
        const double v = rst[(int64_t)(a * d.dim + b) * npts + p];
        g = fma(v, v, g);
      }
      s += sqrt(g) / jac[p];
    }
    s = block_reduce_sum<NT_NS>(s);
    __shared__ double sh;
    if (threadIdx.x == 0) sh = s;
    __syncthreads();
    L[a] = 2.0 / (sh / d.n1e);
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int a = 0; a < d.dim; ++a) {
      double v = 2.0 / L[a];
      for (int o = 0; o < d.dim; ++o)
        if (o != a) v *= 0.5 * L[o];
      c[e * 3 + a] = v;
    }
}

// z_e = E_e^-1 r_e ; rc[e] = sum of r_e (restriction to the coarse space)
__global__ void __launch_bounds__(NT_NS)
fdm_kernel(NsDims d, const int *__restrict__ done, const double *__restrict__ r, const double *__restrict__ Sg,
           const double *__restrict__ lamg, const double *__restrict__ c, double *__restrict__ z,
           double *__restrict__ rc) {
  if (*done) return;
  extern __shared__ double sm[];
  const int l2 = d.l2, dim = d.dim, lz2 = dim == 3 ? l2 : 1;
  double *Sm = sm, *lam = Sm + l2 * l2, *b0 = lam + l2, *b1 = b0 + d.n2e;
  for (int t = threadIdx.x; t < l2 * l2; t += blockDim.x) Sm[t] = Sg[t];
  if (threadIdx.x < l2) lam[threadIdx.x] = lamg[threadIdx.x];
  const int64_t e = blockIdx.x;
  double s = 0.0;
  for (int t = threadIdx.x; t < d.n2e; t += blockDim.x) {
    const double v = r[e * d.n2e + t];
    b0[t] = v;
    s += v;
  }
  s = block_reduce_sum<NT_NS>(s);   // contains the barriers that publish Sm, lam, b0
  if (threadIdx.x == 0 && rc) rc[e] = s;
  __syncthreads();
  stage<true, false>(b1, b0, Sm, lz2 * l2, l2, l2, 1);
  __syncthreads();
  stage<true, false>(b0, b1, Sm, lz2, l2, l2, l2);
  __syncthreads();
  double *cur = b0, *oth = b1;
  if (dim == 3) {
    stage<true, false>(b1, b0, Sm, 1, l2, l2, l2 * l2);
    __syncthreads();
    cur = b1;
    oth = b0;
  }
  const double cr = c[e * 3 + 0], cs = c[e * 3 + 1], ct = dim == 3 ? c[e * 3 + 2] : 0.0;
  for (int t = threadIdx.x; t < d.n2e; t += blockDim.x) {
    const int i = t % l2, j = (t / l2) % l2, k = t / (l2 * l2);
    cur[t] /= cr * lam[i] + cs * lam[j] + (dim == 3 ? ct * lam[k] : 0.0);
  }
  __syncthreads();
  if (dim == 3) {
    stage<false, false>(oth, cur, Sm, 1, l2, l2, l2 * l2);
    __syncthreads();
    double *t_ = cur;
    cur = oth;
    oth = t_;
  }
  stage<false, false>(oth, cur, Sm, lz2, l2, l2, l2);
  __syncthreads();
  stage<false, false>(cur, oth, Sm, lz2 * l2, l2, l2, 1);
  __syncthreads();
  for (int t = threadIdx.x; t < d.n2e; t += blockDim.x) z[e * d.n2e + t] = cur[t];
}

// Coarse Jacobi-CG in the mean-free subspace:  P E_c P x = P rc,  P = 1 - 11^T / n  (E_c 1 = 0 on affine elements and
// nearly so on deformed ones: the projection keeps the solve -- and with it the preconditioner -- a fixed linear,
// symmetric operator).  z = P dinv r is never stored: (r, z) = sum r dinv r - mean(dinv r) sum r.
// State: rtz[2] (double-buffered across iterations), rtz0, done.
struct CcState {
  double rtz[2], rtz0, mz[2];
  int done, it;
};

// every CTA sums the [n][3] partials in the same fixed order: identical scalars everywhere, no extra launch
// (loads with .cg: inside the cooperative kernel the partials are rewritten by other CTAs between grid barriers, so
// they must come from L2, never from a stale L1 line or the non-coherent path)
__device__ __forceinline__ void sum_partials3(const double *partial, int n, double (&out)[3]) {
  __shared__ double red[3][NT_NS / 32];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s0 += __ldcg(partial + 3 * i);
    s1 += __ldcg(partial + 3 * i + 1);
    s2 += __ldcg(partial + 3 * i + 2);
  }
  s0 = warp_reduce_sum(s0);
  s1 = warp_reduce_sum(s1);
  s2 = warp_reduce_sum(s2);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
    red[2][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  out[0] = out[1] = out[2] = 0.0;
  for (int i = 0; i < NT_NS / 32; ++i) {
    out[0] += red[0][i];
    out[1] += red[1][i];
    out[2] += red[2][i];
  }
  __syncthreads();
}

__device__ __forceinline__ void store_partials3(double *__restrict__ partial, double a, double b, double c) {
  a = block_reduce_sum<NT_NS>(a);
  b = block_reduce_sum<NT_NS>(b);
  c = block_reduce_sum<NT_NS>(c);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x] = a;
    partial[3 * blockIdx.x + 1] = b;
    partial[3 * blockIdx.x + 2] = c;
  }
}

__global__ void __launch_bounds__(NT_NS)
cc_sum_kernel(const int *__restrict__ odone, const double *__restrict__ rc, int n, double *__restrict__ partial) {
  if (*odone) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  store_partials3(partial, i < n ? rc[i] : 0.0, 0.0, 0.0);
}

// r = rc - mean ; x = 0 ; p = 0 ; partials (r dinv r, dinv r, r)
__global__ void __launch_bounds__(NT_NS)
cc_init_kernel(const int *__restrict__ odone, const double *__restrict__ rc, int n, const double *__restrict__ partial_in,
               const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r, double *__restrict__ p,
               double *__restrict__ partial_out, CcState *st) {
  if (*odone) return;
  double s[3];
  sum_partials3(partial_in, gridDim.x, s);
  const double mean = s[0] / n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ri = 0.0, zi = 0.0;
  if (i < n) {
    ri = rc[i] - mean;
    zi = dinv[i] * ri;
    r[i] = ri;
    x[i] = 0.0;
    p[i] = 0.0;
  }
  store_partials3(partial_out, ri * zi, zi, ri);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->done = 0;
    st->it = 0;
    st->rtz0 = -1.0;
  }
}

// (first the direction of this iteration)  p = (dinv r - mz) + beta p ;  w = E_c p ; partial (p, w).
// The direction update lives here, not at the end of the previous iteration, so that p is complete before any row of
// the product reads it: a separate kernel computes the product.
__global__ void __launch_bounds__(NT_NS)
cc_dir_kernel(const int *__restrict__ odone, CcState *st, int it, const double *__restrict__ r,
              const double *__restrict__ dinv, int n, double *__restrict__ p, const double *__restrict__ partial_rz,
              double tol) {
  if (*odone || st->done) return;
  double s[3];
  sum_partials3(partial_rz, gridDim.x, s);
  const double mz = s[1] / n;
  const double rtzn = s[0] - mz * s[2];
  const double rtz = it == 0 ? 1.0 : st->rtz[it & 1];
  const double rtz0 = it == 0 ? rtzn : st->rtz0;
  const bool stop = !(rtzn > tol * tol * rtz0) || !(rtz > 0.0);
  const double beta = it == 0 ? 0.0 : rtzn / rtz;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !stop) p[i] = fma(beta, p[i], dinv[i] * r[i] - mz);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->rtz[(it + 1) & 1] = rtzn;
    if (it == 0) st->rtz0 = rtzn;
    st->it = it;
    if (stop) st->done = 1;     // read by the NEXT kernels only (this kernel read it at its first instruction)
  }
}

// E_c in ELL storage: entry k of row i at [k * n + i] (neighbouring threads read neighbouring words); rows are padded
// with zero coefficients on their own diagonal index.
__global__ void __launch_bounds__(NT_NS)
cc_spmv_kernel(const int *__restrict__ odone, const CcState *st, int width, const int *__restrict__ col,
               const double *__restrict__ val, const double *__restrict__ p, int n, double *__restrict__ w,
               double *__restrict__ partial_pw) {
  if (*odone || st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double pw = 0.0;
  if (i < n) {
    double sv = 0.0;
    for (int k = 0; k < width; ++k) sv = fma(val[(size_t)k * n + i], p[col[(size_t)k * n + i]], sv);
    w[i] = sv;
    pw = sv * p[i];
  }
  store_partials3(partial_pw, pw, 0.0, 0.0);
}

// alpha = rtz / (p, w) ; x += alpha p ; r -= alpha w ; partials (r dinv r, dinv r, r)
__global__ void __launch_bounds__(NT_NS)
cc_update_kernel(const int *__restrict__ odone, const CcState *st, int it, const double *__restrict__ p,
                 const double *__restrict__ w, const double *__restrict__ dinv, int n, double *__restrict__ x,
                 double *__restrict__ r, const double *__restrict__ partial_pw, double *__restrict__ partial_rz) {
  if (*odone || st->done) return;
  double s[3];
  sum_partials3(partial_pw, gridDim.x, s);
  const double pw = s[0], rtz = st->rtz[(it + 1) & 1];
  if (!(pw > 0.0)) {                           // nothing (left) to solve: the same decision in every CTA
    store_partials3(partial_rz, 0.0, 0.0, 0.0);
    return;
  }
  const double alpha = rtz / pw;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ri = 0.0, zi = 0.0;
  if (i < n) {
    x[i] = fma(alpha, p[i], x[i]);
    ri = fma(-alpha, w[i], r[i]);
    r[i] = ri;
    zi = dinv[i] * ri;
  }
  store_partials3(partial_rz, ri * zi, zi, ri);
}

// The whole coarse solve as ONE cooperative kernel: a thread owns a row, its x / r / p entries stay in registers, p is
// published for the product, and the three reductions of an iteration are grid-wide barriers instead of kernel
// boundaries (the multi-launch version above spends 14 us per iteration on three 4-5 us launches; it remains for
// grids that are not co-resident).  Same arithmetic, same fixed summation order: every CTA forms identical scalars,
// so every thread of the grid takes the same exits.
__global__ void __launch_bounds__(NT_NS)
cc_solve_kernel(const int *__restrict__ odone, CcState *st, int width, const int *__restrict__ col,
                const double *__restrict__ val, const double *__restrict__ dinv, const double *__restrict__ rc, int n,
                double *__restrict__ x, double *p, double *pa, double *pb, double tol, int maxit) {
  cg::grid_group grid = cg::this_grid();
  if (*odone) return;                       // the same word for the whole grid, written by an earlier kernel
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  double s[3];
  store_partials3(pa, live ? rc[i] : 0.0, 0.0, 0.0);
  grid.sync();
  sum_partials3(pa, gridDim.x, s);
  const double di = live ? dinv[i] : 0.0;
  double ri = live ? rc[i] - s[0] / n : 0.0, xi = 0.0, pi = 0.0;
  store_partials3(pb, ri * di * ri, di * ri, ri);
  grid.sync();
  double rtz_old = 1.0, rtz0 = 0.0;
  int it = 0;
  for (; it < maxit; ++it) {
    sum_partials3(pb, gridDim.x, s);
    const double mz = s[1] / n, rtzn = s[0] - mz * s[2];
    if (it == 0) rtz0 = rtzn;
    if (!(rtzn > tol * tol * rtz0) || !(rtz_old > 0.0)) break;
    const double beta = it == 0 ? 0.0 : rtzn / rtz_old;
    pi = fma(beta, pi, di * ri - mz);
    if (live) p[i] = pi;
    grid.sync();                             // p complete
    double wi = 0.0;
    if (live)
      for (int k = 0; k < width; ++k) wi = fma(val[(size_t)k * n + i], __ldcg(p + col[(size_t)k * n + i]), wi);
    store_partials3(pa, live ? wi * pi : 0.0, 0.0, 0.0);
    grid.sync();
    sum_partials3(pa, gridDim.x, s);
    if (!(s[0] > 0.0)) break;
    const double alpha = rtzn / s[0];
    xi = fma(alpha, pi, xi);
    ri = fma(-alpha, wi, ri);
    store_partials3(pb, ri * di * ri, di * ri, ri);
    grid.sync();
    rtz_old = rtzn;
  }
  if (live) x[i] = xi;
  if (i == 0) {
    st->it = it;
    st->done = 1;
  }
}

// out = a + c * b on n entries (pressure update p* + dp, extrapolation 2 p - plag)
__global__ void lin2_kernel(double *__restrict__ out, const double *__restrict__ a, double ca, const double *__restrict__ b,
                            double cb, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ca * a[i] + cb * b[i];
}

// G[c * dim + b] = bm1 * dU_c / dx_b on the GLL points (collocation, Nek's gradm1), once per operator: the
// base-flow-gradient term of the adjoint equations.
__global__ void __launch_bounds__(256)
grad_base_kernel(int dim, int lx, const double *__restrict__ U, int64_t fs, const double *__restrict__ rst,
                 const double *__restrict__ bm1, const double *__restrict__ jac, const double *__restrict__ D,
                 int64_t npts, double *__restrict__ G) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  const int64_t e0 = p / nloc * nloc;
  const int loc = (int)(p - e0);
  const int i = loc % lx, j = (loc / lx) % lx, k = dim == 3 ? loc / (lx * lx) : 0;
  const double w3 = bm1[p] / jac[p];
  for (int c = 0; c < dim; ++c) {
    const double *u = U + (int64_t)c * fs + e0;
    double g[3] = {0.0, 0.0, 0.0};
    for (int l = 0; l < lx; ++l) {   // D[i + lx * l] = dxm1(i, l)
      g[0] = fma(D[i + lx * l], u[l + lx * (j + lx * k)], g[0]);
      g[1] = fma(D[j + lx * l], u[i + lx * (l + lx * k)], g[1]);
      if (dim == 3) g[2] = fma(D[k + lx * l], u[i + lx * (j + lx * l)], g[2]);
    }
    for (int b = 0; b < dim; ++b) {
      double s = 0.0;
      for (int a = 0; a < dim; ++a) s = fma(rst[(int64_t)(a * dim + b) * npts + p], g[a], s);
      G[(int64_t)(c * dim + b) * npts + p] = s * w3;
    }
  }
}

// bf_b -= sum_c G[c][b] v_c   (pointwise)
__global__ void __launch_bounds__(256)
adj_gradterm_kernel(int dim, const double *__restrict__ G, const double *__restrict__ v, int64_t fs, int64_t npts,
                    double *__restrict__ bf) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x) {
    double vc[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < dim; ++c) vc[c] = v[(int64_t)c * fs + p];
    for (int b = 0; b < dim; ++b) {
      double s = 0.0;
      for (int c = 0; c < dim; ++c) s = fma(G[(int64_t)(c * dim + b) * npts + p], vc[c], s);
      bf[(int64_t)b * fs + p] -= s;
    }
  }
}

// norm_grad (core/utils.f90:446-486): sum over velocity components c and directions b of glsc3(du_c/dx_b, bm1s,
// du_c/dx_b) with the collocation derivatives of gradm1 -- the quantity outpost_ks compares with 1.1 to drop
// spurious Ritz vectors (core/eigensolvers.f90:587-594).  Partial sums per CTA, fixed order.
__global__ void __launch_bounds__(NT_NS)
norm_grad_kernel(int dim, int lx, const double *__restrict__ U, int64_t fs, const double *__restrict__ rst,
                 const double *__restrict__ jac, const double *__restrict__ D, const double *__restrict__ w, int64_t npts,
                 double *__restrict__ partial) {
  int nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  double acc = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = p / nloc * nloc;
    const int loc = (int)(p - e0);
    const int i = loc % lx, j = (loc / lx) % lx, k = dim == 3 ? loc / (lx * lx) : 0;
    const double ji = 1.0 / jac[p];
    for (int c = 0; c < dim; ++c) {
      const double *u = U + (int64_t)c * fs + e0;
      double g[3] = {0.0, 0.0, 0.0};
      for (int l = 0; l < lx; ++l) {   // D[i + lx * l] = dxm1(i, l)
        g[0] = fma(D[i + lx * l], u[l + lx * (j + lx * k)], g[0]);
        g[1] = fma(D[j + lx * l], u[i + lx * (l + lx * k)], g[1]);
        if (dim == 3) g[2] = fma(D[k + lx * l], u[i + lx * (j + lx * l)], g[2]);
      }
      const double wp = w[(int64_t)c * fs + p];
      for (int b = 0; b < dim; ++b) {
        double d = 0.0;
        for (int a = 0; a < dim; ++a) d = fma(rst[(int64_t)(a * dim + b) * npts + p], g[a], d);
        d *= ji;
        acc = fma(wp * d, d, acc);
      }
    }
  }
  acc = block_reduce_sum<NT_NS>(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// compute_cfl ([UPSTREAM-RECALL] Nek5000 navier4.f; call sites core/linear_stab.f90:222,231): per GLL point
//     dt ( |u.grad r| / dr_i + |u.grad s| / ds_j + |u.grad t| / dt_k ),   u.grad r = (u rx + v ry + w rz) / jac,
// dr_i = the reference-space spacing getdr builds (one-sided at the ends, centred inside); the maximum per CTA.
// dri[0..lx): 1 / dr_i.
struct CflDri {
  double v[16];
};
__global__ void __launch_bounds__(NT_NS)
cfl_kernel(int dim, int lx, const double *__restrict__ U, int64_t fs, const double *__restrict__ rst,
           const double *__restrict__ jac, CflDri dri, double dt, int64_t npts, double *__restrict__ partial) {
  int nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  double cfl = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x) {
    const int loc = (int)(p % nloc);
    const int ijk[3] = {loc % lx, (loc / lx) % lx, dim == 3 ? loc / (lx * lx) : 0};
    const double ji = 1.0 / jac[p];
    double u[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < dim; ++c) u[c] = U[(int64_t)c * fs + p];
    double m = 0.0;
    for (int a = 0; a < dim; ++a) {
      double ua = 0.0;
      for (int b = 0; b < dim; ++b) ua += u[b] * rst[(int64_t)(a * dim + b) * npts + p];
      m += fabs(dt * (ua * ji) * dri.v[ijk[a]]);
    }
    cfl = fmax(cfl, m);
  }
  __shared__ double red[NT_NS / 32];
  for (int o = 16; o > 0; o >>= 1) cfl = fmax(cfl, __shfl_xor_sync(0xffffffffu, cfl, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cfl;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < NT_NS / 32; ++i) cfl = fmax(cfl, red[i]);
    partial[blockIdx.x] = cfl;
  }
}

__global__ void max_rows_kernel(const double *__restrict__ partial, int rows, double *__restrict__ out) {
  double m = 0.0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) m = fmax(m, partial[r]);
  __shared__ double red[NT_NS / 32];
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < NT_NS / 32; ++i) m = fmax(m, red[i]);
    *out = m;
  }
}

NsDims ns_dims(nsb_sem_t S) {
  NsDims d;
  d.dim = S->dim;
  d.l1 = S->lx;
  d.l2 = S->lx2;
  d.n1e = S->lx * S->lx * (S->dim == 3 ? S->lx : 1);
  d.n2e = S->lx2 * S->lx2 * (S->dim == 3 ? S->lx2 : 1);
  return d;
}

size_t ns_smem(nsb_sem_t S) { return sizeof(double) * ns_smem_doubles(S->dim, S->lx, S->lx2); }

// rows of dot partials an opdiv launch writes
int opdiv_rows(nsb_sem_t S) {
  return (S->lx == 8 && S->dim == 3 && !S->ctx->ns_generic) ? (int)((S->nel + T8_EPC - 1) / T8_EPC) : (int)S->nel;
}

int launch_opdiv(nsb_sem_t S, const double *vel, int64_t fs, double scale, double *out, const double *dotw,
                 double *dot_partial, const int *done) {
  nsb_context_t ctx = S->ctx;
  // algorithmic bytes: dim velocity fields and dim^2 metric arrays read, one pressure array written
  ProfScope ps(ctx, PC_AXHELM, 8.0 * ((double)S->dim * S->npts + ((double)S->dim * S->dim + 1.0) * S->n2));
  if (S->lx == 8 && S->dim == 3 && !ctx->ns_generic)
    opdiv8_kernel<<<(unsigned)((S->nel + T8_EPC - 1) / T8_EPC), T8_NT, 0, ctx->stream>>>(S->nel, vel, fs, S->rx2_d, S->n2,
                                                                                       scale, out, dotw, dot_partial, done);
  else if (S->lx == 8 && S->dim == 3)
    opdiv_kernel<8, 3><<<(unsigned)S->nel, NT_NS, ns_smem(S), ctx->stream>>>(ns_dims(S), vel, fs, S->rx2_d, S->n2, S->i12_d,
                                                                            S->d12_d, scale, out, dotw, dot_partial, done);
  else
    opdiv_kernel<0, 0><<<(unsigned)S->nel, NT_NS, ns_smem(S), ctx->stream>>>(ns_dims(S), vel, fs, S->rx2_d, S->n2, S->i12_d,
                                                                            S->d12_d, scale, out, dotw, dot_partial, done);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

int launch_opgradt(nsb_sem_t S, const double *p, double *out, int64_t fs, int epi, const double *uin, double alpha,
                   double beta, const int *done) {
  nsb_context_t ctx = S->ctx;
  ProfScope ps(ctx, PC_AXHELM, 8.0 * ((double)S->dim * S->npts * (uin ? 2.0 : 1.0) + ((double)S->dim * S->dim + 1.0) * S->n2));
  if (S->lx == 8 && S->dim == 3 && !ctx->ns_generic)
    opgradt8_kernel<<<(unsigned)((S->nel + T8_EPC - 1) / T8_EPC), T8_NT, 0, ctx->stream>>>(
        S->nel, p, S->rx2_d, S->n2, out, fs, epi, uin, alpha, beta, S->bmask_d, done);
  else if (S->lx == 8 && S->dim == 3)
    opgradt_kernel<8, 3><<<(unsigned)S->nel, NT_NS, ns_smem(S), ctx->stream>>>(ns_dims(S), p, S->rx2_d, S->n2, S->i12_d,
                                                                              S->d12_d, out, fs, epi, uin, alpha, beta,
                                                                              S->bmask_d, done);
  else
    opgradt_kernel<0, 0><<<(unsigned)S->nel, NT_NS, ns_smem(S), ctx->stream>>>(ns_dims(S), p, S->rx2_d, S->n2, S->i12_d,
                                                                              S->d12_d, out, fs, epi, uin, alpha, beta,
                                                                              S->bmask_d, done);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// out = alpha uin + beta B^-1 D^T p  on the dim velocity fields at `out` (uin may alias out only when alpha == 0)
int launch_binv_gradt(nsb_sem_t S, const double *p, double *out, int64_t fs, const double *uin, double alpha, double beta,
                      const int *done) {
  NSB_CHECK(launch_opgradt(S, p, out, fs, 1, uin, alpha, beta, done));
  return launch_gs_ext(S, out, S->dim, fs, 1, uin ? uin : out, uin ? alpha : 0.0, beta, S->bmask_d);
}

// velocity-shaped scratch of the mesh: dim fields of npts after the four pressure work vectors (r, p, w, z)
double *ns_vel_scratch(nsb_sem_t S) { return S->ns_work_d + 4 * ((S->n2 + 31) & ~(int64_t)31); }

int pressure_ptr(nsb_sem_t S, nsb_basis_t B, int col, double **out, const char *who) {
  NSB_REQUIRE(S && B, "%s: NULL argument", who);
  NSB_REQUIRE(S->lx2 > 0, "%s: call nsb_sem_pressure_setup first", who);
  NSB_REQUIRE(col >= 0 && col < B->ncols, "%s: column %d out of range", who, col);
  nsb_layout_t L = B->lay;
  NSB_REQUIRE(L->ctx == S->ctx, "%s: basis and mesh live on different contexts", who);
  NSB_REQUIRE(!L->c0_sem, "%s: the C0 storage layout is not supported by the time-stepper pieces", who);
  NSB_REQUIRE(L->nfields > S->dim && L->len[S->dim] == S->n2,
              "%s: field %d of the layout must be the pressure (%lld Gauss points on this mesh)", who, S->dim,
              (long long)S->n2);
  *out = B->col(col) + L->off[S->dim];
  return NSB_OK;
}

int velocity_ptr(nsb_sem_t S, nsb_basis_t B, int col, double **out, int64_t *fs, const char *who) {
  NSB_REQUIRE(S && B, "%s: NULL argument", who);
  NSB_REQUIRE(S->lx2 > 0, "%s: call nsb_sem_pressure_setup first", who);
  NSB_REQUIRE(col >= 0 && col < B->ncols, "%s: column %d out of range", who, col);
  nsb_layout_t L = B->lay;
  NSB_REQUIRE(L->ctx == S->ctx, "%s: basis and mesh live on different contexts", who);
  NSB_REQUIRE(!L->c0_sem, "%s: the C0 storage layout is not supported by the time-stepper pieces", who);
  NSB_REQUIRE(L->nfields >= S->dim, "%s: the layout has %d fields, the velocity needs %d", who, L->nfields, S->dim);
  const int64_t s = S->dim > 1 ? L->off[1] - L->off[0] : 0;
  for (int f = 0; f < S->dim; ++f) {
    NSB_REQUIRE(L->len[f] == S->npts, "%s: field %d has %lld points, mesh has %lld", who, f, (long long)L->len[f],
                (long long)S->npts);
    NSB_REQUIRE(f == 0 || L->off[f] - L->off[f - 1] == s, "%s: velocity fields are not equally spaced", who);
  }
  *out = B->col(col) + L->off[0];
  *fs = s;
  return NSB_OK;
}

size_t fdm_smem(nsb_sem_t S);

// symmetric-definite generalised eigenproblem A s = lambda B s, n <= 16, by Cholesky + cyclic Jacobi:
// S[I * n + m] = component I of eigenvector m, S^T B S = 1.
int sym_gen_eig(int n, const std::vector<double> &A, const std::vector<double> &B, std::vector<double> &S,
                std::vector<double> &lam) {
  std::vector<double> L(n * n, 0.0), C(n * n, 0.0), Y(n * n, 0.0), T(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = B[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0.0)) return NSB_EINVAL;
        L[i * n + i] = std::sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  // T = L^-1 A ; C = T L^-T
  for (int c = 0; c < n; ++c)
    for (int i = 0; i < n; ++i) {
      double s = A[i * n + c];
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * T[k * n + c];
      T[i * n + c] = s / L[i * n + i];
    }
  for (int r = 0; r < n; ++r)
    for (int i = 0; i < n; ++i) {
      double s = T[r * n + i];
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * C[r * n + k];
      C[r * n + i] = s / L[i * n + i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) C[i * n + j] = C[j * n + i] = 0.5 * (C[i * n + j] + C[j * n + i]);
  for (int i = 0; i < n; ++i) Y[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? dia : off) += C[i * n + j] * C[i * n + j];
    if (off <= 1e-30 * dia) break;
    for (int pq = 0; pq < n; ++pq)
      for (int q = pq + 1; q < n; ++q) {
        const int p_ = pq;
        const double apq = C[p_ * n + q];
        if (apq == 0.0) continue;
        const double th = (C[q * n + q] - C[p_ * n + p_]) / (2.0 * apq);
        const double t = (th >= 0.0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < n; ++k) {
          const double ckp = C[k * n + p_], ckq = C[k * n + q];
          C[k * n + p_] = cs * ckp - sn * ckq;
          C[k * n + q] = sn * ckp + cs * ckq;
        }
        for (int k = 0; k < n; ++k) {
          const double cpk = C[p_ * n + k], cqk = C[q * n + k];
          C[p_ * n + k] = cs * cpk - sn * cqk;
          C[q * n + k] = sn * cpk + cs * cqk;
        }
        for (int k = 0; k < n; ++k) {
          const double ykp = Y[k * n + p_], ykq = Y[k * n + q];
          Y[k * n + p_] = cs * ykp - sn * ykq;
          Y[k * n + q] = sn * ykp + cs * ykq;
        }
      }
  }
  lam.resize(n);
  S.assign(n * n, 0.0);
  for (int m = 0; m < n; ++m) lam[m] = C[m * n + m];
  // S = L^-T Y
  for (int m = 0; m < n; ++m)
    for (int i = n - 1; i >= 0; --i) {
      double s = Y[i * n + m];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * S[k * n + m];
      S[i * n + m] = s / L[i * n + i];
    }
  return NSB_OK;
}

// 1-D operators of the element-wise solves on the reference element and their generalised eigen-decomposition.
int fdm_matrices_host(int N, std::vector<double> &Sm, std::vector<double> &lam) {
  const int l1 = N + 1, l2 = N - 1;
  std::vector<double> z2(l2), w2(l2), I12(l2 * l1), D12(l2 * l1), z1(l1), b(l1), D(l1 * l1);
  NSB_CHECK(nsb_pressure_matrices(N, z2.data(), w2.data(), I12.data(), D12.data()));
  NSB_CHECK(nsb_gll(N, z1.data(), b.data(), D.data()));
  b[0] *= 2.0;            // end weights doubled: assembled with an equal neighbour
  b[l1 - 1] *= 2.0;
  std::vector<double> Eh(l2 * l2, 0.0), Mh(l2 * l2, 0.0);
  for (int I = 0; I < l2; ++I)
    for (int J = 0; J < l2; ++J) {
      double se = 0.0, smm = 0.0;
      for (int i = 0; i < l1; ++i) {
        se += w2[I] * D12[I * l1 + i] * w2[J] * D12[J * l1 + i] / b[i];
        smm += w2[I] * I12[I * l1 + i] * w2[J] * I12[J * l1 + i] / b[i];
      }
      Eh[I * l2 + J] = se;
      Mh[I * l2 + J] = smm;
    }
  NSB_REQUIRE(sym_gen_eig(l2, Eh, Mh, Sm, lam) == NSB_OK, "fdm: the 1-D mass operator is not positive definite");
  return NSB_OK;
}

// One-time set-up of the two-level preconditioner (first solve that asks for it).
int fdm_setup(nsb_sem_t S) {
  if (S->fdm_S_d) return NSB_OK;
  nsb_context_t ctx = S->ctx;
  cudaStream_t st = ctx->stream;
  const int l2 = S->lx2, dim = S->dim;
  const NsDims d = ns_dims(S);
  std::vector<double> Sm, lam;
  NSB_CHECK(fdm_matrices_host(S->N, Sm, lam));
  const int64_t nelp = std::max<int64_t>(S->nel, 1);
  NSB_CUDA(cudaMalloc(&S->fdm_S_d, sizeof(double) * l2 * l2));
  NSB_CUDA(cudaMalloc(&S->fdm_lam_d, sizeof(double) * l2));
  NSB_CUDA(cudaMalloc(&S->fdm_c_d, sizeof(double) * 3 * nelp));
  NSB_CUDA(cudaMemcpyAsync(S->fdm_S_d, Sm.data(), sizeof(double) * l2 * l2, cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(S->fdm_lam_d, lam.data(), sizeof(double) * l2, cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaFuncSetAttribute(fdm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fdm_smem(S)));
  if (S->nel > 0) {
    elem_scales_kernel<<<(unsigned)S->nel, NT_NS, 0, st>>>(d, S->rst_d, S->jac_d, S->npts, S->fdm_c_d);
    ctx->launches++;
  }
  NSB_CUDA(cudaGetLastError());
  // coarse operator E_c[f, e] = sum over shared velocity nodes n of bmask_n g_f(n) . g_e(n),  g_e = D^T 1_e
  const int64_t n2p = (S->n2 + 31) & ~(int64_t)31;
  double *ones = S->ns_work_d + 3 * n2p, *wv = ns_vel_scratch(S);   // z doubles as the all-ones pressure
  {
    std::vector<double> h1v((size_t)std::max<int64_t>(S->n2, 1), 1.0);
    NSB_CUDA(cudaMemcpyAsync(ones, h1v.data(), sizeof(double) * S->n2, cudaMemcpyHostToDevice, st));
    NSB_CUDA(cudaStreamSynchronize(st));
  }
  NSB_CHECK(launch_opgradt(S, ones, wv, S->npts, 0, nullptr, 0.0, 1.0, nullptr));
  std::vector<double> g((size_t)dim * S->npts), bm((size_t)S->npts);
  NSB_CUDA(cudaMemcpyAsync(g.data(), wv, sizeof(double) * dim * S->npts, cudaMemcpyDeviceToHost, st));
  NSB_CUDA(cudaMemcpyAsync(bm.data(), S->bmask_d, sizeof(double) * S->npts, cudaMemcpyDeviceToHost, st));
  NSB_CUDA(cudaStreamSynchronize(st));
  const int n = (int)S->nel;
  std::vector<double> diag(n, 0.0);
  for (int64_t x = 0; x < S->npts; ++x) {
    double s = 0.0;
    for (int bb = 0; bb < dim; ++bb) s += g[(size_t)bb * S->npts + x] * g[(size_t)bb * S->npts + x];
    diag[x / d.n1e] += bm[x] * s;
  }
  std::vector<std::vector<std::pair<int, double>>> rows(n);
  auto add = [&](int r_, int c_, double v) {
    for (auto &pr : rows[r_])
      if (pr.first == c_) {
        pr.second += v;
        return;
      }
    rows[r_].push_back({c_, v});
  };
  NSB_REQUIRE((int64_t)S->gs_off_h.size() == S->nshared + 1, "fdm_setup: host copy of the gather-scatter lists missing");
  for (int64_t nd = 0; nd < S->nshared; ++nd)
    for (int64_t qa = S->gs_off_h[nd]; qa < S->gs_off_h[nd + 1]; ++qa)
      for (int64_t qb = qa + 1; qb < S->gs_off_h[nd + 1]; ++qb) {
        const int64_t xa = S->gs_idx_h[qa], xb = S->gs_idx_h[qb];
        double v = 0.0;
        for (int bb = 0; bb < dim; ++bb) v += g[(size_t)bb * S->npts + xa] * g[(size_t)bb * S->npts + xb];
        v *= bm[xa];
        const int ea = (int)(xa / d.n1e), eb = (int)(xb / d.n1e);
        if (ea == eb) diag[ea] += 2.0 * v;
        else {
          add(ea, eb, v);
          add(eb, ea, v);
        }
      }
  int width = 1;
  for (int e = 0; e < n; ++e) width = std::max<int>(width, 1 + (int)rows[e].size());
  std::vector<int> col((size_t)width * std::max(n, 1));
  std::vector<double> val((size_t)width * std::max(n, 1), 0.0), dinv(std::max(n, 1), 0.0);
  for (int e = 0; e < n; ++e) {
    for (int k = 0; k < width; ++k) col[(size_t)k * n + e] = e;
    val[e] = diag[e];
    for (size_t k = 0; k < rows[e].size(); ++k) {
      col[(k + 1) * n + e] = rows[e][k].first;
      val[(k + 1) * n + e] = rows[e][k].second;
    }
    dinv[e] = diag[e] > 0.0 ? 1.0 / diag[e] : 0.0;
  }
  S->cc_nnz = (int64_t)width * n;
  S->cc_width = width;
  const int nb = (n + NT_NS - 1) / NT_NS;
  const int64_t np = ((int64_t)n + 31) & ~(int64_t)31;
  NSB_CUDA(cudaMalloc(&S->cc_col_d, sizeof(int) * std::max<size_t>(col.size(), 1)));
  NSB_CUDA(cudaMalloc(&S->cc_val_d, sizeof(double) * std::max<size_t>(val.size(), 1)));
  NSB_CUDA(cudaMalloc(&S->cc_dinv_d, sizeof(double) * std::max(n, 1)));
  NSB_CUDA(cudaMalloc(&S->cc_vec_d, sizeof(double) * 5 * std::max<int64_t>(np, 32)));
  NSB_CUDA(cudaMalloc(&S->cc_partial_d, sizeof(double) * 6 * std::max(nb, 1)));
  NSB_CUDA(cudaMalloc(&S->cc_state_d, sizeof(CcState)));
  NSB_CUDA(cudaMemcpyAsync(S->cc_col_d, col.data(), sizeof(int) * col.size(), cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(S->cc_val_d, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(S->cc_dinv_d, dinv.data(), sizeof(double) * std::max(n, 1), cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemsetAsync(S->cc_state_d, 0, sizeof(CcState), st));
  NSB_CUDA(cudaStreamSynchronize(st));
  // the one-kernel coarse solve needs the whole grid resident at once
  int coop = 0, per_sm = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cc_solve_kernel, NT_NS, 0);
  S->cc_coop = coop && nb >= 1 && nb <= per_sm * ctx->num_sms;
  return NSB_OK;
}

size_t fdm_smem(nsb_sem_t S) {
  const int l2 = S->lx2, n2e = l2 * l2 * (S->dim == 3 ? l2 : 1);
  return sizeof(double) * (l2 * l2 + l2 + 2 * n2e);
}

// z = M^-1 r and the partial sums (r, z), sum z, sum r in partial[grid][3].
//   precond 0 : 1 / bm2 (Nek's uzprec without the Schwarz part)
//   precond 1 : element-wise fast diagonalisation + coarse correction on the element constants
int precondition(nsb_sem_t S, int precond, const int *done, const double *r, double *z, int grid, double *partial,
                 double tol) {
  nsb_context_t ctx = S->ctx;
  cudaStream_t st = ctx->stream;
  const NsDims d = ns_dims(S);
  if (precond == 0) {
    precond_sums_kernel<<<grid, NT_NS, 0, st>>>(done, r, S->bm2inv_d, nullptr, d.n2e, z, S->n2, partial);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
  const int n = (int)S->nel, nb = (n + NT_NS - 1) / NT_NS;
  const int64_t np = ((int64_t)n + 31) & ~(int64_t)31;
  double *xc = S->cc_vec_d, *rc = xc + np, *pc = rc + np, *wc = pc + np, *rr = wc + np;
  double *pa = S->cc_partial_d, *pb = pa + 3 * nb;
  CcState *cs = reinterpret_cast<CcState *>(S->cc_state_d);
  fdm_kernel<<<(unsigned)S->nel, NT_NS, fdm_smem(S), st>>>(d, done, r, S->fdm_S_d, S->fdm_lam_d, S->fdm_c_d, z, rr);
  ctx->launches++;
  const bool coarse = S->cc_nnz > 0 && !ctx->ns_no_coarse;
  // the coarse solve is part of a preconditioner CG takes for a fixed linear map: two orders below the outer
  // tolerance, at most 1e-6, at least 1e-10
  const double ctol = std::max(1e-10, std::min(1e-6, 1e-2 * tol));
  if (coarse && S->cc_coop && !ctx->ns_generic) {
    int width = S->cc_width, nn = n, mx = S->cc_maxit;
    double tl = ctol;
    const int *colp = S->cc_col_d;
    const double *valp = S->cc_val_d, *dinvp = S->cc_dinv_d, *rrp = rr;
    void *args[] = {(void *)&done, (void *)&cs, (void *)&width, (void *)&colp, (void *)&valp, (void *)&dinvp, (void *)&rrp,
                    (void *)&nn, (void *)&xc, (void *)&pc, (void *)&pa, (void *)&pb, (void *)&tl, (void *)&mx};
    NSB_CUDA(cudaLaunchCooperativeKernel((void *)cc_solve_kernel, dim3(nb), dim3(NT_NS), args, 0, st));
    ctx->launches++;
  } else if (coarse) {
    cc_sum_kernel<<<nb, NT_NS, 0, st>>>(done, rr, n, pa);
    cc_init_kernel<<<nb, NT_NS, 0, st>>>(done, rr, n, pa, S->cc_dinv_d, xc, rc, pc, pb, cs);
    ctx->launches += 2;
    for (int it = 0; it < S->cc_launch; ++it) {
      cc_dir_kernel<<<nb, NT_NS, 0, st>>>(done, cs, it, rc, S->cc_dinv_d, n, pc, pb, ctol);
      cc_spmv_kernel<<<nb, NT_NS, 0, st>>>(done, cs, S->cc_width, S->cc_col_d, S->cc_val_d, pc, n, wc, pa);
      cc_update_kernel<<<nb, NT_NS, 0, st>>>(done, cs, it, pc, wc, S->cc_dinv_d, n, xc, rc, pa, pb);
      ctx->launches += 3;
    }
  }
  precond_sums_kernel<<<grid, NT_NS, 0, st>>>(done, r, nullptr, coarse ? xc : nullptr, d.n2e, z, S->n2, partial);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// E dp = rhs on device arrays; x receives dp.  Host involvement: one poll of the state block every 8 iterations.
int esolve_d(nsb_sem_t S, const double *rhs, double *x, double tol, int maxit, int mean_free, int precond, int *iters,
             double *res) {
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int64_t n2 = S->n2, n2p = (n2 + 31) & ~(int64_t)31;
  double *r = S->ns_work_d, *p = r + n2p, *w = p + n2p, *z = w + n2p, *wv = ns_vel_scratch(S);
  PcgP *state = reinterpret_cast<PcgP *>(S->ns_state_d);
  double *sums = S->ns_state_d + offsetof(PcgP, sums) / sizeof(double);
  const int *done = reinterpret_cast<const int *>(reinterpret_cast<const char *>(state) + offsetof(PcgP, done));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n2 + NT_NS - 1) / NT_NS, (int64_t)ctx->num_sms * 8));
  NSB_CHECK(ensure_partial(ctx, std::max<int64_t>((3 * (int64_t)grid + kMaxK + 7) / (kMaxK + 8) + 1,
                                                  (S->nel + kMaxK + 7) / (kMaxK + 8) + 1)));
  if (precond == 1) NSB_CHECK(fdm_setup(S));
  double *partial = ctx->partial_d;
  double ntot = (double)n2;
  PcgP h0;
  memset(&h0, 0, sizeof(h0));
  h0.mean_free = mean_free;
  if (ctx->nranks > 1) {
    // total number of pressure points: one all-reduce of a device scalar through the state block
    h0.ntot = ntot;
    NSB_CUDA(cudaMemcpyAsync(state, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
    NSB_CHECK(allreduce_sum_d(ctx, S->ns_state_d + offsetof(PcgP, ntot) / sizeof(double), 1));
    NSB_CUDA(cudaMemcpyAsync(&ntot, S->ns_state_d + offsetof(PcgP, ntot) / sizeof(double), sizeof(double),
                             cudaMemcpyDeviceToHost, st));
    NSB_CUDA(cudaStreamSynchronize(st));
  }
  h0.ntot = ntot;
  NSB_CUDA(cudaMemcpyAsync(state, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
  double *sum_rhs = nullptr;
  if (mean_free) {   // mean of the right-hand side over all ranks -> sums[3], read by the init kernel
    sum_kernel<<<grid, NT_NS, 0, st>>>(rhs, n2, partial);
    reduce_rows_kernel<<<1, NT_NS, 0, st>>>(partial, grid, 1, sums + 3, nullptr);
    ctx->launches += 2;
    if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, sums + 3, 1));
    sum_rhs = sums + 3;
  }
  pcg_init_kernel<<<grid, NT_NS, 0, st>>>(rhs, n2, sum_rhs, ntot, mean_free, r, x, p);
  ctx->launches++;
  NSB_CHECK(precondition(S, precond, done, r, z, grid, partial, tol));
  reduce_rows_kernel<<<3, NT_NS, 0, st>>>(partial, grid, 3, sums + 1, nullptr);
  ctx->launches++;
  if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, sums + 1, 3));
  pcg_after_r_kernel<<<1, 1, 0, st>>>(state, tol, 1);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  const int64_t fs = S->npts;
  PcgP *hst = reinterpret_cast<PcgP *>(ctx->hpin + 3 * (kMaxK + 8) + 64);
  for (int it = 1; it <= maxit; ++it) {
    pcg_p_kernel<<<grid, NT_NS, 0, st>>>(state, z, p, n2);
    ctx->launches++;
    NSB_CHECK(launch_binv_gradt(S, p, wv, fs, nullptr, 0.0, 1.0, done));           // wv = B^-1 D^T p
    NSB_CHECK(launch_opdiv(S, wv, fs, 1.0, w, p, partial, done));                 // w = D wv, partial (w, p)
    reduce_rows_kernel<<<1, NT_NS, 0, st>>>(partial, opdiv_rows(S), 1, sums, done);
    ctx->launches++;
    if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, sums, 1));
    pcg_after_w_kernel<<<1, 1, 0, st>>>(state);
    pcg_xr_kernel<<<grid, NT_NS, 0, st>>>(state, p, w, x, r, n2);
    ctx->launches += 2;
    NSB_CHECK(precondition(S, precond, done, r, z, grid, partial, tol));
    reduce_rows_kernel<<<3, NT_NS, 0, st>>>(partial, grid, 3, sums + 1, done);
    ctx->launches++;
    if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, sums + 1, 3));
    pcg_after_r_kernel<<<1, 1, 0, st>>>(state, tol, 0);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    if (it % 8 == 0 || it == maxit) {
      NSB_CUDA(cudaMemcpyAsync(hst, state, sizeof(PcgP), cudaMemcpyDeviceToHost, st));
      CcState *hcs = reinterpret_cast<CcState *>(hst + 1);
      const bool two_level = precond == 1 && S->cc_nnz > 0 && !ctx->ns_no_coarse && !(S->cc_coop && !ctx->ns_generic);
      if (two_level) NSB_CUDA(cudaMemcpyAsync(hcs, S->cc_state_d, sizeof(CcState), cudaMemcpyDeviceToHost, st));
      NSB_CUDA(cudaStreamSynchronize(st));
      if (hst->done) break;
      if (two_level) {
        // the coarse solve stops on the device; enqueue a quarter more iterations than the last one used (all of
        // cc_maxit while it still runs out of them)
        S->cc_launch = hcs->done ? std::min(S->cc_maxit, hcs->it + hcs->it / 4 + 8) : S->cc_maxit;
      }
    }
  }
  NSB_CUDA(cudaMemcpyAsync(hst, state, sizeof(PcgP), cudaMemcpyDeviceToHost, st));
  NSB_CUDA(cudaStreamSynchronize(st));
  NSB_CHECK(check_dev_err(ctx));
  if (iters) *iters = hst->it;
  if (res) *res = hst->r0 > 0.0 ? hst->rn / hst->r0 : 0.0;
  return NSB_OK;
}

const double kBDn[4][4] = {{0, 0, 0, 0}, {1.0, 1.0, 0, 0}, {1.5, 2.0, -0.5, 0}, {11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0}};
const double kABn[4][3] = {{0, 0, 0}, {1.0, 0, 0}, {2.0, -1.0, 0}, {3.0, -3.0, 1.0}};

}  // namespace

void nsb::ns_free(nsb_sem_t S) {
  for (double *q : {S->i12_d, S->d12_d, S->rx2_d, S->bm2inv_d, S->ns_work_d, S->ns_state_d, S->fdm_S_d, S->fdm_lam_d,
                    S->fdm_c_d, S->cc_val_d, S->cc_dinv_d, S->cc_vec_d, S->cc_partial_d, S->cc_state_d})
    if (q) cudaFree(q);
  if (S->cc_rowptr_d) cudaFree(S->cc_rowptr_d);
  if (S->cc_col_d) cudaFree(S->cc_col_d);
  S->i12_d = S->d12_d = S->rx2_d = S->bm2inv_d = S->ns_work_d = S->ns_state_d = nullptr;
  S->fdm_S_d = S->fdm_lam_d = S->fdm_c_d = S->cc_val_d = S->cc_dinv_d = S->cc_vec_d = S->cc_partial_d = S->cc_state_d = nullptr;
  S->cc_rowptr_d = S->cc_col_d = nullptr;
  S->cc_nnz = 0;
  S->lx2 = 0;
}

// Host-only: the matrices of the pressure mesh (compared with an independent construction by the CPU tests).
extern "C" int nsb_pressure_matrices(int N, double *z2, double *w2, double *I12, double *D12) {
  NSB_REQUIRE(N >= 3 && N <= 15, "nsb_pressure_matrices: N=%d (3..15)", N);
  const int l1 = N + 1, l2 = N - 1;
  std::vector<double> z1(l1), w1(l1), D(l1 * l1), zz, ww;
  NSB_CHECK(nsb_gll(N, z1.data(), w1.data(), D.data()));   // D[i + l1 * j] = dxm1(i, j)
  gauss_legendre(l2, zz, ww);
  const std::vector<double> J = interp_matrix(z1, zz);      // [l2][l1]
  for (int I = 0; I < l2; ++I) {
    if (z2) z2[I] = zz[I];
    if (w2) w2[I] = ww[I];
    for (int i = 0; i < l1; ++i) {
      if (I12) I12[I * l1 + i] = J[(size_t)I * l1 + i];
      if (D12) {
        double s = 0.0;
        for (int m = 0; m < l1; ++m) s += J[(size_t)I * l1 + m] * D[m + l1 * i];
        D12[I * l1 + i] = s;
      }
    }
  }
  return NSB_OK;
}

// Host-only: the 1-D generalised eigenpairs of the element-wise pressure solves, E^ S = M^ S Lambda with S^T M^ S = 1
// (S[I * lx2 + m] = component I of eigenvector m).
extern "C" int nsb_fdm_matrices(int N, double *S, double *lam) {
  NSB_REQUIRE(N >= 3 && N <= 15 && S && lam, "nsb_fdm_matrices: bad argument");
  std::vector<double> Sm, lm;
  NSB_CHECK(fdm_matrices_host(N, Sm, lm));
  memcpy(S, Sm.data(), sizeof(double) * Sm.size());
  memcpy(lam, lm.data(), sizeof(double) * lm.size());
  return NSB_OK;
}

extern "C" int nsb_sem_pressure_setup(nsb_sem_t S) {
  NSB_REQUIRE(S, "nsb_sem_pressure_setup: NULL argument");
  NSB_REQUIRE(S->N >= 3, "nsb_sem_pressure_setup: P_N - P_N-2 needs N >= 3 (N = %d)", S->N);
  if (S->lx2 > 0) return NSB_OK;
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  const int l1 = S->lx, l2 = l1 - 2, dim = S->dim;
  const int n2e = l2 * l2 * (dim == 3 ? l2 : 1);
  NSB_REQUIRE(n2e <= MAXQ * NT_NS, "nsb_sem_pressure_setup: N=%d not supported in %d-D", S->N, dim);
  std::vector<double> z2(l2), w2(l2), I12(l2 * l1), D12(l2 * l1), w3(n2e);
  NSB_CHECK(nsb_pressure_matrices(S->N, z2.data(), w2.data(), I12.data(), D12.data()));
  for (int q = 0; q < n2e; ++q) {
    const int i = q % l2, j = (q / l2) % l2, k = q / (l2 * l2);
    w3[q] = w2[i] * w2[j] * (dim == 3 ? w2[k] : 1.0);
  }
  S->lx2 = l2;
  S->n2 = S->nel * n2e;
  const int64_t n2p = (S->n2 + 31) & ~(int64_t)31;
  double *w3_d = nullptr;
  auto fail = [&](const char *what) {
    if (w3_d) cudaFree(w3_d);
    ns_free(S);
    set_error("nsb_sem_pressure_setup: %s", what);
    return NSB_ECUDA;
  };
  if (cudaMalloc(&S->i12_d, sizeof(double) * l2 * l1) != cudaSuccess ||
      cudaMalloc(&S->d12_d, sizeof(double) * l2 * l1) != cudaSuccess ||
      cudaMalloc(&S->rx2_d, sizeof(double) * dim * dim * std::max<int64_t>(S->n2, 1)) != cudaSuccess ||
      cudaMalloc(&S->bm2inv_d, sizeof(double) * std::max<int64_t>(S->n2, 1)) != cudaSuccess ||
      cudaMalloc(&S->ns_work_d, sizeof(double) * (4 * n2p + dim * std::max<int64_t>(S->npts, 1))) != cudaSuccess ||
      cudaMalloc(&S->ns_state_d, sizeof(PcgP)) != cudaSuccess || cudaMalloc(&w3_d, sizeof(double) * n2e) != cudaSuccess)
    return fail("out of device memory");
  S->lx2 = l2;
  cudaStream_t st = ctx->stream;
  if (l1 == 8) {   // constant-memory copies for the line-per-thread kernels (the values depend on N only)
    NSB_CUDA(cudaMemcpyToSymbol(c_I12, I12.data(), sizeof(double) * 48));
    NSB_CUDA(cudaMemcpyToSymbol(c_D12, D12.data(), sizeof(double) * 48));
  }
  NSB_CUDA(cudaMemcpyAsync(S->i12_d, I12.data(), sizeof(double) * l2 * l1, cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(S->d12_d, D12.data(), sizeof(double) * l2 * l1, cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(w3_d, w3.data(), sizeof(double) * n2e, cudaMemcpyHostToDevice, st));
  NSB_CUDA(cudaMemsetAsync(S->ns_state_d, 0, sizeof(PcgP), st));
  const size_t smem = ns_smem(S);
  NSB_CUDA(cudaFuncSetAttribute(opdiv_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NSB_CUDA(cudaFuncSetAttribute(opgradt_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NSB_CUDA(cudaFuncSetAttribute(opdiv_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NSB_CUDA(cudaFuncSetAttribute(opgradt_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NSB_CUDA(cudaFuncSetAttribute(map12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (S->nel > 0) {
    const NsDims d = ns_dims(S);
    for (int m = 0; m < dim * dim; ++m)   // rxm2 = w3m2 * map12(rxm1)
      map12_kernel<<<(unsigned)S->nel, NT_NS, smem, st>>>(d, S->rst_d + (int64_t)m * S->npts, S->i12_d, w3_d,
                                                         S->rx2_d + (int64_t)m * S->n2, 0);
    map12_kernel<<<(unsigned)S->nel, NT_NS, smem, st>>>(d, S->jac_d, S->i12_d, w3_d, S->bm2inv_d, 1);
    ctx->launches += dim * dim + 1;
  }
  NSB_CUDA(cudaGetLastError());
  NSB_CUDA(cudaStreamSynchronize(st));
  cudaFree(w3_d);
  return NSB_OK;
}

extern "C" int64_t nsb_sem_npres(nsb_sem_t S) { return (S && S->lx2 > 0) ? S->n2 : -1; }

// which: 0 rx2 [dim*dim][n2], 1 bm2inv [n2]
extern "C" int nsb_sem_pressure_get(nsb_sem_t S, int which, double *out) {
  NSB_REQUIRE(S && out && S->lx2 > 0, "nsb_sem_pressure_get: bad argument (nsb_sem_pressure_setup first)");
  NSB_REQUIRE(which == 0 || which == 1, "nsb_sem_pressure_get: unknown selector %d", which);
  cudaSetDevice(S->ctx->device);
  NSB_CUDA(cudaStreamSynchronize(S->ctx->stream));
  const size_t n = which == 0 ? (size_t)S->dim * S->dim * S->n2 : (size_t)S->n2;
  NSB_CUDA(cudaMemcpy(out, which == 0 ? S->rx2_d : S->bm2inv_d, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return NSB_OK;
}

extern "C" int nsb_sem_opdiv(nsb_sem_t S, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  double *v, *p;
  int64_t fs;
  NSB_CHECK(velocity_ptr(S, bin, cin, &v, &fs, "nsb_sem_opdiv"));
  NSB_CHECK(pressure_ptr(S, bout, cout, &p, "nsb_sem_opdiv"));
  cudaSetDevice(S->ctx->device);
  return launch_opdiv(S, v, fs, 1.0, p, nullptr, nullptr, nullptr);
}

extern "C" int nsb_sem_opgradt(nsb_sem_t S, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  double *v, *p;
  int64_t fs;
  NSB_CHECK(pressure_ptr(S, bin, cin, &p, "nsb_sem_opgradt"));
  NSB_CHECK(velocity_ptr(S, bout, cout, &v, &fs, "nsb_sem_opgradt"));
  cudaSetDevice(S->ctx->device);
  return launch_opgradt(S, p, v, fs, 0, nullptr, 0.0, 1.0, nullptr);
}

extern "C" int nsb_sem_cdabdtp(nsb_sem_t S, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  double *pi, *po;
  NSB_CHECK(pressure_ptr(S, bin, cin, &pi, "nsb_sem_cdabdtp"));
  NSB_CHECK(pressure_ptr(S, bout, cout, &po, "nsb_sem_cdabdtp"));
  NSB_REQUIRE(S->exchange_ready || S->ctx->nranks == 1, "nsb_sem_cdabdtp: call nsb_sem_setup_exchange first");
  cudaSetDevice(S->ctx->device);
  double *wv = ns_vel_scratch(S);
  NSB_CHECK(launch_binv_gradt(S, pi, wv, S->npts, nullptr, 0.0, 1.0, nullptr));
  return launch_opdiv(S, wv, S->npts, 1.0, po, nullptr, nullptr, nullptr);
}

extern "C" int nsb_sem_esolve(nsb_sem_t S, nsb_basis_t brhs, int crhs, nsb_basis_t bx, int cx, double tol, int maxit,
                              int mean_free, int precond, int *iters, double *res) {
  double *rhs, *x;
  NSB_CHECK(pressure_ptr(S, brhs, crhs, &rhs, "nsb_sem_esolve"));
  NSB_CHECK(pressure_ptr(S, bx, cx, &x, "nsb_sem_esolve"));
  NSB_REQUIRE(rhs != x, "nsb_sem_esolve: right-hand side and solution are the same vector");
  NSB_REQUIRE(tol > 0.0 && maxit >= 1, "nsb_sem_esolve: bad tolerance / iteration limit");
  NSB_REQUIRE(precond == 0 || precond == 1, "nsb_sem_esolve: unknown preconditioner %d", precond);
  NSB_REQUIRE(S->exchange_ready || S->ctx->nranks == 1, "nsb_sem_esolve: call nsb_sem_setup_exchange first");
  return esolve_d(S, rhs, x, tol, maxit, mean_free, precond, iters, res);
}

// norm_grad of the velocity fields of (b, col) (core/utils.f90:446-486; the spurious-mode filter of outpost_ks,
// core/eigensolvers.f90:587-594): sum_c sum_b (du_c/dx_b, du_c/dx_b)_w with the layout's weight (bm1s) -- not the
// square root, exactly what the reference compares with 1.1.  Summed over all ranks.
extern "C" int nsb_sem_norm_grad(nsb_sem_t S, nsb_basis_t B, int col, double *norma) {
  NSB_REQUIRE(S && B && norma, "nsb_sem_norm_grad: NULL argument");
  NSB_REQUIRE(col >= 0 && col < B->ncols, "nsb_sem_norm_grad: column %d out of range", col);
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = S->ctx;
  NSB_REQUIRE(L->ctx == ctx, "nsb_sem_norm_grad: basis and mesh live on different contexts");
  NSB_REQUIRE(!L->c0_sem, "nsb_sem_norm_grad: the C0 storage layout is not supported");
  NSB_REQUIRE(L->nfields >= S->dim, "nsb_sem_norm_grad: the layout has %d fields, the velocity needs %d", L->nfields, S->dim);
  const int64_t fs = S->dim > 1 ? L->off[1] - L->off[0] : 0;
  for (int f = 0; f < S->dim; ++f) {
    NSB_REQUIRE(L->len[f] == S->npts && L->in_dot[f], "nsb_sem_norm_grad: field %d is not a weighted velocity field of the mesh", f);
    NSB_REQUIRE(f == 0 || L->off[f] - L->off[f - 1] == fs, "nsb_sem_norm_grad: velocity fields are not equally spaced");
  }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((S->npts + NT_NS - 1) / NT_NS, (int64_t)ctx->num_sms * 8));
  NSB_CHECK(ensure_partial(ctx, (grid + kMaxK + 7) / (kMaxK + 8) + 1));
  norm_grad_kernel<<<grid, NT_NS, 0, st>>>(S->dim, S->lx, B->col(col) + L->off[0], fs, S->rst_d, S->jac_d, S->D_d,
                                           L->w_d + L->off[0], S->npts, ctx->partial_d);
  reduce_rows_kernel<<<1, NT_NS, 0, st>>>(ctx->partial_d, grid, 1, ctx->hvec_d, nullptr);
  ctx->launches += 2;
  NSB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, ctx->hvec_d, 1));
  NSB_CUDA(cudaMemcpyAsync(norma, ctx->hvec_d, sizeof(double), cudaMemcpyDeviceToHost, st));
  NSB_CUDA(cudaStreamSynchronize(st));
  return check_dev_err(ctx);
}

// compute_cfl(cfl, vx, vy, vz, dt) for the velocity fields of (b, col): what set_linear_solver derives the time step and
// the number of steps of the linearised solver from (core/linear_stab.f90:220-236: dt = ctarg / compute_cfl(.., 1),
// nsteps = ceiling(T / dt), dt = T / nsteps).  Maximum over all ranks (each rank's value in its own slot of a summed
// vector: the library's only collective is a sum).
extern "C" int nsb_sem_cfl(nsb_sem_t S, nsb_basis_t B, int col, double dt, double *cfl) {
  NSB_REQUIRE(S && B && cfl, "nsb_sem_cfl: NULL argument");
  NSB_REQUIRE(col >= 0 && col < B->ncols, "nsb_sem_cfl: column %d out of range", col);
  nsb_layout_t L = B->lay;
  nsb_context_t ctx = S->ctx;
  NSB_REQUIRE(L->ctx == ctx, "nsb_sem_cfl: basis and mesh live on different contexts");
  NSB_REQUIRE(!L->c0_sem, "nsb_sem_cfl: the C0 storage layout is not supported");
  NSB_REQUIRE(L->nfields >= S->dim, "nsb_sem_cfl: the layout has %d fields, the velocity needs %d", L->nfields, S->dim);
  NSB_REQUIRE(S->lx <= 16, "nsb_sem_cfl: lx1 = %d above 16", S->lx);
  const int64_t fs = S->dim > 1 ? L->off[1] - L->off[0] : 0;
  for (int f = 0; f < S->dim; ++f) {
    NSB_REQUIRE(L->len[f] == S->npts, "nsb_sem_cfl: field %d is not a velocity field of the mesh", f);
    NSB_REQUIRE(f == 0 || L->off[f] - L->off[f - 1] == fs, "nsb_sem_cfl: velocity fields are not equally spaced");
  }
  std::vector<double> z(S->lx), w(S->lx), D((size_t)S->lx * S->lx);
  NSB_CHECK(nsb_gll(S->N, z.data(), w.data(), D.data()));
  CflDri dri;
  memset(&dri, 0, sizeof(dri));
  const int lx = S->lx;
  dri.v[0] = 1.0 / (z[1] - z[0]);                                   // getdr: one-sided at the ends,
  for (int i = 1; i < lx - 1; ++i) dri.v[i] = 1.0 / (0.5 * (z[i + 1] - z[i - 1]));   // centred inside
  dri.v[lx - 1] = 1.0 / (z[lx - 1] - z[lx - 2]);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((S->npts + NT_NS - 1) / NT_NS, (int64_t)ctx->num_sms * 8));
  NSB_CHECK(ensure_partial(ctx, (grid + kMaxK + 7) / (kMaxK + 8) + 1));
  NSB_REQUIRE(ctx->nranks <= kMaxK, "nsb_sem_cfl: %d ranks", ctx->nranks);
  NSB_CUDA(cudaMemsetAsync(ctx->hvec_d, 0, sizeof(double) * ctx->nranks, st));
  cfl_kernel<<<grid, NT_NS, 0, st>>>(S->dim, lx, B->col(col) + L->off[0], fs, S->rst_d, S->jac_d, dri, dt, S->npts,
                                     ctx->partial_d);
  max_rows_kernel<<<1, NT_NS, 0, st>>>(ctx->partial_d, grid, ctx->hvec_d + ctx->rank);
  ctx->launches += 2;
  NSB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, ctx->hvec_d, ctx->nranks));
  std::vector<double> all(ctx->nranks);
  NSB_CUDA(cudaMemcpyAsync(all.data(), ctx->hvec_d, sizeof(double) * ctx->nranks, cudaMemcpyDeviceToHost, st));
  NSB_CUDA(cudaStreamSynchronize(st));
  NSB_CHECK(check_dev_err(ctx));
  double m = 0.0;
  for (double a : all) m = a > m ? a : m;
  *cfl = m;
  return NSB_OK;
}

// Operator handle with the structure of exponential_prop%matvec for the linearised Navier-Stokes equations:
// fields 0..dim-1 of the layout are the velocity, field dim the pressure (lx2 mesh); the input vector's velocity AND
// pressure start nsteps BDF/EXT steps (order ramp 1, 2, 3: the reference restarts the time-stepper for every
// matvec), the output vector receives the final velocity and pressure.  base/col_base: base flow U (velocity fields
// of that column), NULL for the Stokes operator.  Uses both convection slots of the mesh (0: U, 1: v).
extern "C" int nsb_op_create_ns_stepper(nsb_sem_t S, nsb_layout_t layout, nsb_basis_t base, int col_base, double nu,
                                        double dt, int nsteps, double tol_v, double tol_p, int maxit, int mean_free,
                                        int precond, nsb_op_t *out) {
  NSB_REQUIRE(S && layout && out, "nsb_op_create_ns_stepper: NULL argument");
  NSB_REQUIRE(nu > 0.0 && dt > 0.0 && nsteps >= 1 && maxit >= 1 && tol_v > 0.0 && tol_p > 0.0 &&
                  (precond == 0 || precond == 1),
              "nsb_op_create_ns_stepper: bad parameter");
  NSB_REQUIRE(layout->ctx == S->ctx, "nsb_op_create_ns_stepper: layout and mesh live on different contexts");
  NSB_REQUIRE(S->exchange_ready, "nsb_op_create_ns_stepper: call nsb_sem_setup_exchange first");
  NSB_CHECK(nsb_sem_pressure_setup(S));
  NSB_REQUIRE(!layout->c0_sem, "nsb_op_create_ns_stepper: the C0 storage layout is not supported");
  NSB_REQUIRE(layout->nfields > S->dim && layout->len[S->dim] == S->n2,
              "nsb_op_create_ns_stepper: field %d of the layout must be the pressure (%lld Gauss points)", S->dim,
              (long long)S->n2);
  for (int f = 0; f < S->dim; ++f)
    NSB_REQUIRE(layout->len[f] == S->npts, "nsb_op_create_ns_stepper: field %d has %lld points, mesh has %lld", f,
                (long long)layout->len[f], (long long)S->npts);
  if (base) {
    NSB_REQUIRE(base->lay == layout && col_base >= 0 && col_base < base->ncols,
                "nsb_op_create_ns_stepper: the base flow must be a column of a basis with the same layout");
    NSB_REQUIRE(S->lxd > 0, "nsb_op_create_ns_stepper: call nsb_sem_dealias_setup first");
  }
  nsb_op_t op = new nsb_op_s();
  op->kind = 4;
  op->sem = S;
  op->lay = layout;
  op->nfields_apply = S->dim + 1;
  op->nu = nu;
  op->dt = dt;
  op->nsteps = nsteps;
  op->tol = tol_v;
  op->tol_p = tol_p;
  op->maxit = maxit;
  op->mean_free = mean_free;
  op->precond = precond;
  op->has_base = base != nullptr;
  int r = nsb_basis_create(layout, 10, &op->tmp);
  if (r == NSB_OK && base) {
    r = nsb_vec_copy(op->tmp, 9, base, col_base);
    if (r == NSB_OK) r = nsb_sem_set_convect(S, 0, op->tmp, 9, 0);
  }
  if (r != NSB_OK) {
    if (op->tmp) nsb_basis_destroy(op->tmp);
    delete op;
    return r;
  }
  *out = op;
  return NSB_OK;
}

// exponential_prop%rmatvec (core/linear_operators.f90:84-103) for the same equations: Nek's stepper in adjoint mode,
// i.e. the same BDF/EXT splitting on the continuous adjoint equations
//     -dw/dt ... :  dw/dt - (U.grad) w + sum_c w_c grad U_c = -grad q + nu lap w,  div w = 0
// -- the explicit term becomes +(U.grad) w (dealiased, like the forward one) - sum_c w_c grad U_c (pointwise on the
// GLL mesh with gradm1 of the base flow, computed once here).  <A v, w>_B = <v, A+ w>_B then holds up to the
// discretisation error (first order in dt), not to rounding (tests/test_gpu_ns.py).
extern "C" int nsb_op_create_ns_stepper_adjoint(nsb_sem_t S, nsb_layout_t layout, nsb_basis_t base, int col_base, double nu,
                                                double dt, int nsteps, double tol_v, double tol_p, int maxit,
                                                int mean_free, int precond, nsb_op_t *out) {
  NSB_CHECK(nsb_op_create_ns_stepper(S, layout, base, col_base, nu, dt, nsteps, tol_v, tol_p, maxit, mean_free, precond, out));
  nsb_op_t op = *out;
  op->adjoint = true;
  if (!op->has_base) return NSB_OK;
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  const int dim = S->dim;
  if (cudaMalloc(&op->c_d, sizeof(double) * dim * dim * S->npts) != cudaSuccess) {
    nsb_op_destroy(op);
    *out = nullptr;
    set_error("nsb_op_create_ns_stepper_adjoint: out of device memory");
    return NSB_ECUDA;
  }
  const int64_t fs = dim > 1 ? layout->off[1] - layout->off[0] : 0;
  grad_base_kernel<<<(unsigned)((S->npts + 255) / 256), 256, 0, ctx->stream>>>(
      dim, S->lx, op->tmp->col(9) + layout->off[0], fs, S->rst_d, S->bm1_d, S->jac_d, S->D_d, S->npts, op->c_d);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// Time-periodic base flows (Floquet analysis, Newton for periodic orbits): the reference stores the orbit of the
// nonlinear solution, uor / vor / wor(:, istep), and copies it into vx after every step of the linearised solver
// (core/linear_operators.f90:254-275, core/matvec.f90:347-362), so step istep linearises about uor(:, istep-1) with
// uor(:, 0) = ubase.  Here the orbit is a run of columns of a device basis: step n uses column col0 + (n-1) stride
// (stride -1 walks it backwards for the adjoint).  orbit = NULL: back to the steady base flow of the constructor.
extern "C" int nsb_op_ns_set_orbit(nsb_op_t op, nsb_basis_t orbit, int col0, int stride) {
  NSB_REQUIRE(op && op->kind == 4, "nsb_op_ns_set_orbit: not a Navier-Stokes stepper operator");
  if (!orbit) {
    op->orbit_b = nullptr;
    return NSB_OK;
  }
  NSB_REQUIRE(op->has_base, "nsb_op_ns_set_orbit: the operator was created without a base flow (Stokes)");
  NSB_REQUIRE(orbit->lay == op->lay, "nsb_op_ns_set_orbit: the orbit must be stored in a basis with the operator's layout");
  const int last = col0 + (op->nsteps - 1) * stride;
  NSB_REQUIRE(col0 >= 0 && col0 < orbit->ncols && last >= 0 && last < orbit->ncols,
              "nsb_op_ns_set_orbit: %d steps from column %d with stride %d leave the basis (%d columns)", op->nsteps, col0,
              stride, orbit->ncols);
  op->orbit_b = orbit;
  op->orbit_c0 = col0;
  op->orbit_stride = stride;
  return NSB_OK;
}

extern "C" int nsb_op_ns_iterations(nsb_op_t op, int64_t *helmholtz, int64_t *pressure) {
  NSB_REQUIRE(op && op->kind == 4, "nsb_op_ns_iterations: not a Navier-Stokes stepper operator");
  if (helmholtz) *helmholtz = op->helm_iters;
  if (pressure) *pressure = op->pres_iters;
  return NSB_OK;
}

int nsb::ns_stepper_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  NSB_REQUIRE(bin->lay == op->lay, "nsb_op_apply: time-stepper operator built for another layout");
  nsb_sem_t S = op->sem;
  nsb_basis_t W = op->tmp;
  nsb_layout_t L = op->lay;
  nsb_context_t ctx = L->ctx;
  const int dim = S->dim;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  // work columns of W.  velocity fields: lag[3] rolling (0, 1, 2 + the free one), 4 bf, 5 e1, 6 e2, 7 v*, 9 U
  //                     pressure field:  col 0 p, 1 plag, 2 p*, 3 rhs of E, 4 dp
  for (int c = 0; c < 9; ++c) NSB_CHECK(nsb_vec_zero(W, c));
  int lag[3] = {0, 1, 2}, nw = 3;
  const int bf = 4, e1 = 5, e2 = 6, vs = 7, cU = 9;
  const int64_t fs = dim > 1 ? L->off[1] - L->off[0] : 0;
  const int64_t n2 = S->n2;
  const int g2 = (int)std::min<int64_t>((n2 + 255) / 256 + 1, (int64_t)ctx->num_sms * 8);
  auto vel = [&](int c) { return W->col(c) + L->off[0]; };
  auto prs = [&](int c) { return W->col(c) + L->off[dim]; };
  // cold start: velocity and pressure of the input vector
  for (int f = 0; f < dim; ++f)
    NSB_CUDA(cudaMemcpyAsync(vel(lag[0]) + f * fs, bin->col(cin) + L->off[f], sizeof(double) * S->npts,
                             cudaMemcpyDeviceToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(prs(0), bin->col(cin) + L->off[dim], sizeof(double) * n2, cudaMemcpyDeviceToDevice, st));
  // the base flow this step linearises about: the steady one (column cU of W) or a column of the stored orbit.  Slot 0
  // of the mesh is shared by every operator built on it, so it is refreshed at every application.
  nsb_basis_t Ub = W;
  int Uc = cU;
  auto set_base = [&](bool grad) -> int {
    NSB_CHECK(nsb_sem_set_convect(S, 0, Ub, Uc, 0));
    if (grad && op->adjoint) {
      grad_base_kernel<<<(unsigned)((S->npts + 255) / 256), 256, 0, st>>>(dim, S->lx, Ub->col(Uc) + L->off[0], fs, S->rst_d,
                                                                        S->bm1_d, S->jac_d, S->D_d, S->npts, op->c_d);
      ctx->launches++;
      NSB_CUDA(cudaGetLastError());
    }
    return NSB_OK;
  };
  if (op->has_base && !op->orbit_b) {
    NSB_CHECK(set_base(op->c_dirty));
    op->c_dirty = false;
  }
  for (int n = 1; n <= op->nsteps; ++n) {
    const int o = n < 3 ? n : 3;
    const double bd0 = kBDn[o][0];
    if (op->has_base && op->orbit_b) {
      Ub = op->orbit_b;
      Uc = op->orbit_c0 + (n - 1) * op->orbit_stride;
      NSB_CHECK(set_base(true));
      op->c_dirty = true;          // the gradient of the steady base flow has to be rebuilt once the orbit is dropped
    }
    // advabp: bf = -[(U.grad) v + (v.grad) U], mass matrix inside the dealiased quadrature
    if (op->has_base && op->adjoint) {
      // adjoint advabp: +(U.grad) w - sum_c w_c grad U_c
      NSB_CHECK(nsb_sem_convect(S, 0, W, lag[0], W, bf, 0, dim, 1.0, 0));
      adj_gradterm_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(dim, op->c_d, vel(lag[0]), fs, S->npts, vel(bf));
      ctx->launches++;
    } else if (op->has_base) {
      NSB_CHECK(nsb_sem_convect(S, 0, W, lag[0], W, bf, 0, dim, -1.0, 0));
      NSB_CHECK(nsb_sem_set_convect(S, 1, W, lag[0], 0));
      NSB_CHECK(nsb_sem_convect(S, 1, Ub, Uc, W, bf, 0, dim, -1.0, 1));
    } else {
      for (int f = 0; f < dim; ++f) NSB_CUDA(cudaMemsetAsync(vel(bf) + f * fs, 0, sizeof(double) * S->npts, st));
    }
    NSB_CHECK(nsb_sem_bdf_ext(S, W, bf, e1, e2, lag, o, 0, dim, kABn[o], kBDn[o], 1.0 / op->dt));
    // p* (extrapprp): p^(n-1), from the third step on 2 p^(n-1) - p^(n-2)
    lin2_kernel<<<g2, 256, 0, st>>>(prs(2), prs(0), o < 3 ? 1.0 : 2.0, prs(1), o < 3 ? 0.0 : -1.0, n2);
    ctx->launches++;
    // cresvipp + ophinv in the non-incremental form: H v* = QQ^T (bf + D^T p*)
    NSB_CHECK(launch_opgradt(S, prs(2), vel(bf), fs, 0, vel(bf), 1.0, 1.0, nullptr));
    NSB_CHECK(launch_gs_ext(S, vel(bf), dim, fs, 0, nullptr, 0.0, 0.0, nullptr));
    {
      int it[3] = {0, 0, 0};
      double res[3];
      NSB_CHECK(nsb_sem_hmholtz_vec(S, W, bf, W, vs, 0, dim, op->nu, bd0 / op->dt, op->tol, op->maxit, it, res));
      for (int g = 0; g < dim; ++g) op->helm_iters += it[g];
    }
    // incomprp: E dp = -(bd0/dt) D v* ; v = v* + (dt/bd0) B^-1 D^T dp ; p = p* + dp
    NSB_CHECK(launch_opdiv(S, vel(vs), fs, -bd0 / op->dt, prs(3), nullptr, nullptr, nullptr));
    {
      int it = 0;
      double res = 0.0;
      NSB_CHECK(esolve_d(S, prs(3), prs(4), op->tol_p, op->maxit, op->mean_free, op->precond, &it, &res));
      op->pres_iters += it;
    }
    NSB_CHECK(launch_binv_gradt(S, prs(4), vel(nw), fs, vel(vs), 1.0, op->dt / bd0, nullptr));
    NSB_CUDA(cudaMemcpyAsync(prs(1), prs(0), sizeof(double) * n2, cudaMemcpyDeviceToDevice, st));
    lin2_kernel<<<g2, 256, 0, st>>>(prs(0), prs(2), 1.0, prs(4), 1.0, n2);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    const int freed = lag[2];
    lag[2] = lag[1];
    lag[1] = lag[0];
    lag[0] = nw;
    nw = freed;
  }
  // fields outside the operator and %time are carried through
  NSB_CHECK(nsb_vec_copy(bout, cout, bin, cin));
  for (int f = 0; f < dim; ++f)
    NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->off[f], vel(lag[0]) + f * fs, sizeof(double) * S->npts,
                             cudaMemcpyDeviceToDevice, st));
  NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->off[dim], prs(0), sizeof(double) * n2, cudaMemcpyDeviceToDevice, st));
  return NSB_OK;
}
