// Time-stepper pieces either side of `ax`, on the device:
//   * dealiased convection of Nek5000's perturbation step  [UPSTREAM-RECALL convect.f:
//     set_dealias_rx, set_convect_new, convect_new, intp_rstd, grad_rst] -- the term advabp adds
//     for (U.grad)u' and (u'.grad)U, evaluated on the lxd = 3 lx1 / 2 Gauss-Legendre mesh;
//   * the EXT / BDF sums of the same step  [UPSTREAM-RECALL perturb.f: makextp, makebdfp] as one
//     fused streaming pass.
// Nek5000 is not vendored in the reference tree: the kernels are checked against a CPU restatement
// of those routines held by the test suite (tests/test_gpu_conv.py); parity unpinned.
#include <cmath>
#include <cstring>
#include <vector>

#include "nsb_internal.h"
#include "nsb_quadrature.h"
#include "nsb_device.cuh"

using namespace nsb;

namespace {

// ---- device: tensor-product interpolation of one element, whole CTA ------------------------------
// Shared layout of a work area (doubles): src [LX^3] | t1 [LX^2 LD] + t2 [LX LD^2] | fine [LD^3].
template <int LX, int LD>
struct ConvSmem {
  static constexpr int NC = LX * LX * LX, NF = LD * LD * LD;
  static constexpr int T1 = LX * LX * LD, T2 = LX * LD * LD;
  static constexpr int SRC = 0, TMP = NC, FINE = NC + T1 + T2, MAT = FINE + NF;
  static constexpr int TOTAL = MAT + LD * LX + LD * LD;   // + J + Dg
  static_assert(T1 + T2 >= NF, "the second fine buffer aliases the temporaries");
};

// fine[K][J][I] = sum_kji Jm[K][k] Jm[J][j] Jm[I][i] src[k][j][i]   (intp_rstd, idir = 0)
template <int LX, int LD>
__device__ __forceinline__ void interp3(const double *__restrict__ src, double *__restrict__ t1,
                                        double *__restrict__ t2, double *__restrict__ fine,
                                        const double *__restrict__ Jm) {
  for (int o = threadIdx.x; o < LX * LX * LD; o += blockDim.x) {     // t1[k][j][I]
    const int I = o % LD, kj = o / LD;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < LX; ++i) s = fma(Jm[I * LX + i], src[kj * LX + i], s);
    t1[o] = s;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < LX * LD * LD; o += blockDim.x) {     // t2[k][J][I]
    const int I = o % LD, Jx = (o / LD) % LD, k = o / (LD * LD);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < LX; ++j) s = fma(Jm[Jx * LX + j], t1[(k * LX + j) * LD + I], s);
    t2[o] = s;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < LD * LD * LD; o += blockDim.x) {     // fine[K][J][I]
    const int JI = o % (LD * LD), K = o / (LD * LD);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < LX; ++k) s = fma(Jm[K * LX + k], t2[k * LD * LD + JI], s);
    fine[o] = s;
  }
  __syncthreads();
}

template <int LX, int LD>
__device__ __forceinline__ void load_mats(double *sm, const double *__restrict__ Jg, const double *__restrict__ Dgg) {
  using L = ConvSmem<LX, LD>;
  for (int i = threadIdx.x; i < LD * LX; i += blockDim.x) sm[L::MAT + i] = Jg[i];
  for (int i = threadIdx.x; i < LD * LD; i += blockDim.x) sm[L::MAT + LD * LX + i] = Dgg[i];
}

// set_dealias_rx: rxf[a][e] = (w_I w_J w_K) * J(rst[a][e]);  grid (nel, 9)
template <int LX, int LD>
__global__ void __launch_bounds__(256)
dealias_rx_kernel(const double *__restrict__ rst, int64_t npts, int64_t nfine, const double *__restrict__ Jg,
                  const double *__restrict__ Dgg, const double *__restrict__ wd, double *__restrict__ rxf) {
  using L = ConvSmem<LX, LD>;
  extern __shared__ double sm[];
  const int64_t e = blockIdx.x;
  const int a = blockIdx.y;
  load_mats<LX, LD>(sm, Jg, Dgg);
  for (int i = threadIdx.x; i < L::NC; i += blockDim.x) sm[L::SRC + i] = rst[(int64_t)a * npts + e * L::NC + i];
  __syncthreads();
  interp3<LX, LD>(sm + L::SRC, sm + L::TMP, sm + L::TMP + L::T1, sm + L::FINE, sm + L::MAT);
  for (int o = threadIdx.x; o < L::NF; o += blockDim.x) {
    const int I = o % LD, Jx = (o / LD) % LD, K = o / (LD * LD);
    rxf[(int64_t)a * nfine + e * L::NF + o] = wd[I] * wd[Jx] * wd[K] * sm[L::FINE + o];
  }
}

// ---- line kernels: one thread owns a whole grid line of a contraction ---------------------------
// The first version of convect (one output value per thread, both operands of every FMA read from
// shared memory) ran at 4.7 TFLOP/s.  Here a thread loads the LX or LD values of one grid line into
// registers once and produces every output of that line from them; the small matrices J (GLL -> GL)
// and Dg (derivative on GL) sit in constant memory and, with the loops fully unrolled, become
// immediate constant operands of the DFMAs: no shared-memory load per FMA.  Lines run along i, then
// j, then k, so every stage changes line ownership through shared memory; the fastest dimension is
// padded by one double (stride LX+1 / LD+1) so that line reads along i are conflict-free.
// CTA = LD x LD threads (one per (J, I) fine-mesh column), one (element, field) per CTA, fields of an
// element on consecutive CTAs so that the contravariant field c is re-read from L2, not DRAM.
template <int LX, int LD>
__constant__ double c_convJ[LD * LX];    // J[I][i]
template <int LX, int LD>
__constant__ double c_convD[LD * LD];    // Dg[I][m]

template <int LX, int LD>
struct LineSmem {
  static constexpr int LXp = LX + 1, LDp = LD + 1, NT = LD * LD;
  static constexpr int NSRC = LX * LX * LXp, NA1 = LX * LX * LDp, NA2 = LX * LD * LDp, NUF = LD * LD * LDp;
  static constexpr int SRC = 0, TMP = NSRC, UF = NSRC + NA1 + NA2, TOTAL = UF + NUF;
  static_assert(NA1 + NA2 >= NUF, "wf aliases the interpolation temporaries");
  static_assert(NA2 <= NUF && NA1 <= NA1 + NA2, "projection temporaries fit");
};

template <int LX, int LD>
__device__ __forceinline__ void load_src(double *__restrict__ src, const double *__restrict__ g) {
  using L = LineSmem<LX, LD>;
  for (int o = threadIdx.x; o < LX * LX * LX; o += L::NT) src[(o / LX) * L::LXp + o % LX] = g[o];
}

// uf[K][J][I] = (J x J x J) src, line by line; on return thread (J, I) = (tid / LD, tid % LD) also
// holds its fine-mesh column uf[0..LD)[J][I] in col[].
template <int LX, int LD>
__device__ __forceinline__ void interp_lines(const double *__restrict__ src, double *__restrict__ a1,
                                             double *__restrict__ a2, double *__restrict__ uf, double (&col)[LD]) {
  using L = LineSmem<LX, LD>;
  const int tid = threadIdx.x;
  if (tid < LX * LX) {                                       // line (k, j) along i
    double in[LX];
#pragma unroll
    for (int i = 0; i < LX; ++i) in[i] = src[tid * L::LXp + i];
#pragma unroll
    for (int I = 0; I < LD; ++I) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < LX; ++i) s = fma(c_convJ<LX, LD>[I * LX + i], in[i], s);
      a1[tid * L::LDp + I] = s;
    }
  }
  __syncthreads();
  if (tid < LX * LD) {                                       // line (k, I) along j
    const int k = tid / LD, I = tid % LD;
    double in[LX];
#pragma unroll
    for (int j = 0; j < LX; ++j) in[j] = a1[(k * LX + j) * L::LDp + I];
#pragma unroll
    for (int Jx = 0; Jx < LD; ++Jx) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < LX; ++j) s = fma(c_convJ<LX, LD>[Jx * LX + j], in[j], s);
      a2[(k * LD + Jx) * L::LDp + I] = s;
    }
  }
  __syncthreads();
  {                                                          // line (J, I) along k
    const int Jx = tid / LD, I = tid % LD;
    double in[LX];
#pragma unroll
    for (int k = 0; k < LX; ++k) in[k] = a2[(k * LD + Jx) * L::LDp + I];
#pragma unroll
    for (int K = 0; K < LD; ++K) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < LX; ++k) s = fma(c_convJ<LX, LD>[K * LX + k], in[k], s);
      col[K] = s;
      uf[(K * LD + Jx) * L::LDp + I] = s;
    }
  }
  __syncthreads();
}

// set_convect_new: c_a = sum_b rxf[3a+b] * J(v_b);  one CTA per element
template <int LX, int LD>
__global__ void __launch_bounds__(LD * LD)
set_convect_kernel(const double *__restrict__ v, int64_t fstride, int64_t nfine, const double *__restrict__ rxf,
                   double *__restrict__ cf) {
  using L = LineSmem<LX, LD>;
  constexpr int NC = LX * LX * LX, NF = LD * LD * LD;
  extern __shared__ double sm[];
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x;                               // = J * LD + I
  double acc[3][LD];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int K = 0; K < LD; ++K) acc[a][K] = 0.0;
  for (int b = 0; b < 3; ++b) {
    load_src<LX, LD>(sm + L::SRC, v + (int64_t)b * fstride + e * NC);
    __syncthreads();
    double col[LD];
    interp_lines<LX, LD>(sm + L::SRC, sm + L::TMP, sm + L::TMP + L::NA1, sm + L::UF, col);
#pragma unroll
    for (int K = 0; K < LD; ++K) {
      const int64_t o = e * NF + K * LD * LD + tid;
#pragma unroll
      for (int a = 0; a < 3; ++a) acc[a][K] = fma(rxf[(int64_t)(3 * a + b) * nfine + o], col[K], acc[a][K]);
    }
  }
#pragma unroll
  for (int K = 0; K < LD; ++K) {
#pragma unroll
    for (int a = 0; a < 3; ++a) cf[(int64_t)a * nfine + e * NF + K * LD * LD + tid] = acc[a][K];
  }
}

// convect_new: out (+)= scale * J^T [ (c . grad_rst)(J u) ];  CTA = (element, field), field fastest
// ADJ: the exact transpose, out (+)= scale * J^T [ D_r^T (c_r o J u) + D_s^T (c_s o J u) + D_t^T (c_t o J u) ]
// (the factors are applied before the differentiation and the derivative matrix is read transposed) -- the
// convective term of the discrete adjoint of the time-stepper (exponential_prop%rmatvec,
// core/linear_operators.f90:84-103)
template <int LX, int LD, bool ADJ>
__global__ void __launch_bounds__(LD * LD)
convect_kernel(const double *__restrict__ u, double *__restrict__ out, int64_t fstride_in, int64_t fstride_out,
               int64_t nfine, int nf, const double *__restrict__ cf, double scale, int accumulate) {
  using L = LineSmem<LX, LD>;
  constexpr int NC = LX * LX * LX, NF = LD * LD * LD;
  extern __shared__ double sm[];
  const int64_t e = blockIdx.x / nf;
  const int f = (int)(blockIdx.x % nf);
  const int tid = threadIdx.x;
  double *src = sm + L::SRC, *uf = sm + L::UF, *wf = sm + L::TMP;   // wf aliases a1/a2 after the interpolation
  load_src<LX, LD>(src, u + (int64_t)f * fstride_in + e * NC);
  __syncthreads();
  double col[LD];
  interp_lines<LX, LD>(src, sm + L::TMP, sm + L::TMP + L::NA1, uf, col);
  const double *cr = cf + e * NF, *cs = cr + nfine, *ct = cs + nfine;
  {                                                          // r: line (K, J) = tid along I
    double in[LD];
#pragma unroll
    for (int I = 0; I < LD; ++I) in[I] = ADJ ? cr[tid * LD + I] * uf[tid * L::LDp + I] : uf[tid * L::LDp + I];
#pragma unroll
    for (int I = 0; I < LD; ++I) {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < LD; ++m) s = fma(c_convD<LX, LD>[ADJ ? m * LD + I : I * LD + m], in[m], s);
      wf[tid * L::LDp + I] = ADJ ? s : cr[tid * LD + I] * s;
    }
  }
  __syncthreads();
  {                                                          // s: line (K, I) along J
    const int K = tid / LD, I = tid % LD;
    double in[LD];
#pragma unroll
    for (int Jx = 0; Jx < LD; ++Jx)
      in[Jx] = ADJ ? cs[(K * LD + Jx) * LD + I] * uf[(K * LD + Jx) * L::LDp + I] : uf[(K * LD + Jx) * L::LDp + I];
#pragma unroll
    for (int Jx = 0; Jx < LD; ++Jx) {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < LD; ++m) s = fma(c_convD<LX, LD>[ADJ ? m * LD + Jx : Jx * LD + m], in[m], s);
      wf[(K * LD + Jx) * L::LDp + I] += ADJ ? s : cs[(K * LD + Jx) * LD + I] * s;
    }
  }
  __syncthreads();
  {                                                          // t: line (J, I) along K, from the registers
    const int Jx = tid / LD, I = tid % LD;
    double w[LD];
    if (ADJ) {
#pragma unroll
      for (int K = 0; K < LD; ++K) col[K] *= ct[(K * LD + Jx) * LD + I];
    }
#pragma unroll
    for (int K = 0; K < LD; ++K) {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < LD; ++m) s = fma(c_convD<LX, LD>[ADJ ? m * LD + K : K * LD + m], col[m], s);
      w[K] = wf[(K * LD + Jx) * L::LDp + I] + (ADJ ? s : ct[(K * LD + Jx) * LD + I] * s);
    }
    // project back along k at once: the thread owns the whole (J, I) column of w
    double *b2 = uf;                                         // [k][J][I], uf is dead (other threads' s/r reads are done)
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      double s = 0.0;
#pragma unroll
      for (int K = 0; K < LD; ++K) s = fma(c_convJ<LX, LD>[K * LX + k], w[K], s);
      b2[(k * LD + Jx) * L::LDp + I] = s;
    }
  }
  __syncthreads();
  double *b1 = wf;                                           // [k][j][I]; wf is dead
  if (tid < LX * LD) {                                       // line (k, I) along J
    const int k = tid / LD, I = tid % LD;
    double in[LD];
#pragma unroll
    for (int Jx = 0; Jx < LD; ++Jx) in[Jx] = uf[(k * LD + Jx) * L::LDp + I];
#pragma unroll
    for (int j = 0; j < LX; ++j) {
      double s = 0.0;
#pragma unroll
      for (int Jx = 0; Jx < LD; ++Jx) s = fma(c_convJ<LX, LD>[Jx * LX + j], in[Jx], s);
      b1[(k * LX + j) * L::LDp + I] = s;
    }
  }
  __syncthreads();
  if (tid < LX * LX) {                                       // line (k, j) along I
    double in[LD];
#pragma unroll
    for (int I = 0; I < LD; ++I) in[I] = b1[tid * L::LDp + I];
#pragma unroll
    for (int i = 0; i < LX; ++i) {
      double s = 0.0;
#pragma unroll
      for (int I = 0; I < LD; ++I) s = fma(c_convJ<LX, LD>[I * LX + i], in[I], s);
      src[tid * L::LXp + i] = s;
    }
  }
  __syncthreads();
  double *dst = out + (int64_t)f * fstride_out + e * NC;
  for (int o = tid; o < NC; o += L::NT) {
    const double s = src[(o / LX) * L::LXp + o % LX];
    dst[o] = accumulate ? fma(scale, s, dst[o]) : scale * s;
  }
}

// ---- 2-D (the reference's cylinder / backward-facing-step cases): runtime sizes, one output per
// thread; these meshes have a few thousand elements of 36 points, the kernels are not hot -----------
// shared layout: Jm [ld*lx] | Dg [ld*ld] | src [lx*lx] | t1 [lx*ld] | uf [ld*ld] | wf [ld*ld]
__device__ __forceinline__ void interp2(int lx, int ld, const double *Jm, const double *src, double *t1, double *uf) {
  for (int o = threadIdx.x; o < lx * ld; o += blockDim.x) {      // t1[j][I]
    const int I = o % ld, j = o / ld;
    double s = 0.0;
    for (int i = 0; i < lx; ++i) s = fma(Jm[I * lx + i], src[j * lx + i], s);
    t1[o] = s;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ld * ld; o += blockDim.x) {      // uf[J][I]
    const int I = o % ld, Jx = o / ld;
    double s = 0.0;
    for (int j = 0; j < lx; ++j) s = fma(Jm[Jx * lx + j], t1[j * ld + I], s);
    uf[o] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void load_mats2(int lx, int ld, double *sm, const double *__restrict__ Jg,
                                           const double *__restrict__ Dgg) {
  for (int i = threadIdx.x; i < ld * lx; i += blockDim.x) sm[i] = Jg[i];
  for (int i = threadIdx.x; i < ld * ld; i += blockDim.x) sm[ld * lx + i] = Dgg[i];
}

// mode 0: rxf[a][e] = (w_I w_J) J(rst[a][e]), grid (nel, 4)
// mode 1: c_a = sum_b rxf[2a+b] J(v_b), grid (nel, 1)
// mode 2: out (+)= scale J^T[(c . grad)(J u)], grid (nel, nf)
// mode 3: the transpose of mode 2, out (+)= scale J^T[D_r^T (c_r o J u) + D_s^T (c_s o J u)]
__global__ void __launch_bounds__(128)
conv2d_kernel(int mode, int lx, int ld, const double *__restrict__ in, double *__restrict__ out, int64_t npts,
              int64_t nfine, int64_t fs_in, int64_t fs_out, const double *__restrict__ Jg,
              const double *__restrict__ Dgg, const double *__restrict__ wd, const double *__restrict__ fine_in,
              double scale, int accumulate) {
  extern __shared__ double sm[];
  const int nc = lx * lx, nf2 = ld * ld;
  double *Jm = sm, *Dg = Jm + ld * lx, *src = Dg + ld * ld, *t1 = src + nc, *uf = t1 + lx * ld, *wf = uf + nf2;
  const int64_t e = blockIdx.x;
  const int y = blockIdx.y;
  load_mats2(lx, ld, sm, Jg, Dgg);
  if (mode == 0) {
    for (int i = threadIdx.x; i < nc; i += blockDim.x) src[i] = in[(int64_t)y * npts + e * nc + i];
    __syncthreads();
    interp2(lx, ld, Jm, src, t1, uf);
    for (int o = threadIdx.x; o < nf2; o += blockDim.x)
      out[(int64_t)y * nfine + e * nf2 + o] = wd[o % ld] * wd[o / ld] * uf[o];
    return;
  }
  if (mode == 1) {
    for (int b = 0; b < 2; ++b) {
      __syncthreads();
      for (int i = threadIdx.x; i < nc; i += blockDim.x) src[i] = in[(int64_t)b * fs_in + e * nc + i];
      __syncthreads();
      interp2(lx, ld, Jm, src, t1, uf);
      for (int o = threadIdx.x; o < nf2; o += blockDim.x) {
        const int64_t q = e * nf2 + o;
        for (int a = 0; a < 2; ++a) {
          const double t = fine_in[(int64_t)(2 * a + b) * nfine + q] * uf[o];
          out[(int64_t)a * nfine + q] = b == 0 ? t : out[(int64_t)a * nfine + q] + t;
        }
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < nc; i += blockDim.x) src[i] = in[(int64_t)y * fs_in + e * nc + i];
  __syncthreads();
  interp2(lx, ld, Jm, src, t1, uf);
  const double *cr = fine_in + e * nf2, *cs = cr + nfine;
  for (int o = threadIdx.x; o < nf2; o += blockDim.x) {
    const int I = o % ld, Jx = o / ld;
    double ur = 0.0, us = 0.0;
    if (mode == 3) {
      for (int m = 0; m < ld; ++m) {
        ur = fma(Dg[m * ld + I], cr[Jx * ld + m] * uf[Jx * ld + m], ur);
        us = fma(Dg[m * ld + Jx], cs[m * ld + I] * uf[m * ld + I], us);
      }
      wf[o] = ur + us;
    } else {
      for (int m = 0; m < ld; ++m) {
        ur = fma(Dg[I * ld + m], uf[Jx * ld + m], ur);
        us = fma(Dg[Jx * ld + m], uf[m * ld + I], us);
      }
      wf[o] = cr[o] * ur + cs[o] * us;
    }
  }
  __syncthreads();
  double *p1 = uf;                                               // [J][i]
  for (int o = threadIdx.x; o < ld * lx; o += blockDim.x) {
    const int i = o % lx, Jx = o / lx;
    double s = 0.0;
    for (int I = 0; I < ld; ++I) s = fma(Jm[I * lx + i], wf[Jx * ld + I], s);
    p1[o] = s;
  }
  __syncthreads();
  double *dst = out + (int64_t)y * fs_out + e * nc;
  for (int o = threadIdx.x; o < nc; o += blockDim.x) {
    const int i = o % lx, j = o / lx;
    double s = 0.0;
    for (int Jx = 0; Jx < ld; ++Jx) s = fma(Jm[Jx * lx + j], p1[Jx * lx + i], s);
    dst[o] = accumulate ? fma(scale, s, dst[o]) : scale * s;
  }
}

inline size_t conv2d_smem(int lx, int ld) {
  return sizeof(double) * ((size_t)ld * lx + (size_t)ld * ld + (size_t)lx * lx + (size_t)lx * ld + 2 * (size_t)ld * ld);
}

// makextp + makebdfp, one pass; grid-stride over the points of one field, grid.y = field
__global__ void __launch_bounds__(256)
bdf_ext_kernel(double *__restrict__ bf, double *__restrict__ e1, double *__restrict__ e2, const double *v0,
               const double *v1, const double *v2, const double *__restrict__ bm1, int64_t npts, int64_t fstride,
               double ab0, double ab1, double ab2, double bd1, double bd2, double bd3, double rho_dt) {
  const int64_t off = (int64_t)blockIdx.y * fstride;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = off + p;
    const double b = bf[q], x1 = e1[q], x2 = e2[q];
    const double ta = ab1 * x1 + ab2 * x2;
    e2[q] = x1;
    e1[q] = b;
    double r = ab0 * b + ta;
    const double m = bm1[p];
    double tb = bd1 * m * v0[q];
    if (v1) tb = tb + bd2 * m * v1[q];
    if (v2) tb = tb + bd3 * m * v2[q];
    bf[q] = r + rho_dt * tb;
  }
}

int conv_field_ptr(nsb_sem_t S, nsb_basis_t B, int col, int field, int nf, double **out, int64_t *fstride,
                   const char *who) {
  NSB_REQUIRE(S && B, "%s: NULL argument", who);
  NSB_REQUIRE(col >= 0 && col < B->ncols, "%s: column %d out of range", who, col);
  nsb_layout_t L = B->lay;
  NSB_REQUIRE(field >= 0 && nf >= 1 && field + nf <= L->nfields, "%s: fields %d..%d out of range", who, field,
              field + nf - 1);
  NSB_REQUIRE(L->ctx == S->ctx, "%s: basis and mesh live on different contexts", who);
  for (int f = field; f < field + nf; ++f)
    NSB_REQUIRE(L->len[f] == S->npts, "%s: field %d has %lld points, mesh has %lld", who, f, (long long)L->len[f],
                (long long)S->npts);
  for (int f = field + 1; f < field + nf; ++f)
    NSB_REQUIRE(L->off[f] - L->off[f - 1] == L->off[field + 1] - L->off[field], "%s: fields are not equally spaced",
                who);
  *out = B->col(col) + L->off[field];
  *fstride = nf > 1 ? L->off[field + 1] - L->off[field] : 0;
  return NSB_OK;
}

template <int LX, int LD>
int dealias_setup_t(nsb_sem_t S, const double *wd_d, const double *J_h, const double *Dg_h) {
  using L = ConvSmem<LX, LD>;
  using LL = LineSmem<LX, LD>;
  const size_t smem = sizeof(double) * L::TOTAL, smem_l = sizeof(double) * LL::TOTAL;
  const int64_t nfine = S->nel * L::NF;
  NSB_CUDA(cudaMemcpyToSymbol(c_convJ<LX, LD>, J_h, sizeof(double) * LD * LX));
  NSB_CUDA(cudaMemcpyToSymbol(c_convD<LX, LD>, Dg_h, sizeof(double) * LD * LD));
  NSB_CUDA(cudaFuncSetAttribute(dealias_rx_kernel<LX, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NSB_CUDA(cudaFuncSetAttribute(set_convect_kernel<LX, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
  NSB_CUDA(cudaFuncSetAttribute(convect_kernel<LX, LD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
  NSB_CUDA(cudaFuncSetAttribute(convect_kernel<LX, LD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
  dealias_rx_kernel<LX, LD><<<dim3((unsigned)S->nel, 9), 256, smem, S->ctx->stream>>>(S->rst_d, S->npts, nfine, S->J_d,
                                                                                  S->Dg_d, wd_d, S->rxf_d);
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

template <int LX, int LD>
int set_convect_t(nsb_sem_t S, const double *v, int64_t fstride, double *cf) {
  using LL = LineSmem<LX, LD>;
  set_convect_kernel<LX, LD><<<(unsigned)S->nel, LL::NT, sizeof(double) * LL::TOTAL, S->ctx->stream>>>(
      v, fstride, S->nel * (int64_t)(LD * LD * LD), S->rxf_d, cf);
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

template <int LX, int LD>
int convect_t(nsb_sem_t S, const double *u, double *out, int64_t fsi, int64_t fso, int nf, const double *cf,
              double scale, int accumulate, bool adj) {
  using LL = LineSmem<LX, LD>;
  if (adj)
    convect_kernel<LX, LD, true><<<(unsigned)(S->nel * nf), LL::NT, sizeof(double) * LL::TOTAL, S->ctx->stream>>>(
        u, out, fsi, fso, S->nel * (int64_t)(LD * LD * LD), nf, cf, scale, accumulate);
  else
    convect_kernel<LX, LD, false><<<(unsigned)(S->nel * nf), LL::NT, sizeof(double) * LL::TOTAL, S->ctx->stream>>>(
        u, out, fsi, fso, S->nel * (int64_t)(LD * LD * LD), nf, cf, scale, accumulate);
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// (lx, lxd) pairs with compiled kernels: Nek's lxd = 3 lx1 / 2 for lx1 = 4, 6, 8 (+ lx1 = 5 with lxd = 8)
#define NSB_CONV_DISPATCH(S, CALL)                                                       \
  do {                                                                                   \
    if ((S)->lx == 8 && (S)->lxd == 12) return CALL(8, 12);                              \
    if ((S)->lx == 6 && (S)->lxd == 9) return CALL(6, 9);                                \
    if ((S)->lx == 5 && (S)->lxd == 8) return CALL(5, 8);                                \
    if ((S)->lx == 4 && (S)->lxd == 6) return CALL(4, 6);                                \
    set_error("dealiased convection: no kernel for lx1=%d, lxd=%d", (S)->lx, (S)->lxd);  \
    return NSB_EINVAL;                                                                   \
  } while (0)

}  // namespace

// Host only (no CUDA): the small matrices of the dealiasing operators, for checking them without a device.
extern "C" int nsb_dealias_matrices(int N, int lxd, double *zd, double *wd, double *J, double *Dg) {
  NSB_REQUIRE(N >= 1 && N <= 31 && lxd >= 1 && lxd <= 64, "nsb_dealias_matrices: N=%d, lxd=%d out of range", N, lxd);
  std::vector<double> zg(N + 1), z, w;
  NSB_CHECK(nsb_gll(N, zg.data(), nullptr, nullptr));
  gauss_legendre(lxd, z, w);
  if (zd) memcpy(zd, z.data(), sizeof(double) * lxd);
  if (wd) memcpy(wd, w.data(), sizeof(double) * lxd);
  if (J) {
    const std::vector<double> Jm = interp_matrix(zg, z);      // [lxd][N+1], row-major
    memcpy(J, Jm.data(), sizeof(double) * Jm.size());
  }
  if (Dg) {
    const std::vector<double> D = deriv_matrix(z);            // [lxd][lxd], row-major
    memcpy(Dg, D.data(), sizeof(double) * D.size());
  }
  return NSB_OK;
}

extern "C" int nsb_sem_dealias_setup(nsb_sem_t S, int lxd) {
  NSB_REQUIRE(S, "nsb_sem_dealias_setup: NULL");
  if (lxd <= 0) lxd = (S->dim == 3 && S->lx == 5) ? 8 : 3 * S->lx / 2;
  NSB_REQUIRE(S->dim == 3 || (lxd >= S->lx && lxd <= 24), "nsb_sem_dealias_setup: lxd=%d out of range", lxd);
  cudaSetDevice(S->ctx->device);
  if (S->lxd == lxd && S->rxf_d) return NSB_OK;
  const int lx = S->lx;
  std::vector<double> zd, wd;
  gauss_legendre(lxd, zd, wd);
  const std::vector<double> J = interp_matrix(S->z_h, zd), Dg = deriv_matrix(zd);
  for (double **p : {&S->J_d, &S->Dg_d, &S->rxf_d, &S->cfine_d[0], &S->cfine_d[1]}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  S->lxd = lxd;
  const int64_t nfine = S->nel * (int64_t)lxd * lxd * (S->dim == 3 ? lxd : 1);
  double *wd_d = nullptr;
  NSB_CUDA(cudaMalloc(&S->J_d, sizeof(double) * lxd * lx));
  NSB_CUDA(cudaMalloc(&S->Dg_d, sizeof(double) * lxd * lxd));
  NSB_CUDA(cudaMalloc(&wd_d, sizeof(double) * lxd));
  NSB_CUDA(cudaMalloc(&S->rxf_d, sizeof(double) * S->dim * S->dim * nfine));
  NSB_CUDA(cudaMemcpy(S->J_d, J.data(), sizeof(double) * lxd * lx, cudaMemcpyHostToDevice));
  NSB_CUDA(cudaMemcpy(S->Dg_d, Dg.data(), sizeof(double) * lxd * lxd, cudaMemcpyHostToDevice));
  NSB_CUDA(cudaMemcpy(wd_d, wd.data(), sizeof(double) * lxd, cudaMemcpyHostToDevice));
  int rc;
  if (S->dim == 2) {
    conv2d_kernel<<<dim3((unsigned)S->nel, 4), 128, conv2d_smem(lx, lxd), S->ctx->stream>>>(
        0, lx, lxd, S->rst_d, S->rxf_d, S->npts, nfine, 0, 0, S->J_d, S->Dg_d, wd_d, nullptr, 1.0, 0);
    S->ctx->launches++;
    rc = cudaGetLastError() == cudaSuccess ? NSB_OK : NSB_ECUDA;
    if (rc != NSB_OK) set_error("nsb_sem_dealias_setup: 2-D metric kernel failed to launch");
  } else {
    auto run = [&]() -> int {
#define CALL_SETUP(A, B) dealias_setup_t<A, B>(S, wd_d, J.data(), Dg.data())
      NSB_CONV_DISPATCH(S, CALL_SETUP);
#undef CALL_SETUP
    };
    rc = run();
  }
  cudaStreamSynchronize(S->ctx->stream);
  cudaFree(wd_d);
  if (rc != NSB_OK) {
    S->lxd = 0;
    return rc;
  }
  return NSB_OK;
}

extern "C" int nsb_sem_set_convect(nsb_sem_t S, int slot, nsb_basis_t B, int col, int field0) {
  NSB_REQUIRE(S && B, "nsb_sem_set_convect: NULL argument");
  NSB_REQUIRE(slot == 0 || slot == 1, "nsb_sem_set_convect: slot %d (0 or 1)", slot);
  NSB_REQUIRE(S->lxd > 0 && S->rxf_d, "nsb_sem_set_convect: call nsb_sem_dealias_setup first");
  double *v;
  int64_t fs;
  NSB_CHECK(conv_field_ptr(S, B, col, field0, S->dim, &v, &fs, "nsb_sem_set_convect"));
  cudaSetDevice(S->ctx->device);
  const int64_t nfine = S->nel * (int64_t)S->lxd * S->lxd * (S->dim == 3 ? S->lxd : 1);
  if (!S->cfine_d[slot]) NSB_CUDA(cudaMalloc(&S->cfine_d[slot], sizeof(double) * S->dim * nfine));
  double *cf = S->cfine_d[slot];
  if (S->dim == 2) {
    conv2d_kernel<<<dim3((unsigned)S->nel, 1), 128, conv2d_smem(S->lx, S->lxd), S->ctx->stream>>>(
        1, S->lx, S->lxd, v, cf, S->npts, nfine, fs, 0, S->J_d, S->Dg_d, nullptr, S->rxf_d, 1.0, 0);
    S->ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
#define CALL_SC(A, B) set_convect_t<A, B>(S, v, fs, cf)
  NSB_CONV_DISPATCH(S, CALL_SC);
#undef CALL_SC
}

static int convect_impl(nsb_sem_t S, int slot, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout, int field0, int nf,
                        double scale, int accumulate, bool adj);

extern "C" int nsb_sem_convect(nsb_sem_t S, int slot, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout,
                               int field0, int nf, double scale, int accumulate) {
  return convect_impl(S, slot, bin, cin, bout, cout, field0, nf, scale, accumulate, false);
}

// The exact transpose of nsb_sem_convect (as a matrix on the local points of a field): for all u, v
//   sum_p v_p (C u)_p = sum_p u_p (C^T v)_p .   It is the convective term of the discrete adjoint stepper.
extern "C" int nsb_sem_convect_t(nsb_sem_t S, int slot, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout,
                                 int field0, int nf, double scale, int accumulate) {
  return convect_impl(S, slot, bin, cin, bout, cout, field0, nf, scale, accumulate, true);
}

static int convect_impl(nsb_sem_t S, int slot, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout, int field0, int nf,
                        double scale, int accumulate, bool adj) {
  NSB_REQUIRE(S && bin && bout, "nsb_sem_convect: NULL argument");
  NSB_REQUIRE(slot == 0 || slot == 1, "nsb_sem_convect: slot %d (0 or 1)", slot);
  NSB_REQUIRE(S->lxd > 0 && S->cfine_d[slot], "nsb_sem_convect: slot %d has no convecting field (nsb_sem_set_convect)",
              slot);
  double *u, *w;
  int64_t fsi, fso;
  NSB_CHECK(conv_field_ptr(S, bin, cin, field0, nf, &u, &fsi, "nsb_sem_convect"));
  NSB_CHECK(conv_field_ptr(S, bout, cout, field0, nf, &w, &fso, "nsb_sem_convect"));
  NSB_REQUIRE(u != w, "nsb_sem_convect: input and output are the same vector");
  cudaSetDevice(S->ctx->device);
  const double *cf = S->cfine_d[slot];
  if (S->dim == 2) {
    conv2d_kernel<<<dim3((unsigned)S->nel, nf), 128, conv2d_smem(S->lx, S->lxd), S->ctx->stream>>>(
        adj ? 3 : 2, S->lx, S->lxd, u, w, S->npts, S->nel * (int64_t)S->lxd * S->lxd, fsi, fso, S->J_d, S->Dg_d, nullptr, cf,
        scale, accumulate);
    S->ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
#define CALL_CV(A, B) convect_t<A, B>(S, u, w, fsi, fso, nf, cf, scale, accumulate, adj)
  NSB_CONV_DISPATCH(S, CALL_CV);
#undef CALL_CV
}

extern "C" int nsb_sem_bdf_ext(nsb_sem_t S, nsb_basis_t B, int col_bf, int col_e1, int col_e2, const int *col_vlag,
                               int nbd, int field0, int nf, const double *ab, const double *bd,
                               double rho_over_dt) {
  NSB_REQUIRE(S && B && col_vlag && ab && bd, "nsb_sem_bdf_ext: NULL argument");
  NSB_REQUIRE(nbd >= 1 && nbd <= 3, "nsb_sem_bdf_ext: nbd=%d (1..3)", nbd);
  double *bf, *e1, *e2, *v[3] = {nullptr, nullptr, nullptr};
  int64_t fs, fs2;
  NSB_CHECK(conv_field_ptr(S, B, col_bf, field0, nf, &bf, &fs, "nsb_sem_bdf_ext"));
  NSB_CHECK(conv_field_ptr(S, B, col_e1, field0, nf, &e1, &fs2, "nsb_sem_bdf_ext"));
  NSB_CHECK(conv_field_ptr(S, B, col_e2, field0, nf, &e2, &fs2, "nsb_sem_bdf_ext"));
  for (int i = 0; i < nbd; ++i) NSB_CHECK(conv_field_ptr(S, B, col_vlag[i], field0, nf, &v[i], &fs2, "nsb_sem_bdf_ext"));
  NSB_REQUIRE(col_bf != col_e1 && col_bf != col_e2 && col_e1 != col_e2, "nsb_sem_bdf_ext: bf, e1, e2 must differ");
  for (int i = 0; i < nbd; ++i)
    NSB_REQUIRE(col_vlag[i] != col_bf && col_vlag[i] != col_e1 && col_vlag[i] != col_e2,
                "nsb_sem_bdf_ext: a velocity column aliases bf / e1 / e2");
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  // bf, e1, e2 read + written, nbd velocities and bm1 read
  ProfScope ps(ctx, PC_BLAS1, 8.0 * (double)S->npts * (nf * (6.0 + nbd) + 1.0));
  bdf_ext_kernel<<<dim3(ctx->num_sms * 8, nf), 256, 0, ctx->stream>>>(bf, e1, e2, v[0], v[1], v[2], S->bm1_d, S->npts, fs,
                                                                     ab[0], ab[1], ab[2], bd[1], nbd > 1 ? bd[2] : 0.0,
                                                                     nbd > 2 ? bd[3] : 0.0, rho_over_dt);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Device time-stepper operator: the structure of exponential_prop%matvec
// (core/linear_operators.f90:225-274 -- integrate the linearised equations over tau from a cold start,
// return the final state) for Nek's scalar step cdscal [UPSTREAM-RECALL]:
//     per step: bq = -rho (U.grad) T  ->  makeabq / makebdq (EXT/BDF, order ramp 1, 2, 3)
//               -> dssum -> hmholtz (kappa A + rho bd1/dt B) T_new = bq
// applied independently to the first nfields_apply fields.  Every piece is a kernel sequence above /
// in nsb_sem.cu; only the k = O(10) scalars of the CG recurrences visit the host.
// ------------------------------------------------------------------------------------------------
namespace {
const double kBD[4][4] = {{0, 0, 0, 0}, {1.0, 1.0, 0, 0}, {1.5, 2.0, -0.5, 0}, {11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0}};
const double kAB[4][3] = {{0, 0, 0}, {1.0, 0, 0}, {2.0, -1.0, 0}, {3.0, -3.0, 1.0}};
}  // namespace

extern "C" int nsb_op_create_stepper(nsb_sem_t S, nsb_layout_t layout, int nfields_apply, int slot, double kappa,
                                     double rho, double dt, int nsteps, double tol, int maxit, nsb_op_t *out) {
  NSB_REQUIRE(S && layout && out, "nsb_op_create_stepper: NULL argument");
  NSB_REQUIRE(nfields_apply >= 1 && nfields_apply <= layout->nfields, "nsb_op_create_stepper: nfields_apply=%d",
              nfields_apply);
  NSB_REQUIRE(slot >= -1 && slot <= 1, "nsb_op_create_stepper: slot %d (-1: no convection, 0, 1)", slot);
  NSB_REQUIRE(slot < 0 || (S->lxd > 0 && S->cfine_d[slot]),
              "nsb_op_create_stepper: slot %d has no convecting field (nsb_sem_dealias_setup, nsb_sem_set_convect)", slot);
  NSB_REQUIRE(kappa > 0.0 && rho > 0.0 && dt > 0.0 && nsteps >= 1 && maxit >= 1 && tol > 0.0,
              "nsb_op_create_stepper: bad parameter");
  NSB_REQUIRE(S->exchange_ready, "nsb_op_create_stepper: call nsb_sem_setup_exchange first");
  NSB_REQUIRE(layout->ctx == S->ctx, "nsb_op_create_stepper: layout and mesh live on different contexts");
  for (int f = 0; f < nfields_apply; ++f)
    NSB_REQUIRE(layout->len[f] == S->npts, "nsb_op_create_stepper: field %d has %lld points, mesh has %lld", f,
                (long long)layout->len[f], (long long)S->npts);
  nsb_op_t op = new nsb_op_s();
  op->kind = 3;
  op->sem = S;
  op->lay = layout;
  op->nfields_apply = nfields_apply;
  op->slot = slot;
  op->kappa = kappa;
  op->rho = rho;
  op->dt = dt;
  op->nsteps = nsteps;
  op->tol = tol;
  op->maxit = maxit;
  int r = nsb_basis_create(layout, 7, &op->tmp);
  if (r != NSB_OK) {
    delete op;
    return r;
  }
  *out = op;
  return NSB_OK;
}

extern "C" int nsb_op_create_stepper_adjoint(nsb_sem_t S, nsb_layout_t layout, int nfields_apply, int slot,
                                             double kappa, double rho, double dt, int nsteps, double tol, int maxit,
                                             nsb_op_t *out) {
  NSB_CHECK(nsb_op_create_stepper(S, layout, nfields_apply, slot, kappa, rho, dt, nsteps, tol, maxit, out));
  (*out)->adjoint = true;
  return NSB_OK;
}

namespace {
// y += a ct + b bm1 g   on nf equally spaced fields
__global__ void __launch_bounds__(256)
adj_accum_kernel(double *__restrict__ y, const double *__restrict__ ct, const double *__restrict__ g,
                 const double *__restrict__ bm1, int64_t npts, int64_t fstride, double a, double b) {
  const int64_t off = (int64_t)blockIdx.y * fstride;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = off + p;
    y[q] = y[q] + a * (ct ? ct[q] : 0.0) + b * bm1[p] * g[q];
  }
}

// out = c1 * c2 * in  (pointwise, c1 / c2 mesh arrays)
__global__ void __launch_bounds__(256)
scale2_kernel(double *__restrict__ out, const double *__restrict__ in, const double *__restrict__ c1,
              const double *__restrict__ c2, int64_t npts, int64_t fs_out, int64_t fs_in) {
  const int64_t oo = (int64_t)blockIdx.y * fs_out, oi = (int64_t)blockIdx.y * fs_in;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x)
    out[oo + p] = (c1 ? c1[p] : 1.0) * (c2 ? c2[p] : 1.0) * in[oi + p];
}
}  // namespace

// The DISCRETE adjoint of stepper_apply with respect to the BM1 inner product: for continuous masked u, v
//     <A u, v>_B = <u, A^+ v>_B        (to rounding and the Helmholtz tolerance),
// A^+ = B_a^-1 A^T B_a.  With x^n = S_n [ sum_j ab_j E^(n-1-j) + (rho/dt) B sum_i bd_(i+1) x^(n-1-i) ],
// E^m = -rho C x^m, S_n = hmholtz o dssum (symmetric), the transposed recurrence runs backwards in time on dual
// vectors y^m kept in local right-hand-side form:
//     y^N = B v ;  for n = N .. 1:  g = S_n y^n ;  y^(n-1-j) += ab_j (-rho C^T g) + (rho/dt) bd_(j+1) B g  (j < order_n)
//     A^+ v = binvm1 mask dssum(y^0).
// Same kernels as the forward step plus the transposed convection (exponential_prop%rmatvec,
// core/linear_operators.f90:84-103, for Nek's scalar step).
static int stepper_apply_adjoint(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  nsb_sem_t S = op->sem;
  nsb_basis_t W = op->tmp;
  nsb_layout_t L = op->lay;
  nsb_context_t ctx = L->ctx;
  const int nfa = op->nfields_apply;
  const int N = op->nsteps;
  const int64_t fs = nfa > 1 ? L->off[1] - L->off[0] : 0;
  for (int f = 1; f < nfa; ++f) NSB_REQUIRE(L->off[f] - L->off[f - 1] == fs, "stepper adjoint: fields are not equally spaced");
  for (int c = 0; c < 7; ++c) NSB_CHECK(nsb_vec_zero(W, c));
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const dim3 grid(ctx->num_sms * 8, nfa);
  // columns: y[0..3] rolling duals (y[n mod 4] = y^n), 4 = right-hand side / g, 5 = C^T g, 6 = solution
  const int crhs = 4, cct = 5, cg = 6;
  auto ycol = [&](int n) { return ((n % 4) + 4) % 4; };
  scale2_kernel<<<grid, 256, 0, st>>>(W->col(ycol(N)) + L->off[0], bin->col(cin) + L->off[0], S->bm1_d, nullptr, S->npts, fs, fs);
  ctx->launches++;
  for (int n = N; n >= 1; --n) {
    const int o = n < 3 ? n : 3;
    NSB_CHECK(nsb_vec_copy(W, crhs, W, ycol(n)));
    for (int f = 0; f < nfa; ++f) NSB_CHECK(nsb_sem_dssum(S, W, crhs, f));
    for (int f = 0; f < nfa; f += 3) {
      const int nb3 = nfa - f < 3 ? nfa - f : 3;
      int it[3] = {0, 0, 0};
      double res[3];
      NSB_CHECK(nsb_sem_hmholtz_vec(S, W, crhs, W, cg, f, nb3, op->kappa, op->rho * kBD[o][0] / op->dt, op->tol, op->maxit,
                                    it, res));
      for (int g = 0; g < nb3; ++g) op->helm_iters += it[g];
    }
    NSB_CHECK(nsb_vec_zero(W, ycol(n)));                    // y^n is consumed; the slot becomes y^(n-4)
    if (op->slot >= 0) NSB_CHECK(nsb_sem_convect_t(S, op->slot, W, cg, W, cct, 0, nfa, -op->rho, 0));
    for (int j = 0; j < o; ++j) {
      const int m = n - 1 - j;
      if (m < 0) continue;                                   // cold start: x^m = 0 for m < 0
      adj_accum_kernel<<<grid, 256, 0, st>>>(W->col(ycol(m)) + L->off[0], op->slot >= 0 ? W->col(cct) + L->off[0] : nullptr,
                                             W->col(cg) + L->off[0], S->bm1_d, S->npts, fs, kAB[o][j],
                                             op->rho / op->dt * kBD[o][j + 1]);
      ctx->launches++;
    }
    NSB_CUDA(cudaGetLastError());
  }
  NSB_CHECK(nsb_vec_copy(W, crhs, W, ycol(0)));
  for (int f = 0; f < nfa; ++f) NSB_CHECK(nsb_sem_dssum(S, W, crhs, f));
  NSB_CHECK(nsb_vec_copy(bout, cout, bin, cin));             // fields outside the operator and %time carried through
  scale2_kernel<<<grid, 256, 0, st>>>(bout->col(cout) + L->off[0], W->col(crhs) + L->off[0], S->binv_d, S->mask_d, S->npts, fs,
                                      fs);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

int nsb::stepper_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  NSB_REQUIRE(bin->lay == op->lay, "nsb_op_apply: time-stepper operator built for another layout");
  if (op->adjoint) return stepper_apply_adjoint(op, bin, cin, bout, cout);
  nsb_sem_t S = op->sem;
  nsb_basis_t W = op->tmp;
  nsb_layout_t L = op->lay;
  const int nfa = op->nfields_apply;
  for (int c = 0; c < 7; ++c) NSB_CHECK(nsb_vec_zero(W, c));
  NSB_CHECK(nsb_vec_copy(W, 0, bin, cin));
  int lag[3] = {0, 1, 2}, nw = 6;
  const int bq = 3, e1 = 4, e2 = 5;
  for (int n = 1; n <= op->nsteps; ++n) {
    const int o = n < 3 ? n : 3;
    if (op->slot >= 0) NSB_CHECK(nsb_sem_convect(S, op->slot, W, lag[0], W, bq, 0, nfa, -op->rho, 0));
    else NSB_CHECK(nsb_vec_zero(W, bq));
    NSB_CHECK(nsb_sem_bdf_ext(S, W, bq, e1, e2, lag, o, 0, nfa, kAB[o], kBD[o], op->rho / op->dt));
    for (int f = 0; f < nfa; ++f) NSB_CHECK(nsb_sem_dssum(S, W, bq, f));
    for (int f = 0; f < nfa; f += 3) {                       // up to three systems side by side
      const int nb3 = nfa - f < 3 ? nfa - f : 3;
      int it[3] = {0, 0, 0};
      double res[3];
      NSB_CHECK(nsb_sem_hmholtz_vec(S, W, bq, W, nw, f, nb3, op->kappa, op->rho * kBD[o][0] / op->dt, op->tol, op->maxit,
                                    it, res));
      for (int g = 0; g < nb3; ++g) op->helm_iters += it[g];
    }
    const int freed = lag[2];
    lag[2] = lag[1];
    lag[1] = lag[0];
    lag[0] = nw;
    nw = freed;
  }
  // fields outside the operator and %time are carried through; the stepped fields come from the last lag
  NSB_CHECK(nsb_vec_copy(bout, cout, bin, cin));
  cudaStream_t s = L->ctx->stream;
  for (int f = 0; f < nfa; ++f)
    NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->off[f], W->col(lag[0]) + L->off[f], sizeof(double) * L->len[f],
                             cudaMemcpyDeviceToDevice, s));
  return NSB_OK;
}
