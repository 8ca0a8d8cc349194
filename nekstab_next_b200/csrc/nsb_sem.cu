// Spectral-element operator kernels: geometry, axhelm (tensor-product sum factorisation with
// geometric factors G1..G6 on GLL points), dssum (gather-scatter), col2, and the fused linear
// operator used as the Arnoldi matvec on the synthetic configurations.
//
// Reference: these are the Nek5000 kernels the reference reaches through nek_advance
// (core/linear_operators.f90:247, core/matvec.f90:211); in-tree uses of dssum/col2:
// core/utils.f90:287-290, 338-346.  [UPSTREAM-RECALL] hmholtz.f axhelm, navier5.f local_grad3,
// coef.f glmapm1/geodat1, speclib.f zwgll/dgll, gslib gs_op.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "nsb_internal.h"
#include "nsb_device.cuh"

using namespace nsb;

namespace nsb {
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nblk, int pstride, int k,
                                       double *__restrict__ out, int accumulate_into, double *__restrict__ out2);
}


// ------------------------------------------------------------------------------------------------
// GLL quadrature (host)
// ------------------------------------------------------------------------------------------------
static void legendre(int n, double x, double *pn, double *pnm1) {
  double p0 = 1.0, p1 = x;
  if (n == 0) { *pn = 1.0; *pnm1 = 0.0; return; }
  for (int m = 1; m < n; ++m) {
    double p2 = ((2 * m + 1) * x * p1 - m * p0) / (m + 1);
    p0 = p1;
    p1 = p2;
  }
  *pn = p1;
  *pnm1 = p0;
}

extern "C" int nsb_gll(int N, double *z, double *w, double *D) {
  NSB_REQUIRE(N >= 1 && N <= 31, "nsb_gll: N=%d out of range", N);
  const int n1 = N + 1;
  std::vector<double> x(n1), pn(n1);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n1; ++i) x[i] = -std::cos(pi * i / N);
  for (int i = 1; i < N; ++i) {
    double xi = x[i];
    for (int it = 0; it < 100; ++it) {
      double p, pm;
      legendre(N, xi, &p, &pm);
      double q = N * (pm - xi * p);             // (1-x^2) P_N'(x)
      double dq = -(double)N * (N + 1) * p;     // derivative of q
      double dx = q / dq;
      xi -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    x[i] = xi;
  }
  x[0] = -1.0;
  x[N] = 1.0;
  for (int i = 0; i < n1 / 2; ++i) {  // enforce antisymmetry
    double a = 0.5 * (x[i] - x[N - i]);
    x[i] = a;
    x[N - i] = -a;
  }
  if (n1 % 2) x[N / 2] = 0.0;
  for (int i = 0; i < n1; ++i) {
    double pm;
    legendre(N, x[i], &pn[i], &pm);
  }
  for (int i = 0; i < n1; ++i) {
    if (z) z[i] = x[i];
    if (w) w[i] = 2.0 / (N * (N + 1) * pn[i] * pn[i]);
  }
  if (D) {
    for (int j = 0; j < n1; ++j)
      for (int i = 0; i < n1; ++i)
        D[i + n1 * j] = (i == j) ? 0.0 : pn[i] / (pn[j] * (x[i] - x[j]));
    D[0] = -N * (N + 1) / 4.0;
    D[N + n1 * N] = N * (N + 1) / 4.0;
  }
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------------
namespace {

// Geometry: one thread per point, D read through the read-only cache (setup only, not hot).
__global__ void geom3d_kernel(const double *__restrict__ x, const double *__restrict__ y,
                              const double *__restrict__ z, const double *__restrict__ D,
                              const double *__restrict__ wq, int lx, int64_t npts,
                              double *__restrict__ g, double *__restrict__ bm1,
                              double *__restrict__ jac_out, double *__restrict__ rst) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const int n3 = lx * lx * lx;
  const int64_t e0 = p / n3 * n3;
  const int loc = (int)(p - e0);
  const int i = loc % lx, j = (loc / lx) % lx, k = loc / (lx * lx);
  double d[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // d[c][r] = d coord_c / d (r,s,t)
  const double *c[3] = {x, y, z};
  for (int l = 0; l < lx; ++l) {
    const double dr = D[i + lx * l], ds = D[j + lx * l], dt = D[k + lx * l];
    const int64_t pr = e0 + l + lx * (j + lx * k), ps = e0 + i + lx * (l + lx * k),
                  pt = e0 + i + lx * (j + lx * l);
    for (int a = 0; a < 3; ++a) {
      d[a][0] = fma(dr, c[a][pr], d[a][0]);
      d[a][1] = fma(ds, c[a][ps], d[a][1]);
      d[a][2] = fma(dt, c[a][pt], d[a][2]);
    }
  }
  const double xr = d[0][0], xs = d[0][1], xt = d[0][2], yr = d[1][0], ys = d[1][1], yt = d[1][2],
               zr = d[2][0], zs = d[2][1], zt = d[2][2];
  const double jac = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr);
  const double rx = ys * zt - yt * zs, ry = xt * zs - xs * zt, rz = xs * yt - xt * ys;
  const double sx = yt * zr - yr * zt, sy = xr * zt - xt * zr, sz = xt * yr - xr * yt;
  const double tx = yr * zs - ys * zr, ty = xs * zr - xr * zs, tz = xr * ys - xs * yr;
  const double w3 = wq[i] * wq[j] * wq[k];
  const double sc = w3 / jac;
  g[0 * npts + p] = (rx * rx + ry * ry + rz * rz) * sc;
  g[1 * npts + p] = (sx * sx + sy * sy + sz * sz) * sc;
  g[2 * npts + p] = (tx * tx + ty * ty + tz * tz) * sc;
  g[3 * npts + p] = (rx * sx + ry * sy + rz * sz) * sc;
  g[4 * npts + p] = (rx * tx + ry * ty + rz * tz) * sc;
  g[5 * npts + p] = (sx * tx + sy * ty + sz * tz) * sc;
  bm1[p] = jac * w3;
  jac_out[p] = jac;
  const double r9[9] = {rx, ry, rz, sx, sy, sz, tx, ty, tz};
  for (int a = 0; a < 9; ++a) rst[a * npts + p] = r9[a];
}

__global__ void geom2d_kernel(const double *__restrict__ x, const double *__restrict__ y,
                              const double *__restrict__ D, const double *__restrict__ wq, int lx,
                              int64_t npts, double *__restrict__ g, double *__restrict__ bm1,
                              double *__restrict__ jac_out, double *__restrict__ rst) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const int n2 = lx * lx;
  const int64_t e0 = p / n2 * n2;
  const int loc = (int)(p - e0);
  const int i = loc % lx, j = loc / lx;
  double xr = 0, xs = 0, yr = 0, ys = 0;
  for (int l = 0; l < lx; ++l) {
    const double dr = D[i + lx * l], ds = D[j + lx * l];
    const int64_t pr = e0 + l + lx * j, ps = e0 + i + lx * l;
    xr = fma(dr, x[pr], xr);
    yr = fma(dr, y[pr], yr);
    xs = fma(ds, x[ps], xs);
    ys = fma(ds, y[ps], ys);
  }
  const double jac = xr * ys - xs * yr;
  const double rx = ys, ry = -xs, sx = -yr, sy = xr;
  const double w2 = wq[i] * wq[j];
  const double sc = w2 / jac;
  g[0 * npts + p] = (rx * rx + ry * ry) * sc;
  g[1 * npts + p] = (sx * sx + sy * sy) * sc;
  g[2 * npts + p] = (rx * sx + ry * sy) * sc;
  bm1[p] = jac * w2;
  jac_out[p] = jac;
  rst[0 * npts + p] = rx;
  rst[1 * npts + p] = ry;
  rst[2 * npts + p] = sx;
  rst[3 * npts + p] = sy;
}

// ---- axhelm, 3-D -------------------------------------------------------------------------------
// Thread (i,j) of an element marches over k with its u-column and w-column in registers; the
// r- and s-derivatives go through one shared plane per element, the t-derivative stays in
// registers.  D (row-major) and D^T are staged in shared memory once per CTA.
//   EPI = 0 : w = h1 * D^T G D u + h2 * bm1 * u [+ C . grad u]         (raw, element-local)
//   EPI = 1 : as 0 on element-boundary points (to be summed by the gather-scatter kernel),
//             final value alpha*u + beta*bmask*w on element-interior points
template <int LX, bool CONV, int EPI>
__global__ void __launch_bounds__((256 / (LX * LX) > 0 ? 256 / (LX * LX) : 1) * LX * LX)
axhelm3d_kernel(const double *__restrict__ u, double *__restrict__ w, const double *__restrict__ g,
                const double *__restrict__ bm1, const double *__restrict__ Dg, int64_t nel,
                int64_t npts, double h1, double h2, const double *__restrict__ cv, double alpha,
                double beta, const double *__restrict__ bmask, int64_t fstride) {
  u += (int64_t)blockIdx.y * fstride;
  w += (int64_t)blockIdx.y * fstride;
  constexpr int EPC = (256 / (LX * LX) > 0 ? 256 / (LX * LX) : 1);
  constexpr int N2 = LX * LX, N3 = LX * LX * LX;
  __shared__ double sD[LX * LX], sDt[LX * LX];
  __shared__ double s_u[EPC][LX][LX], s_wr[EPC][LX][LX], s_ws[EPC][LX][LX];
  const int tid = threadIdx.x;
  for (int t = tid; t < N2; t += EPC * N2) {
    const int a = t / LX, b = t % LX;   // Dg[a + LX*b] = D_ab
    sD[a * LX + b] = Dg[a + LX * b];
    sDt[b * LX + a] = Dg[a + LX * b];
  }
  const int el = tid / N2, ij = tid % N2, i = ij % LX, j = ij / LX;
  int64_t e = (int64_t)blockIdx.x * EPC + el;
  const bool active = e < nel;
  if (!active) e = nel - 1;
  const int64_t base = e * N3 + ij;
  double uk[LX], wk[LX];
#pragma unroll
  for (int k = 0; k < LX; ++k) {
    uk[k] = u[base + k * N2];
    wk[k] = 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < LX; ++k) {
    const int64_t p = base + k * N2;
    // geometric factors of this plane: issued before the barrier so they overlap it
    const double g1 = ld_stream1(g + 0 * npts + p), g2 = ld_stream1(g + 1 * npts + p),
                 g3 = ld_stream1(g + 2 * npts + p), g4 = ld_stream1(g + 3 * npts + p),
                 g5 = ld_stream1(g + 4 * npts + p), g6 = ld_stream1(g + 5 * npts + p);
    double c1 = 0, c2 = 0, c3 = 0;
    if (CONV) {
      c1 = ld_stream1(cv + 0 * npts + p);
      c2 = ld_stream1(cv + 1 * npts + p);
      c3 = ld_stream1(cv + 2 * npts + p);
    }
    s_u[el][j][i] = uk[k];
    __syncthreads();
    double ur = 0, us = 0, ut = 0;
#pragma unroll
    for (int l = 0; l < LX; ++l) {
      ur = fma(sDt[l * LX + i], s_u[el][j][l], ur);
      us = fma(sDt[l * LX + j], s_u[el][l][i], us);
      ut = fma(sD[k * LX + l], uk[l], ut);
    }
    const double wr = h1 * (g1 * ur + g4 * us + g5 * ut);
    const double ws = h1 * (g2 * us + g4 * ur + g6 * ut);
    const double wt = h1 * (g3 * ut + g5 * ur + g6 * us);
    s_wr[el][j][i] = wr;
    s_ws[el][j][i] = ws;
    __syncthreads();
    double acc = CONV ? (c1 * ur + c2 * us + c3 * ut) : 0.0;
#pragma unroll
    for (int l = 0; l < LX; ++l) {
      acc = fma(sD[l * LX + i], s_wr[el][j][l], acc);
      acc = fma(sD[l * LX + j], s_ws[el][l][i], acc);
      wk[l] = fma(sD[k * LX + l], wt, wk[l]);
    }
    wk[k] += acc;
  }
  if (!active) return;
  const bool ij_bnd = (i == 0 || i == LX - 1 || j == 0 || j == LX - 1);
#pragma unroll
  for (int k = 0; k < LX; ++k) {
    const int64_t p = base + k * N2;
    double v = wk[k];
    if (h2 != 0.0) v = fma(h2 * ld_stream1(bm1 + p), uk[k], v);
    if (EPI == 1) {
      const bool bnd = ij_bnd || k == 0 || k == LX - 1;
      if (!bnd) v = alpha * uk[k] + beta * ld_stream1(bmask + p) * v;
    }
    w[p] = v;
  }
}

// ---- axhelm, 3-D, N = 7: one warp per element -------------------------------------------------
// ncu on the kernel above (64 threads per element, two CTA-wide barriers per plane) showed 25 %
// warps active, 3.75 shared wavefronts per point and 36 % DRAM.  Here a single warp owns an
// element, so planes are separated by __syncwarp only: lane (i = lane % 8, jp = lane / 8) holds
// the two u/w columns (i, jp, :) and (i, jp + 4, :) in registers, which lets the r-derivative
// share its D row and the s-derivative its u value between the two points (3 shared loads per
// point per l instead of 4); D(k,l) for the register-resident t-derivative comes from constant
// memory with compile-time indices; plane rows are padded to 10 doubles (conflict-free).
__constant__ double c_D8[64];    // c_D8[a * 8 + b] = D_ab for N = 7

template <bool CONV, int EPI>
__global__ void __launch_bounds__(128, 3)
axhelm3d_warp8_kernel(const double *__restrict__ u, double *__restrict__ w, const double *__restrict__ g,
                      const double *__restrict__ bm1, int64_t nel, int64_t npts, double h1, double h2,
                      const double *__restrict__ cv, double alpha, double beta,
                      const double *__restrict__ bmask, int64_t fstride) {
  u += (int64_t)blockIdx.y * fstride;
  w += (int64_t)blockIdx.y * fstride;
  constexpr int LX = 8, N2 = 64, N3 = 512, PS = 10, WPC = 4;   // PS: padded plane row stride
  __shared__ double sD[LX * LX], sDt[LX * LX];
  __shared__ double s_u[WPC][LX * PS], s_wr[WPC][LX * PS], s_ws[WPC][LX * PS];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (tid < N2) {
    const int a = tid / LX, b = tid % LX;
    sD[a * LX + b] = c_D8[a * LX + b];
    sDt[b * LX + a] = c_D8[a * LX + b];
  }
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * WPC + wp;
  if (e >= nel) return;
  const int i = lane & 7, jp = lane >> 3, j0 = jp, j1 = jp + 4;
  const int64_t base0 = e * N3 + j0 * LX + i, base1 = e * N3 + j1 * LX + i;
  double *su = s_u[wp], *swr = s_wr[wp], *sws = s_ws[wp];
  double uk0[LX], uk1[LX], wk0[LX], wk1[LX];
#pragma unroll
  for (int k = 0; k < LX; ++k) {
    uk0[k] = u[base0 + k * N2];
    uk1[k] = u[base1 + k * N2];
    wk0[k] = 0.0;
    wk1[k] = 0.0;
  }
  // geometric factors are prefetched one plane ahead so that a warp always has two planes of
  // loads in flight (ncu on the first warp-per-element version: 80 % of cycles with no eligible
  // warp, stalled on the loads issued at the top of each plane)
  double gn_a[6], gn_b[6], cn_a[3] = {0, 0, 0}, cn_b[3] = {0, 0, 0};
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    gn_a[c] = ld_stream1(g + c * npts + base0);
    gn_b[c] = ld_stream1(g + c * npts + base1);
  }
  if (CONV) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cn_a[c] = ld_stream1(cv + c * npts + base0);
      cn_b[c] = ld_stream1(cv + c * npts + base1);
    }
  }
#pragma unroll
  for (int k = 0; k < LX; ++k) {
    double ga[6], gb[6], ca[3], cb[3];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      ga[c] = gn_a[c];
      gb[c] = gn_b[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ca[c] = cn_a[c];
      cb[c] = cn_b[c];
    }
    if (k + 1 < LX) {
      const int64_t q0 = base0 + (k + 1) * N2, q1 = base1 + (k + 1) * N2;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        gn_a[c] = ld_stream1(g + c * npts + q0);
        gn_b[c] = ld_stream1(g + c * npts + q1);
      }
      if (CONV) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          cn_a[c] = ld_stream1(cv + c * npts + q0);
          cn_b[c] = ld_stream1(cv + c * npts + q1);
        }
      }
    }
    su[j0 * PS + i] = uk0[k];
    su[j1 * PS + i] = uk1[k];
    __syncwarp();
    double ur0 = 0, ur1 = 0, us0 = 0, us1 = 0, ut0 = 0, ut1 = 0;
#pragma unroll
    for (int l = 0; l < LX; ++l) {
      const double di = sDt[l * LX + i];          // D(i,l)
      ur0 = fma(di, su[j0 * PS + l], ur0);
      ur1 = fma(di, su[j1 * PS + l], ur1);
      const double b = su[l * PS + i];            // u(i,l,k)
      us0 = fma(sDt[l * LX + j0], b, us0);        // D(j0,l)
      us1 = fma(sDt[l * LX + j1], b, us1);
      const double dk = c_D8[k * LX + l];         // D(k,l), compile-time constant-bank address
      ut0 = fma(dk, uk0[l], ut0);
      ut1 = fma(dk, uk1[l], ut1);
    }
    const double wr0 = h1 * (ga[0] * ur0 + ga[3] * us0 + ga[4] * ut0);
    const double ws0 = h1 * (ga[1] * us0 + ga[3] * ur0 + ga[5] * ut0);
    const double wt0 = h1 * (ga[2] * ut0 + ga[4] * ur0 + ga[5] * us0);
    const double wr1 = h1 * (gb[0] * ur1 + gb[3] * us1 + gb[4] * ut1);
    const double ws1 = h1 * (gb[1] * us1 + gb[3] * ur1 + gb[5] * ut1);
    const double wt1 = h1 * (gb[2] * ut1 + gb[4] * ur1 + gb[5] * us1);
    swr[j0 * PS + i] = wr0;
    swr[j1 * PS + i] = wr1;
    sws[j0 * PS + i] = ws0;
    sws[j1 * PS + i] = ws1;
    __syncwarp();
    double a0 = CONV ? (ca[0] * ur0 + ca[1] * us0 + ca[2] * ut0) : 0.0;
    double a1 = CONV ? (cb[0] * ur1 + cb[1] * us1 + cb[2] * ut1) : 0.0;
#pragma unroll
    for (int l = 0; l < LX; ++l) {
      const double ci = sD[l * LX + i];           // D(l,i)
      a0 = fma(ci, swr[j0 * PS + l], a0);
      a1 = fma(ci, swr[j1 * PS + l], a1);
      const double ev = sws[l * PS + i];
      a0 = fma(sD[l * LX + j0], ev, a0);          // D(l,j0)
      a1 = fma(sD[l * LX + j1], ev, a1);
      const double dk = c_D8[k * LX + l];
      wk0[l] = fma(dk, wt0, wk0[l]);
      wk1[l] = fma(dk, wt1, wk1[l]);
    }
    wk0[k] += a0;
    wk1[k] += a1;
  }
  const bool b0 = (i == 0 || i == LX - 1 || j0 == 0), b1 = (i == 0 || i == LX - 1 || j1 == LX - 1);
#pragma unroll
  for (int k = 0; k < LX; ++k) {
    const int64_t p0 = base0 + k * N2, p1 = base1 + k * N2;
    double v0 = wk0[k], v1 = wk1[k];
    if (h2 != 0.0) {
      v0 = fma(h2 * ld_stream1(bm1 + p0), uk0[k], v0);
      v1 = fma(h2 * ld_stream1(bm1 + p1), uk1[k], v1);
    }
    if (EPI == 1) {
      const bool kb = (k == 0 || k == LX - 1);
      if (!(b0 || kb)) v0 = alpha * uk0[k] + beta * ld_stream1(bmask + p0) * v0;
      if (!(b1 || kb)) v1 = alpha * uk1[k] + beta * ld_stream1(bmask + p1) * v1;
    }
    w[p0] = v0;
    w[p1] = v1;
  }
}

// ---- axhelm, 3-D, N = 7: TMA ring, geometric factors shared by the velocity components ----------
// ncu on the warp-per-element kernel: 51 % DRAM, 80 % of cycles without an eligible warp, stalled
// on the loads of G1..G6 that every plane issues into registers.  Here a producer warp streams
// whole elements (u of NF fields, G1..G6, bm1, bmask: 4 KB bulk copies) into a ring of NSTAGE
// shared-memory stages on mbarriers, so NSTAGE-1 elements of loads are always in flight without
// occupying registers; the NF consumer warps of a stage each take one velocity component of that
// element and all read the SAME staged geometric factors: G, bm1 and bmask cross HBM once per
// element instead of once per component (8 (8 + 2 NF) bytes per point for NF fields).
template <int NF, bool CONV, int EPI, int NSTAGE>
__global__ void __launch_bounds__((NF * NSTAGE + 1) * 32, 1)
axhelm3d_ring8_kernel(const double *__restrict__ u, double *__restrict__ w, const double *__restrict__ g,
                      const double *__restrict__ bm1, int64_t nel, int64_t npts, double h1, double h2,
                      const double *__restrict__ cv, double alpha, double beta,
                      const double *__restrict__ bmask, int64_t fstride) {
  constexpr int LX = 8, N2 = 64, N3 = 512, PS = 10, NCW = NF * NSTAGE;
  constexpr int NARR = 6 + 2 + (CONV ? 3 : 0) + NF;       // arrays per stage: G1..G6, bm1, bmask, [C], u
  constexpr int STAGE = NARR * N3;                        // doubles per stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stage0 = reinterpret_cast<double *>(smem_raw);
  double *planes = stage0 + (size_t)NSTAGE * STAGE;       // [NCW][3][LX*PS]
  double *sD = planes + NCW * 3 * LX * PS;                // [64] D_ab, then [64] D_ba
  double *sDt = sD + 64;
  uint64_t *full = reinterpret_cast<uint64_t *>(sDt + 64);   // [NSTAGE]
  uint64_t *empty = full + NSTAGE;                           // [NSTAGE]
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (tid < N2) {
    const int a = tid / LX, b = tid % LX;
    sD[a * LX + b] = c_D8[a * LX + b];
    sDt[b * LX + a] = c_D8[a * LX + b];
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NF);
    }
  }
  __syncthreads();
  const int64_t nit = (nel - blockIdx.x + gridDim.x - 1) / gridDim.x;   // elements of this CTA

  if (wp == NCW) {
    // ===== producer warp: one lane per array =====
    for (int64_t it = 0; it < nit; ++it) {
      const int s = (int)(it % NSTAGE);
      const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
      mbar_wait(empty + s, ph ^ 1u);                       // fresh barrier: passes immediately
      const int64_t e = blockIdx.x + it * gridDim.x;
      double *dst = stage0 + (size_t)s * STAGE;
      if (lane == 0) mbar_expect_tx(full + s, (uint32_t)(STAGE * sizeof(double)));
      __syncwarp();
      if (lane < NARR) {
        const double *src;
        if (lane < 6) src = g + (int64_t)lane * npts + e * N3;
        else if (lane == 6) src = bm1 + e * N3;
        else if (lane == 7) src = bmask + e * N3;
        else if (CONV && lane < 11) src = cv + (int64_t)(lane - 8) * npts + e * N3;
        else src = u + (int64_t)(lane - (CONV ? 11 : 8)) * fstride + e * N3;
        tma_bulk_g2s(dst + lane * N3, src, N3 * sizeof(double), full + s);
      }
    }
    return;
  }

  // ===== consumer warps: stage = wp / NF, velocity component = wp % NF =====
  const int s = wp / NF, f = wp % NF;
  const int i = lane & 7, jp = lane >> 3, j0 = jp, j1 = jp + 4;
  const int q0 = j0 * LX + i, q1 = j1 * LX + i;             // in-plane offsets of the two points
  double *su = planes + (size_t)wp * 3 * LX * PS, *swr = su + LX * PS, *sws = swr + LX * PS;
  const double *sG = stage0 + (size_t)s * STAGE;
  const double *sB = sG + 6 * N3, *sM = sG + 7 * N3, *sC = sG + 8 * N3;
  const double *sU = sG + (size_t)(8 + (CONV ? 3 : 0) + f) * N3;
  double *wout = w + (int64_t)f * fstride;
  const bool b0 = (i == 0 || i == LX - 1 || j0 == 0), b1 = (i == 0 || i == LX - 1 || j1 == LX - 1);
  // row i and column i of D never change for this lane: keep them in registers (ncu: the shared
  // pipe is the limiter of this kernel, 87 % busy, and D reads are 40 % of its loads)
  constexpr bool CACHE_D = (NSTAGE <= 3);   // 10 warps leave 168 registers per thread, 13 warps only 128
  double dri[CACHE_D ? LX : 1], dci[CACHE_D ? LX : 1];
  if (CACHE_D) {
#pragma unroll
    for (int l = 0; l < LX; ++l) {
      dri[l % (CACHE_D ? LX : 1)] = sDt[l * LX + i];   // D(i,l)
      dci[l % (CACHE_D ? LX : 1)] = sD[l * LX + i];    // D(l,i)
    }
  }
  for (int64_t it = s; it < nit; it += NSTAGE) {
    const int64_t e = blockIdx.x + it * gridDim.x;
    mbar_wait(full + s, (uint32_t)((it / NSTAGE) & 1));
    double uk0[LX], uk1[LX], wk0[LX], wk1[LX];
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      uk0[k] = sU[k * N2 + q0];
      uk1[k] = sU[k * N2 + q1];
      wk0[k] = 0.0;
      wk1[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      su[j0 * PS + i] = uk0[k];
      su[j1 * PS + i] = uk1[k];
      __syncwarp();
      double ur0 = 0, ur1 = 0, us0 = 0, us1 = 0, ut0 = 0, ut1 = 0;
#pragma unroll
      for (int l = 0; l < LX; ++l) {
        const double di = CACHE_D ? dri[l % (CACHE_D ? LX : 1)] : sDt[l * LX + i];
        ur0 = fma(di, su[j0 * PS + l], ur0);
        ur1 = fma(di, su[j1 * PS + l], ur1);
        const double b = su[l * PS + i];
        us0 = fma(sDt[l * LX + j0], b, us0);
        us1 = fma(sDt[l * LX + j1], b, us1);
        const double dk = c_D8[k * LX + l];
        ut0 = fma(dk, uk0[l], ut0);
        ut1 = fma(dk, uk1[l], ut1);
      }
      const int p0 = k * N2 + q0, p1 = k * N2 + q1;
      double wt0, wt1;
      {
        const double g1 = sG[p0], g2 = sG[N3 + p0], g3 = sG[2 * N3 + p0], g4 = sG[3 * N3 + p0],
                     g5 = sG[4 * N3 + p0], g6 = sG[5 * N3 + p0];
        swr[j0 * PS + i] = h1 * (g1 * ur0 + g4 * us0 + g5 * ut0);
        sws[j0 * PS + i] = h1 * (g2 * us0 + g4 * ur0 + g6 * ut0);
        wt0 = h1 * (g3 * ut0 + g5 * ur0 + g6 * us0);
      }
      {
        const double g1 = sG[p1], g2 = sG[N3 + p1], g3 = sG[2 * N3 + p1], g4 = sG[3 * N3 + p1],
                     g5 = sG[4 * N3 + p1], g6 = sG[5 * N3 + p1];
        swr[j1 * PS + i] = h1 * (g1 * ur1 + g4 * us1 + g5 * ut1);
        sws[j1 * PS + i] = h1 * (g2 * us1 + g4 * ur1 + g6 * ut1);
        wt1 = h1 * (g3 * ut1 + g5 * ur1 + g6 * us1);
      }
      __syncwarp();
      double a0 = 0.0, a1 = 0.0;
      if (CONV) {
        a0 = sC[p0] * ur0 + sC[N3 + p0] * us0 + sC[2 * N3 + p0] * ut0;
        a1 = sC[p1] * ur1 + sC[N3 + p1] * us1 + sC[2 * N3 + p1] * ut1;
      }
#pragma unroll
      for (int l = 0; l < LX; ++l) {
        const double ci = CACHE_D ? dci[l % (CACHE_D ? LX : 1)] : sD[l * LX + i];
        a0 = fma(ci, swr[j0 * PS + l], a0);
        a1 = fma(ci, swr[j1 * PS + l], a1);
        const double ev = sws[l * PS + i];
        a0 = fma(sD[l * LX + j0], ev, a0);
        a1 = fma(sD[l * LX + j1], ev, a1);
        const double dk = c_D8[k * LX + l];
        wk0[l] = fma(dk, wt0, wk0[l]);
        wk1[l] = fma(dk, wt1, wk1[l]);
      }
      wk0[k] += a0;
      wk1[k] += a1;
    }
    double *we = wout + e * N3;
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      const int p0 = k * N2 + q0, p1 = k * N2 + q1;
      double v0 = wk0[k], v1 = wk1[k];
      if (h2 != 0.0) {
        v0 = fma(h2 * sB[p0], uk0[k], v0);
        v1 = fma(h2 * sB[p1], uk1[k], v1);
      }
      if (EPI == 1) {
        const bool kb = (k == 0 || k == LX - 1);
        if (!(b0 || kb)) v0 = alpha * uk0[k] + beta * sM[p0] * v0;
        if (!(b1 || kb)) v1 = alpha * uk1[k] + beta * sM[p1] * v1;
      }
      we[p0] = v0;
      we[p1] = v1;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);   // this warp is done with the stage
  }
}

// D (8x8, fp64) += A (8x4) * B (4x8) on the fp64 tensor cores (SASS: DMMA.884).  Fragments per lane
// (g = lane / 4, t = lane % 4): a = A[g][t], b = B[t][g], d = {D[g][2t], D[g][2t+1]}.
__device__ __forceinline__ void dmma884(double2 &d, double a, double b) {
  double d0 = d.x, d1 = d.y;
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
  d.x = d0;
  d.y = d1;
}

// ---- axhelm, 3-D, N = 7: TMA ring + fp64 tensor-core contraction -------------------------------
// ncu on the ring kernel above: DRAM 35 %, shared pipe 87 % -- the r/s contractions read D rows and
// plane values from shared memory for every FMA.  BASELINE.json's north-star allows fp64 tensor
// cores exactly in this case.  Each 8x8 plane contraction becomes two DMMA m8n8k4; the only shared
// traffic left is one 128-bit store and two 64-bit loads per B-operand plane (plus the staged
// G1..G6, read as 128-bit pairs).
//
// Ring: NBUF element buffers serve NG warp groups (NF warps each, one per velocity component).
// Element `it` of the CTA lives in buffer it % NBUF and is processed by group it % NG; with
// NBUF = NG + 1 one buffer is always loading while every group computes.
template <int NF, bool CONV>
struct Dmma8Cfg {
  static constexpr int NARR = 6 + 2 + (CONV ? 3 : 0) + NF;      // arrays per buffer: G1..G6, bm1, bmask, [C], u
  static constexpr int PLANE = 80;                              // doubles per conversion plane
  static constexpr size_t per_buf = sizeof(double) * NARR * 512 + 16;
  static constexpr size_t fixed_for(int ng) { return sizeof(double) * ((size_t)NF * ng * 2 * PLANE + 128) + 256; }
  static constexpr int fit_for(int ng) { return (int)((227 * 1024 - fixed_for(ng)) / per_buf); }
  // Warp groups: as many as pay off, but never more than buffers -- with NG > NBUF a group waits for a
  // buffer several uses ahead and the parity waits become ambiguous (tests/test_ring_protocol.py).
  static constexpr int NG0 = NF == 3 ? 3 : (NF == 2 ? 4 : 5);
  static constexpr int NG = fit_for(NG0) < NG0 ? fit_for(NG0) : NG0;
  static constexpr size_t fixed = fixed_for(NG);
  static constexpr int fit = fit_for(NG);
  static constexpr int NBUF = fit < NG + 1 ? fit : NG + 1;
  static constexpr size_t smem = fixed + (size_t)NBUF * per_buf;
  static_assert(NG >= 1 && NBUF >= NG, "every warp group needs a buffer of its own");
};

template <int NF, bool CONV, int EPI, int NBUF>
__global__ void __launch_bounds__((NF * Dmma8Cfg<NF, CONV>::NG + 1) * 32, 1)
axhelm3d_dmma8_kernel(const double *__restrict__ u, double *__restrict__ w, const double *__restrict__ g,
                      const double *__restrict__ bm1, int64_t nel, int64_t npts, double h1, double h2,
                      const double *__restrict__ cv, double alpha, double beta,
                      const double *__restrict__ bmask, int64_t fstride, int geo_evict_first,
                      double *__restrict__ dotp) {
  using Cfg = Dmma8Cfg<NF, CONV>;
  constexpr int LX = 8, N2 = 64, N3 = 512, NG = Cfg::NG, NCW = NF * NG, PLANE = Cfg::PLANE;
  constexpr int NARR = Cfg::NARR;
  constexpr int STAGE = NARR * N3;                        // doubles per buffer
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stage0 = reinterpret_cast<double *>(smem_raw);
  double *planes = stage0 + (size_t)NBUF * STAGE;         // [NCW][2][PLANE]
  double *sD = planes + NCW * 2 * PLANE;                  // [64] D
  uint64_t *full = reinterpret_cast<uint64_t *>(sD + 128);   // [NBUF]
  uint64_t *empty = full + NBUF;                             // [NBUF]
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (tid < N2) sD[tid] = c_D8[tid];
  if (tid == 0) {
    for (int s = 0; s < NBUF; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NF);
    }
  }
  __syncthreads();
  const int64_t nit = (nel - blockIdx.x + gridDim.x - 1) / gridDim.x;   // elements of this CTA

  if (wp == NCW) {
    // ===== producer warp: one lane per array =====
    // slab pipeline: G, bm1, bmask and C are streamed once -> evict_first, so that u and w of the slab stay in
    // L2 for the gather-scatter that follows on the second stream
    const bool hint = geo_evict_first && lane < (CONV ? 11 : 8);
    const uint64_t pol = l2_policy_evict_first();
    for (int64_t it = 0; it < nit; ++it) {
      const int s = (int)(it % NBUF);
      const uint32_t ph = (uint32_t)((it / NBUF) & 1);
      mbar_wait(empty + s, ph ^ 1u);                       // fresh barrier: passes immediately
      const int64_t e = blockIdx.x + it * gridDim.x;
      double *dst = stage0 + (size_t)s * STAGE;
      if (lane == 0) mbar_expect_tx(full + s, (uint32_t)(STAGE * sizeof(double)));
      __syncwarp();
      if (lane < NARR) {
        const double *src;
        if (lane < 6) src = g + (int64_t)lane * npts + e * N3;
        else if (lane == 6) src = bm1 + e * N3;
        else if (lane == 7) src = bmask + e * N3;
        else if (CONV && lane < 11) src = cv + (int64_t)(lane - 8) * npts + e * N3;
        else src = u + (int64_t)(lane - (CONV ? 11 : 8)) * fstride + e * N3;
        if (hint) tma_bulk_g2s_hint(dst + lane * N3, src, N3 * sizeof(double), full + s, pol);
        else tma_bulk_g2s(dst + lane * N3, src, N3 * sizeof(double), full + s);
      }
    }
    return;
  }

  // ===== consumer warps: group = wp / NF, velocity component = wp % NF =====
  // Plane matrices are handled TRANSPOSED, M~[j][i] = m(i,j), in the fragment layouts of
  // mma.m8n8k4.f64 (gq = lane / 4, t = lane % 4):
  //   C / accumulator : rows j = gq, columns i = 2t, 2t+1   -> the lane OWNS points (2t..2t+1, gq, k):
  //                     two adjacent doubles in memory, so stage reads and the final store are 128-bit
  // The two m8n8k4 of a plane product split the contraction index into EVEN (0,2,4,6) and ODD
  // (1,3,5,7) instead of low / high halves.  Lane (gq, t) then supplies A[gq][2t] to the first and
  // A[gq][2t+1] to the second -- exactly the two accumulator values it owns -- so a matrix that is
  // an A operand (U~ in UR~ = U~ D^T, WR~ in WR~ D) never leaves registers.  Only the B operands
  // (U~ in US~ = D U~, WS~ in D^T WS~) change layout: lane (gq, t) needs M~[2t][gq], M~[2t+1][gq].
  // They go through one padded plane each, address(row, col) = 20 (row / 2) + 8 (row % 2) + col:
  // the 128-bit row stores and the 64-bit column loads are both bank-conflict-free.
  //   r-derivative  UR~ = U~ D^T   (A = U~ registers,  B = D^T : lane constants D[gq][2t], D[gq][2t+1])
  //   s-derivative  US~ = D  U~    (A = D    : the same lane constants,  B = U~ via the plane)
  //   transposed:   W~ += WR~ D  +  D^T WS~   (constants D[2t][gq], D[2t+1][gq])
  const int grp = wp / NF, f = wp % NF;
  const int gq = lane >> 2, t = lane & 3;
  const int q = gq * LX + 2 * t;                              // in-plane offset of the lane's point pair
  double *pu = planes + (size_t)wp * 2 * PLANE, *pws = pu + PLANE;
  const int st_off = 20 * (gq >> 1) + 8 * (gq & 1) + 2 * t;   // C layout: row gq, columns 2t, 2t+1
  const int ld_off0 = 20 * t + gq, ld_off1 = ld_off0 + 8;     // B layout: rows 2t, 2t+1, column gq
  double *wout = w + (int64_t)f * fstride;
  const double dR0 = sD[gq * LX + 2 * t], dR1 = sD[gq * LX + 2 * t + 1];     // D[gq][2t], D[gq][2t+1]
  const double dT0 = sD[2 * t * LX + gq], dT1 = sD[(2 * t + 1) * LX + gq];   // D[2t][gq], D[2t+1][gq]
  // element-boundary flags of the two points (i = 2t, 2t+1 ; j = gq)
  const bool bj = (gq == 0 || gq == LX - 1);
  const bool bx = bj || (2 * t == 0), by = bj || (2 * t + 1 == LX - 1);
  // (w, u) over this warp's points: for a continuous u the sum over local points of w_raw u equals the assembled
  // inner product (QQ^T w, u)_mult -- the (w, p) of the conjugate-gradient iteration without a pass of its own
  double dacc = 0.0;
  for (int64_t it = grp; it < nit; it += NG) {
    const int64_t e = blockIdx.x + it * gridDim.x;
    const int s = (int)(it % NBUF);
    const double *sG = stage0 + (size_t)s * STAGE;
    const double *sB = sG + 6 * N3, *sM = sG + 7 * N3, *sC = sG + 8 * N3;
    const double *sU = sG + (size_t)(8 + (CONV ? 3 : 0) + f) * N3;
    // A parity wait is only unambiguous when the waiter is at most one phase ahead of the barrier.  With
    // NBUF > NG consecutive uses of a buffer belong to DIFFERENT warps: the warp of use u can get here
    // while use u-1 is still being loaded, full[s] is then two phases behind and its parity already
    // matches (seen as launch failures on large meshes with NF = 1).  Waiting first for the previous
    // user's release (empty[s], phase u-1) pins full[s] to phase u.
    const uint32_t use = (uint32_t)(it / NBUF);
    if (use > 0) mbar_wait(empty + s, (use - 1u) & 1u);
    mbar_wait(full + s, use & 1u);
    __syncwarp();   // the spin loops may leave the lanes diverged; mma.sync.aligned below needs the full warp
    // always 0, but opaque to the compiler: keeps the 64 D(k,l) constant loads inside the loop
    // (hoisted out of it they occupy 128 registers and spill)
    const int zoff = (int)(it >> 40);
    double2 uk[LX], wk[LX];
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      uk[k] = *reinterpret_cast<const double2 *>(sU + k * N2 + q);
      wk[k] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      *reinterpret_cast<double2 *>(pu + st_off) = uk[k];
      double2 ur = make_double2(0.0, 0.0), us = make_double2(0.0, 0.0);
      dmma884(ur, uk[k].x, dR0);
      dmma884(ur, uk[k].y, dR1);
      double2 ut = make_double2(0.0, 0.0);
#pragma unroll
      for (int l = 0; l < LX; ++l) {
        const double dk = c_D8[k * LX + l + zoff];
        ut.x = fma(dk, uk[l].x, ut.x);
        ut.y = fma(dk, uk[l].y, ut.y);
      }
      __syncwarp();
      const double b0 = pu[ld_off0], b1 = pu[ld_off1];                   // B layout of U~
      dmma884(us, dR0, b0);
      dmma884(us, dR1, b1);
      const int p = k * N2 + q;
      const double2 g1 = *reinterpret_cast<const double2 *>(sG + p), g2 = *reinterpret_cast<const double2 *>(sG + N3 + p),
                    g3 = *reinterpret_cast<const double2 *>(sG + 2 * N3 + p),
                    g4 = *reinterpret_cast<const double2 *>(sG + 3 * N3 + p),
                    g5 = *reinterpret_cast<const double2 *>(sG + 4 * N3 + p),
                    g6 = *reinterpret_cast<const double2 *>(sG + 5 * N3 + p);
      double2 wr, ws, wt;
      ws.x = h1 * (g2.x * us.x + g4.x * ur.x + g6.x * ut.x);
      ws.y = h1 * (g2.y * us.y + g4.y * ur.y + g6.y * ut.y);
      *reinterpret_cast<double2 *>(pws + st_off) = ws;
      wr.x = h1 * (g1.x * ur.x + g4.x * us.x + g5.x * ut.x);
      wr.y = h1 * (g1.y * ur.y + g4.y * us.y + g5.y * ut.y);
      wt.x = h1 * (g3.x * ut.x + g5.x * ur.x + g6.x * us.x);
      wt.y = h1 * (g3.y * ut.y + g5.y * ur.y + g6.y * us.y);
      double2 acc = make_double2(0.0, 0.0);
      if (CONV) {
        const double2 c1 = *reinterpret_cast<const double2 *>(sC + p), c2 = *reinterpret_cast<const double2 *>(sC + N3 + p),
                      c3 = *reinterpret_cast<const double2 *>(sC + 2 * N3 + p);
        acc.x = c1.x * ur.x + c2.x * us.x + c3.x * ut.x;
        acc.y = c1.y * ur.y + c2.y * us.y + c3.y * ut.y;
      }
      dmma884(acc, wr.x, dT0);
      dmma884(acc, wr.y, dT1);
#pragma unroll
      for (int l = 0; l < LX; ++l) {
        const double dk = c_D8[k * LX + l + zoff];
        wk[l].x = fma(dk, wt.x, wk[l].x);
        wk[l].y = fma(dk, wt.y, wk[l].y);
      }
      __syncwarp();
      const double bs0 = pws[ld_off0], bs1 = pws[ld_off1];               // B layout of WS~
      dmma884(acc, dT0, bs0);
      dmma884(acc, dT1, bs1);
      wk[k].x += acc.x;
      wk[k].y += acc.y;
    }
    double *we = wout + e * N3;
#pragma unroll
    for (int k = 0; k < LX; ++k) {
      const int p = k * N2 + q;
      double2 v = wk[k];
      if (h2 != 0.0) {
        const double2 bm = *reinterpret_cast<const double2 *>(sB + p);
        v.x = fma(h2 * bm.x, uk[k].x, v.x);
        v.y = fma(h2 * bm.y, uk[k].y, v.y);
      }
      if (EPI == 1) {
        const bool kb = (k == 0 || k == LX - 1);
        const double2 mk = *reinterpret_cast<const double2 *>(sM + p);
        if (!(bx || kb)) v.x = alpha * uk[k].x + beta * mk.x * v.x;
        if (!(by || kb)) v.y = alpha * uk[k].y + beta * mk.y * v.y;
      }
      if (EPI == 0) dacc = fma(v.x, uk[k].x, fma(v.y, uk[k].y, dacc));
      *reinterpret_cast<double2 *>(we + p) = v;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);   // this warp is done with the buffer
  }
  if (EPI == 0 && dotp) {                    // partial[(cta * NG + group)][field]
    dacc = warp_reduce_sum(dacc);
    if (lane == 0) dotp[((size_t)blockIdx.x * NG + grp) * NF + f] = dacc;
  }
}

template <int NF, bool CONV, int NSTAGE>
constexpr size_t ring8_smem() {
  return sizeof(double) * ((size_t)NSTAGE * (6 + 2 + (CONV ? 3 : 0) + NF) * 512 + NF * NSTAGE * 3 * 80 + 128) +
         sizeof(uint64_t) * 2 * NSTAGE + 128;
}

template <int NF, bool CONV, int EPI, int NSTAGE>
int launch_ring8_s(nsb_sem_t S, const double *u, double *w, int64_t fstride, double h1, double h2,
                   const double *cv, double alpha, double beta, const double *bmask) {
  constexpr size_t smem = ring8_smem<NF, CONV, NSTAGE>();
  static_assert(smem <= 227 * 1024, "axhelm ring does not fit in shared memory");
  auto kfn = axhelm3d_ring8_kernel<NF, CONV, EPI, NSTAGE>;
  NSB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = S->ax_nel < S->ctx->num_sms ? S->ax_nel : S->ctx->num_sms;
  kfn<<<(unsigned)grid, (NF * NSTAGE + 1) * 32, smem, S->ctx->stream>>>(u, w, S->ax_g, S->ax_bm1, S->ax_nel, S->npts, h1,
                                                                       h2, cv, alpha, beta, bmask, fstride);
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

template <int NF, bool CONV, int EPI, int NBUF>
int launch_dmma8_s(nsb_sem_t S, const double *u, double *w, int64_t fstride, double h1, double h2,
                   const double *cv, double alpha, double beta, const double *bmask) {
  using Cfg = Dmma8Cfg<NF, CONV>;
  constexpr size_t smem = Cfg::fixed + (size_t)NBUF * (sizeof(double) * Cfg::NARR * 512 + 16);
  static_assert(smem <= 227 * 1024, "axhelm dmma ring does not fit in shared memory");
  auto kfn = axhelm3d_dmma8_kernel<NF, CONV, EPI, NBUF>;
  NSB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = S->ax_nel < S->ctx->num_sms ? S->ax_nel : S->ctx->num_sms;
  kfn<<<(unsigned)grid, (NF * Cfg::NG + 1) * 32, smem, S->ctx->stream>>>(u, w, S->ax_g, S->ax_bm1, S->ax_nel, S->npts, h1,
                                                                       h2, cv, alpha, beta, bmask, fstride,
                                                                       S->nslab > 1 && S->ax_nel < S->nel ? 1 : 0,
                                                                       EPI == 0 ? S->ax_dotp : nullptr);
  if (EPI == 0 && S->ax_dotp) S->ax_dot_rows = (int)grid * Cfg::NG;
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

template <int NF, bool CONV, int EPI>
int launch_dmma8(nsb_sem_t S, const double *u, double *w, int64_t fstride, double h1, double h2,
                 const double *cv, double alpha, double beta, const double *bmask) {
  using Cfg = Dmma8Cfg<NF, CONV>;
  // NSB_AX_STAGES=<NG> pins the ring to one buffer per warp group (the pre-decoupling behaviour)
  if constexpr (Cfg::NBUF > Cfg::NG) {
    if (S->ctx->ax_stages == Cfg::NG)
      return launch_dmma8_s<NF, CONV, EPI, Cfg::NG>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmask);
  }
  return launch_dmma8_s<NF, CONV, EPI, Cfg::NBUF>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmask);
}

template <int NF, bool CONV, int EPI>
int launch_ring8(nsb_sem_t S, const double *u, double *w, int64_t fstride, double h1, double h2,
                 const double *cv, double alpha, double beta, const double *bmask) {
  if (S->ctx->ax_dmma) return launch_dmma8<NF, CONV, EPI>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmask);
  if (CONV || S->ctx->ax_stages != 4)
    return launch_ring8_s<NF, CONV, EPI, 3>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmask);
  return launch_ring8_s<NF, CONV, EPI, CONV ? 3 : 4>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmask);
}

// ---- axhelm, 2-D (one thread per point; parity configurations only) --------------------------
template <bool CONV, int EPI>
__global__ void axhelm2d_kernel(const double *__restrict__ u, double *__restrict__ w,
                                const double *__restrict__ g, const double *__restrict__ bm1,
                                const double *__restrict__ Dg, int lx, int64_t nel, int64_t npts,
                                double h1, double h2, const double *__restrict__ cv, double alpha,
                                double beta, const double *__restrict__ bmask, int64_t fstride) {
  u += (int64_t)blockIdx.y * fstride;
  w += (int64_t)blockIdx.y * fstride;
  extern __shared__ double sm[];
  const int n2 = lx * lx;
  double *sD = sm, *s_u = sm + n2, *s_wr = sm + 2 * n2, *s_ws = sm + 3 * n2;
  const int64_t e = blockIdx.x;
  const int t = threadIdx.x, i = t % lx, j = t / lx;
  const int64_t p = e * n2 + t;
  const double uv = u[p];
  sD[t] = Dg[i + lx * j];  // sD[j*lx + i] = D_ij  -> sD[b*lx + a] = D_ab
  s_u[t] = uv;
  __syncthreads();
  double ur = 0, us = 0;
  for (int l = 0; l < lx; ++l) {
    ur = fma(sD[l * lx + i], s_u[j * lx + l], ur);  // D_il u(l,j)
    us = fma(sD[l * lx + j], s_u[l * lx + i], us);  // D_jl u(i,l)
  }
  const double g1 = g[p], g2 = g[npts + p], g4 = g[2 * npts + p];
  s_wr[t] = h1 * (g1 * ur + g4 * us);
  s_ws[t] = h1 * (g2 * us + g4 * ur);
  __syncthreads();
  double v = CONV ? (cv[p] * ur + cv[npts + p] * us) : 0.0;
  for (int l = 0; l < lx; ++l) {
    v = fma(sD[i * lx + l], s_wr[j * lx + l], v);  // D_li wr(l,j)
    v = fma(sD[j * lx + l], s_ws[l * lx + i], v);  // D_lj ws(i,l)
  }
  if (h2 != 0.0) v = fma(h2 * bm1[p], uv, v);
  if (EPI == 1) {
    const bool bnd = (i == 0 || i == lx - 1 || j == 0 || j == lx - 1);
    if (!bnd) v = alpha * uv + beta * bmask[p] * v;
  }
  w[p] = v;
}

// ---- gather-scatter ---------------------------------------------------------------------------
// One thread per unique node that owns at least one element-boundary point.
//   EPI = 0 : every copy <- sum of copies                                  (dssum)
//   EPI = 1 : every copy p <- alpha * uin[p] + beta * bmask[p] * sum       (fused operator tail)
//   EPI = 2 : only write the node sums to node_sum (multi-rank first phase)
//   EPI = 3 : scatter node_sum with the EPI=0 rule,  EPI = 4 : scatter with the EPI=1 rule
// All nf fields of a node are handled by the same thread (in groups of up to FG): the offsets, the
// point indices and bmask are fetched once per node instead of once per field -- with one grid row
// per field they were re-read from DRAM for every velocity component (ncu: 1.43 GB read where the
// fields themselves account for 0.94 GB).
template <int EPI>
__global__ void __launch_bounds__(256)
gs_kernel(double *__restrict__ v, const int64_t *__restrict__ off, const int32_t *__restrict__ idx,
          int64_t n0, int64_t n1, const double *__restrict__ uin, double alpha, double beta,
          const double *__restrict__ bmask, double *__restrict__ node_sum, int64_t fstride,
          int64_t ns_stride, int nf, const double *__restrict__ bnode) {
  constexpr int FG = 3;
  const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n1) return;
  const int64_t a = off[n], b = off[n + 1];
  // bmask is the same on every copy of a node: one coalesced read per node when the caller's bmask
  // is the mesh's own (bnode), else one scattered read per copy
  const double bn = ((EPI == 1 || EPI == 4) && bnode) ? beta * bnode[n] : 0.0;
  for (int f0 = 0; f0 < nf; f0 += FG) {
    const int nfg = nf - f0 < FG ? nf - f0 : FG;
    double *vf = v + (int64_t)f0 * fstride;
    double s[FG];
#pragma unroll
    for (int f = 0; f < FG; ++f) s[f] = 0.0;
    if (EPI == 3 || EPI == 4) {
#pragma unroll
      for (int f = 0; f < FG; ++f)
        if (f < nfg) s[f] = node_sum[(int64_t)(f0 + f) * ns_stride + (n - n0)];
    } else {
      for (int64_t q = a; q < b; ++q) {
        const int32_t p = idx[q];
#pragma unroll
        for (int f = 0; f < FG; ++f)
          if (f < nfg) s[f] += vf[(int64_t)f * fstride + p];
      }
    }
    if (EPI == 2) {
#pragma unroll
      for (int f = 0; f < FG; ++f)
        if (f < nfg) node_sum[(int64_t)(f0 + f) * ns_stride + (n - n0)] = s[f];
      continue;
    }
    for (int64_t q = a; q < b; ++q) {
      const int32_t p = idx[q];
      if (EPI == 0 || EPI == 3) {
#pragma unroll
        for (int f = 0; f < FG; ++f)
          if (f < nfg) vf[(int64_t)f * fstride + p] = s[f];
      } else {
        const double bb = bnode ? bn : beta * bmask[p];
        double u[FG];
#pragma unroll
        for (int f = 0; f < FG; ++f)
          if (f < nfg) u[f] = uin[(int64_t)(f0 + f) * fstride + p];
#pragma unroll
        for (int f = 0; f < FG; ++f)
          if (f < nfg) vf[(int64_t)f * fstride + p] = alpha * u[f] + bb * s[f];
      }
    }
  }
}

__global__ void bnode_kernel(const double *__restrict__ bmask, const int64_t *__restrict__ off,
                             const int32_t *__restrict__ idx, int64_t n, double *__restrict__ bnode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bnode[i] = bmask[idx[off[i]]];
}

// interface buffers hold nf fields back to back: buf[f * n + t]
__global__ void pack_kernel(const double *__restrict__ node_sum, const int32_t *__restrict__ nodes,
                            int64_t n, double *__restrict__ buf, int64_t ns_stride) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (t < n) buf[(int64_t)f * n + t] = node_sum[(int64_t)f * ns_stride + nodes[t]];
}

__global__ void unpack_add_kernel(double *__restrict__ node_sum, const int32_t *__restrict__ nodes,
                                  int64_t n, const double *__restrict__ buf, int64_t ns_stride) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (t < n) node_sum[(int64_t)f * ns_stride + nodes[t]] += buf[(int64_t)f * n + t];
}

__global__ void col2_kernel(double *__restrict__ v, const double *__restrict__ c, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) v[t] *= c[t];
}

__global__ void recip_kernel(double *__restrict__ v, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) v[t] = 1.0 / v[t];
}

__global__ void mul3_kernel(double *__restrict__ out, const double *__restrict__ a,
                            const double *__restrict__ b, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) out[t] = a[t] * b[t];
}

// contravariant convecting velocity times quadrature weight: C_r = w3 (cx rx + cy ry + cz rz) ...
__global__ void conv_coeff_kernel(const double *__restrict__ rst, const double *__restrict__ cx,
                                  const double *__restrict__ cy, const double *__restrict__ cz,
                                  const double *__restrict__ wq, int dim, int lx, int64_t npts,
                                  double *__restrict__ cv) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  const int loc = (int)(p % nloc);
  const int i = loc % lx, j = (loc / lx) % lx, k = loc / (lx * lx);
  if (dim == 3) {
    const double w3 = wq[i] * wq[j] * wq[k];
    for (int a = 0; a < 3; ++a)
      cv[a * npts + p] = w3 * (cx[p] * rst[(3 * a + 0) * npts + p] + cy[p] * rst[(3 * a + 1) * npts + p] +
                               cz[p] * rst[(3 * a + 2) * npts + p]);
  } else {
    const double w2 = wq[i] * wq[j];
    for (int a = 0; a < 2; ++a)
      cv[a * npts + p] = w2 * (cx[p] * rst[(2 * a + 0) * npts + p] + cy[p] * rst[(2 * a + 1) * npts + p]);
  }
}

template <int LX, bool CONV, int EPI>
void launch_ax3d_t(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                   const double *cv, double alpha, double beta, const double *bmask) {
  constexpr int EPC = (256 / (LX * LX) > 0 ? 256 / (LX * LX) : 1);
  const int64_t grid = (S->ax_nel + EPC - 1) / EPC;
  axhelm3d_kernel<LX, CONV, EPI><<<dim3((unsigned)grid, nf), EPC * LX * LX, 0, S->ctx->stream>>>(
      u, w, S->ax_g, S->ax_bm1, S->D_d, S->ax_nel, S->npts, h1, h2, cv, alpha, beta, bmask, fstride);
}

template <bool CONV, int EPI>
int launch_ax3d(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                const double *cv, double alpha, double beta, const double *bmask) {
  if (S->lx == 8 && !S->ctx->ax_generic && S->ctx->ax_ring && nf <= 3) {
    // bmask / bm1 are staged unconditionally: hand the kernel valid arrays even when unused
    const double *bmk = bmask ? bmask : S->ax_bmask;
    if (nf == 3) return launch_ring8<3, CONV, EPI>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmk);
    if (nf == 2) return launch_ring8<2, CONV, EPI>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmk);
    return launch_ring8<1, CONV, EPI>(S, u, w, fstride, h1, h2, cv, alpha, beta, bmk);
  }
  if (S->lx == 8 && !S->ctx->ax_generic) {
    const int64_t grid = (S->ax_nel + 3) / 4;
    axhelm3d_warp8_kernel<CONV, EPI><<<dim3((unsigned)grid, nf), 128, 0, S->ctx->stream>>>(
        u, w, S->ax_g, S->ax_bm1, S->ax_nel, S->npts, h1, h2, cv, alpha, beta, bmask, fstride);
    S->ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
  switch (S->lx) {
#define CASE(L) case L: launch_ax3d_t<L, CONV, EPI>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
    default:
      set_error("axhelm: lx1=%d not instantiated (2..12)", S->lx);
      return NSB_EINVAL;
  }
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

template <bool CONV, int EPI>
int launch_ax2d(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                const double *cv, double alpha, double beta, const double *bmask) {
  const int n2 = S->lx * S->lx;
  axhelm2d_kernel<CONV, EPI><<<dim3((unsigned)S->ax_nel, nf), n2, sizeof(double) * 4 * n2, S->ctx->stream>>>(
      u, w, S->ax_g, S->ax_bm1, S->D_d, S->lx, S->ax_nel, S->npts, h1, h2, cv, alpha, beta, bmask, fstride);
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// axhelm on nf fields that sit fstride doubles apart (velocity components of one column)
int launch_axhelm(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                  const double *cv, int epi, double alpha, double beta, const double *bmask, int64_t e0 = 0,
                  int64_t ne = -1) {
  cudaSetDevice(S->ctx->device);
  // element range [e0, e0 + ne) of this launch (the slab pipeline of the fused operator; default: all)
  if (ne < 0) ne = S->nel - e0;
  const int64_t nloc = S->npts / S->nel, sh = e0 * nloc;
  u += sh;
  w += sh;
  if (cv) cv += sh;
  if (bmask) bmask += sh;
  S->ax_g = S->g_d + sh;
  S->ax_bm1 = S->bm1_d + sh;
  S->ax_bmask = S->bmask_d + sh;
  S->ax_nel = ne;
  // algorithmic bytes per point: u, w, G1..G6 (G1,G2,G4 in 2-D), bm1 if h2 != 0, C if convecting,
  // bmask on element-interior points of the fused epilogue
  const double fint = std::pow((double)(S->lx - 2) / S->lx, S->dim);
  // Algorithmic bytes of what is launched: the N = 7 ring kernels stage G, bm1, bmask and C once per
  // element for all nf components, 8 (geo + 2 nf) bytes per point; the other kernels read them per
  // component, 8 (geo + 2) nf.  (SURVEY.md section 8d's per-component figure, 64 n + 8 n for the
  // h2 B term, would credit the ring kernels with about twice the bytes they move.)
  const bool shared_geo = S->dim == 3 && S->lx == 8 && !S->ctx->ax_generic && S->ctx->ax_ring && nf <= 3;
  const double geo = 8.0 * (S->ng + (h2 != 0.0 ? 1 : 0) + (cv ? S->dim : 0) + (epi ? fint : 0.0));
  const double per_pt = shared_geo ? geo + 16.0 * nf : (geo + 16.0) * nf;
  ProfScope ps(S->ctx, PC_AXHELM, per_pt * (double)(ne * nloc));
  if (S->dim == 3) {
    if (cv) return epi ? launch_ax3d<true, 1>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask)
                       : launch_ax3d<true, 0>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask);
    return epi ? launch_ax3d<false, 1>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask)
               : launch_ax3d<false, 0>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask);
  }
  if (cv) return epi ? launch_ax2d<true, 1>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask)
                     : launch_ax2d<true, 0>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask);
  return epi ? launch_ax2d<false, 1>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask)
             : launch_ax2d<false, 0>(S, u, w, nf, fstride, h1, h2, cv, alpha, beta, bmask);
}

inline unsigned blocks_for(int64_t n, int nt = 256) { return (unsigned)((n + nt - 1) / nt); }

template <int EPI>
void gs_launch(nsb_sem_t S, cudaStream_t st, double *v, int64_t n0, int64_t n1, int nf, int64_t fstride,
               const double *uin, double alpha, double beta, const double *bmask, double *ns, int64_t ns_stride) {
  if (n1 <= n0) return;
  gs_kernel<EPI><<<blocks_for(n1 - n0), 256, 0, st>>>(v, S->gs_off_d, S->gs_idx_d, n0, n1, uin, alpha, beta, bmask,
                                                    ns, fstride, ns_stride, nf,
                                                    (bmask && bmask == S->bmask_d) ? S->bnode_d : nullptr);
  S->ctx->launches++;
}

// Gather-scatter on nf fields incl. the inter-rank exchange; epi 0: plain dssum, 1: fused operator
// tail.  Nodes [0, n_local) are private to this rank and go through one fused kernel; the
// interface nodes [n_local, nshared) take the sum -> pack -> NCCL send/recv -> add -> scatter path
// on a second stream, overlapped with the private part.
int launch_gs(nsb_sem_t S, double *v, int nf, int64_t fstride, int epi, const double *uin, double alpha,
              double beta, const double *bmask) {
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  if (S->nshared == 0) return NSB_OK;
  if (ctx->nranks > 1 && !S->exchange_ready) {
    set_error("dssum: nsb_sem_setup_exchange has not been called on this multi-rank context");
    return NSB_EINVAL;
  }
  // algorithmic bytes: every listed point read + written (16), its index (4), node offsets (8 per
  // node), plus uin and bmask per point in the fused tail
  // (the index, the offsets and bmask are shared by the nf fields)
  ProfScope ps(ctx, PC_GS, (double)S->gs_nnz * (4.0 + (epi ? 8.0 : 0.0)) + 8.0 * (double)S->nshared +
                               nf * (double)S->gs_nnz * (16.0 + (epi ? 8.0 : 0.0)));
  const int64_t nloc = S->n_local, nifc = S->nshared - S->n_local;
  if (ctx->nranks == 1 || nifc == 0 || S->peers.empty()) {
    if (epi == 0) gs_launch<0>(S, ctx->stream, v, 0, S->nshared, nf, fstride, nullptr, 0, 0, nullptr, nullptr, 0);
    else gs_launch<1>(S, ctx->stream, v, 0, S->nshared, nf, fstride, uin, alpha, beta, bmask, nullptr, 0);
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
  NSB_REQUIRE(nf <= S->ns_fields, "dssum: %d fields in one call, interface buffers hold %d", nf, S->ns_fields);
  cudaStream_t s2 = ctx->copy_stream;
  double *ns = S->node_sum_d;
  NSB_CUDA(cudaEventRecord(S->ev_a, ctx->stream));
  NSB_CUDA(cudaStreamWaitEvent(s2, S->ev_a, 0));
  if (S->p2p_halo && ctx->halo_fused) {
    // peer-memory transport: two kernels for the whole interface path
    NSB_CHECK(halo_exchange_fused(S, v, nf, fstride, epi, uin, alpha, beta, bmask, s2));
    NSB_CUDA(cudaEventRecord(S->ev_b, s2));
    if (epi == 0) gs_launch<0>(S, ctx->stream, v, 0, nloc, nf, fstride, nullptr, 0, 0, nullptr, nullptr, 0);
    else gs_launch<1>(S, ctx->stream, v, 0, nloc, nf, fstride, uin, alpha, beta, bmask, nullptr, 0);
    NSB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_b, 0));
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
  gs_launch<2>(S, s2, v, nloc, S->nshared, nf, fstride, nullptr, 0, 0, nullptr, ns, nifc);
  if (S->p2p_halo) {
    NSB_CHECK(halo_exchange_p2p(S, nf, s2));   // stores into the peers' mailboxes, no NCCL launch
  } else {
    for (auto &P : S->peers) {
      pack_kernel<<<dim3(blocks_for(P.n), nf), 256, 0, s2>>>(ns, P.idx_d, P.n, P.send_d, nifc);
      ctx->launches++;
    }
    NSB_CUDA(cudaGetLastError());
    NSB_CHECK(sendrecv_d(ctx, S->peers, nf, s2));
    for (auto &P : S->peers) {
      unpack_add_kernel<<<dim3(blocks_for(P.n), nf), 256, 0, s2>>>(ns, P.idx_d, P.n, P.recv_d, nifc);
      ctx->launches++;
    }
  }
  if (epi == 0) gs_launch<3>(S, s2, v, nloc, S->nshared, nf, fstride, nullptr, 0, 0, nullptr, ns, nifc);
  else gs_launch<4>(S, s2, v, nloc, S->nshared, nf, fstride, uin, alpha, beta, bmask, ns, nifc);
  NSB_CUDA(cudaEventRecord(S->ev_b, s2));
  // private nodes, concurrently on the main stream
  if (epi == 0) gs_launch<0>(S, ctx->stream, v, 0, nloc, nf, fstride, nullptr, 0, 0, nullptr, nullptr, 0);
  else gs_launch<1>(S, ctx->stream, v, 0, nloc, nf, fstride, uin, alpha, beta, bmask, nullptr, 0);
  NSB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_b, 0));
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

// Fused operator w = alpha u + beta bmask QQ^T (h1 A + h2 B [+ C.grad]) u in element slabs: the axhelm of slab
// s + 1 (main stream, DRAM-bound) runs while the gather-scatter of the nodes completed by slab s works out of
// L2 on a second stream -- w and u of a slab are read and re-written by the gather-scatter before they leave
// the cache, instead of crossing HBM a second time (ncu, round 1: 0.91 GB read + 0.36 GB written by gs_kernel
// on top of the 1.86 GB of axhelm).  Interface nodes (multi-rank) take the exchange path after the last slab.
int launch_ax_gs_pipelined(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                           const double *cv, double alpha, double beta) {
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  const double *bmask = S->bmask_d;
  const bool multi = ctx->nranks > 1 && (S->nshared - S->n_local) > 0 && !S->peers.empty();
  if (ctx->nranks > 1 && !S->exchange_ready) {
    set_error("dssum: nsb_sem_setup_exchange has not been called on this multi-rank context");
    return NSB_EINVAL;
  }
  for (int q = 0; q < S->nslab; ++q) {
    const int64_t e0 = S->slab_e0[q], ne = S->slab_e0[q + 1] - e0;
    NSB_CHECK(launch_axhelm(S, u, w, nf, fstride, h1, h2, cv, 1, alpha, beta, bmask, e0, ne));
    const int64_t n0 = q ? S->slab_node_end[q - 1] : 0, n1 = S->slab_node_end[q];
    if (n1 > n0) {
      NSB_CUDA(cudaEventRecord(S->slab_ev[q], ctx->stream));
      NSB_CUDA(cudaStreamWaitEvent(ctx->gs_stream, S->slab_ev[q], 0));
      gs_launch<1>(S, ctx->gs_stream, w, n0, n1, nf, fstride, u, alpha, beta, bmask, nullptr, 0);
    }
  }
  NSB_CUDA(cudaEventRecord(S->ev_c, ctx->gs_stream));
  if (multi) {
    NSB_REQUIRE(nf <= S->ns_fields, "dssum: %d fields in one call, interface buffers hold %d", nf, S->ns_fields);
    const int64_t nloc = S->n_local, nifc = S->nshared - S->n_local;
    cudaStream_t s2 = ctx->copy_stream;
    double *ns = S->node_sum_d;
    NSB_CUDA(cudaEventRecord(S->ev_a, ctx->stream));
    NSB_CUDA(cudaStreamWaitEvent(s2, S->ev_a, 0));
    gs_launch<2>(S, s2, w, nloc, S->nshared, nf, fstride, nullptr, 0, 0, nullptr, ns, nifc);
    if (S->p2p_halo) {
      NSB_CHECK(halo_exchange_p2p(S, nf, s2));
    } else {
      for (auto &P : S->peers) {
        pack_kernel<<<dim3(blocks_for(P.n), nf), 256, 0, s2>>>(ns, P.idx_d, P.n, P.send_d, nifc);
        ctx->launches++;
      }
      NSB_CUDA(cudaGetLastError());
      NSB_CHECK(sendrecv_d(ctx, S->peers, nf, s2));
      for (auto &P : S->peers) {
        unpack_add_kernel<<<dim3(blocks_for(P.n), nf), 256, 0, s2>>>(ns, P.idx_d, P.n, P.recv_d, nifc);
        ctx->launches++;
      }
    }
    gs_launch<4>(S, s2, w, nloc, S->nshared, nf, fstride, u, alpha, beta, bmask, ns, nifc);
    NSB_CUDA(cudaEventRecord(S->ev_b, s2));
    NSB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_b, 0));
  }
  NSB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_c, 0));
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

int field_ptr(nsb_sem_t S, nsb_basis_t B, int col, int field, double **out, const char *who) {
  NSB_REQUIRE(S && B, "%s: NULL argument", who);
  NSB_REQUIRE(col >= 0 && col < B->ncols, "%s: column %d out of range", who, col);
  nsb_layout_t L = B->lay;
  NSB_REQUIRE(field >= 0 && field < L->nfields, "%s: field %d out of range", who, field);
  NSB_REQUIRE(L->len[field] == S->npts, "%s: field %d has %lld points, mesh has %lld", who, field,
              (long long)L->len[field], (long long)S->npts);
  NSB_REQUIRE(L->ctx == S->ctx, "%s: basis and mesh live on different contexts", who);
  *out = B->col(col) + L->off[field];
  return NSB_OK;
}

}  // namespace

namespace nsb {
int launch_axhelm_ext(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                      const double *cv, int epi, double alpha, double beta, const double *bmask) {
  return launch_axhelm(S, u, w, nf, fstride, h1, h2, cv, epi, alpha, beta, bmask);
}
int launch_gs_ext(nsb_sem_t S, double *v, int nf, int64_t fstride, int epi, const double *uin, double alpha, double beta,
                  const double *bmask) {
  return launch_gs(S, v, nf, fstride, epi, uin, alpha, beta, bmask);
}
// Host plan of the gather-scatter: unique nodes owning at least one element-boundary point, CSR,
// ordered by first local index so neighbouring threads touch neighbouring memory.
int gs_plan(int dim, int lx, int64_t nel, const int64_t *glo_num, const double *mask, std::vector<int64_t> &off,
            std::vector<int32_t> &idx, std::vector<int64_t> &gid, std::vector<double> *vmult, int64_t slab_elems,
            std::vector<int32_t> *node_slab) {
  int64_t nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  const int64_t npts = nel * nloc;
  std::vector<int32_t> bpts;
  bpts.reserve((size_t)(npts * 0.6));
  for (int64_t p = 0; p < npts; ++p) {
    const int loc = (int)(p % nloc);
    const int i = loc % lx, j = (loc / lx) % lx, k = dim == 3 ? loc / (lx * lx) : 1;
    const bool bnd = i == 0 || i == lx - 1 || j == 0 || j == lx - 1 ||
                     (dim == 3 && (k == 0 || k == lx - 1));
    if (bnd) bpts.push_back((int32_t)p);
    if (glo_num[p] < 0) {
      set_error("gather-scatter plan: negative global id at point %lld", (long long)p);
      return NSB_EINVAL;
    }
  }
  std::stable_sort(bpts.begin(), bpts.end(),
                   [&](int32_t a, int32_t b) { return glo_num[a] < glo_num[b]; });
  std::vector<int64_t> grp_start;
  for (size_t q = 0; q < bpts.size(); ++q)
    if (q == 0 || glo_num[bpts[q]] != glo_num[bpts[q - 1]]) grp_start.push_back((int64_t)q);
  const int64_t ngrp = (int64_t)grp_start.size();
  grp_start.push_back((int64_t)bpts.size());
  std::vector<int64_t> order(ngrp);
  std::iota(order.begin(), order.end(), 0);
  // slab pipeline: nodes grouped by the element slab of their LAST copy (all copies of such a node have been
  // written once that slab's axhelm is done), by first local index inside a group
  auto slab_of = [&](int64_t g) -> int64_t {
    return slab_elems > 0 ? (bpts[grp_start[g + 1] - 1] / nloc) / slab_elems : 0;
  };
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    const int64_t sa = slab_of(a), sb = slab_of(b);
    return sa != sb ? sa < sb : bpts[grp_start[a]] < bpts[grp_start[b]];
  });
  if (node_slab) node_slab->resize(ngrp);
  off.assign(ngrp + 1, 0);
  idx.resize(bpts.size());
  gid.resize(ngrp);
  if (vmult) vmult->assign(npts, 1.0);
  int64_t w0 = 0;
  for (int64_t n = 0; n < ngrp; ++n) {
    const int64_t gI = order[n];
    off[n] = w0;
    const int64_t cnt = grp_start[gI + 1] - grp_start[gI];
    for (int64_t q = grp_start[gI]; q < grp_start[gI + 1]; ++q) {
      idx[w0++] = bpts[q];
      if (vmult) (*vmult)[bpts[q]] = 1.0 / (double)cnt;
      if (mask && mask[bpts[q]] != mask[bpts[grp_start[gI]]]) {
        set_error("nsb_sem_create: mask differs between copies of global node %lld",
                  (long long)glo_num[bpts[q]]);
        return NSB_EINVAL;
      }
    }
    gid[n] = glo_num[bpts[grp_start[gI]]];
    if (node_slab) (*node_slab)[n] = (int32_t)slab_of(gI);
  }
  off[ngrp] = w0;
  return NSB_OK;
}

// slab_node_end[s] = number of PRIVATE nodes (the first n_local ones, slab-sorted) whose last copy lies in a
// slab <= s
void compute_slab_ends(nsb_sem_t S) {
  S->slab_node_end.assign(S->nslab, 0);
  if (S->nslab <= 1) {
    if (S->nslab == 1) S->slab_node_end[0] = S->n_local;
    return;
  }
  for (int64_t n = 0; n < S->n_local; ++n) S->slab_node_end[S->node_slab[n]]++;
  for (int s = 1; s < S->nslab; ++s) S->slab_node_end[s] += S->slab_node_end[s - 1];
}
}  // namespace nsb

// Host-only (no CUDA): sizes first (off/idx/gid NULL), then the lists.
extern "C" int nsb_host_gs_plan(int dim, int N, int64_t nel, const int64_t *glo_num, int64_t *nnodes,
                                int64_t *nnz, int64_t *off, int32_t *idx, int64_t *gid) {
  NSB_REQUIRE(glo_num && nnodes && nnz && (dim == 2 || dim == 3) && N >= 1 && nel >= 1,
              "nsb_host_gs_plan: bad argument");
  std::vector<int64_t> o, g;
  std::vector<int32_t> ix;
  NSB_CHECK(nsb::gs_plan(dim, N + 1, nel, glo_num, nullptr, o, ix, g, nullptr));
  *nnodes = (int64_t)g.size();
  *nnz = (int64_t)ix.size();
  if (off) memcpy(off, o.data(), sizeof(int64_t) * o.size());
  if (idx) memcpy(idx, ix.data(), sizeof(int32_t) * ix.size());
  if (gid) memcpy(gid, g.data(), sizeof(int64_t) * g.size());
  return NSB_OK;
}

// ------------------------------------------------------------------------------------------------
// mesh object
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_sem_create(nsb_context_t ctx, int dim, int N, int64_t nel, const double *x,
                              const double *y, const double *z, const double *mask,
                              const int64_t *glo_num, nsb_sem_t *out) {
  NSB_REQUIRE(ctx && out && x && y && glo_num, "nsb_sem_create: NULL argument");
  NSB_REQUIRE(dim == 2 || dim == 3, "nsb_sem_create: dim=%d", dim);
  NSB_REQUIRE(dim == 2 || z != nullptr, "nsb_sem_create: z is NULL in 3-D");
  NSB_REQUIRE(N >= 1 && N <= 11, "nsb_sem_create: N=%d (1..11 supported)", N);
  NSB_REQUIRE(nel >= 1, "nsb_sem_create: nel=%lld", (long long)nel);
  const int lx = N + 1;
  int64_t nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  const int64_t npts = nel * nloc;
  NSB_REQUIRE(npts < ((int64_t)1 << 31), "nsb_sem_create: %lld points exceed the int32 index range",
              (long long)npts);
  cudaSetDevice(ctx->device);
  nsb_sem_t S = new nsb_sem_s();
  S->ctx = ctx;
  S->dim = dim;
  S->N = N;
  S->lx = lx;
  S->nel = nel;
  S->npts = npts;
  S->ng = dim == 3 ? 6 : 3;
  S->D_h.resize(lx * lx);
  S->z_h.resize(lx);
  S->w_h.resize(lx);
  nsb_gll(N, S->z_h.data(), S->w_h.data(), S->D_h.data());
  if (lx == 8) {
    double Drm[64];  // row-major D_ab
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b) Drm[a * 8 + b] = S->D_h[a + 8 * b];
    NSB_CUDA(cudaMemcpyToSymbol(c_D8, Drm, sizeof(Drm)));
  }
  const size_t nb = sizeof(double) * npts;
  double *wq_d = nullptr, *xyz_d = nullptr;
  NSB_CUDA(cudaMalloc(&S->D_d, sizeof(double) * lx * lx));
  NSB_CUDA(cudaMalloc(&wq_d, sizeof(double) * lx));
  NSB_CUDA(cudaMalloc(&xyz_d, nb * dim));
  NSB_CUDA(cudaMalloc(&S->g_d, nb * S->ng));
  NSB_CUDA(cudaMalloc(&S->bm1_d, nb));
  NSB_CUDA(cudaMalloc(&S->jac_d, nb));
  NSB_CUDA(cudaMalloc(&S->binv_d, nb));
  NSB_CUDA(cudaMalloc(&S->vmult_d, nb));
  NSB_CUDA(cudaMalloc(&S->mask_d, nb));
  NSB_CUDA(cudaMalloc(&S->bmask_d, nb));
  NSB_CUDA(cudaMalloc(&S->rst_d, nb * dim * dim));
  cudaStream_t s = ctx->stream;
  NSB_CUDA(cudaMemcpyAsync(S->D_d, S->D_h.data(), sizeof(double) * lx * lx, cudaMemcpyHostToDevice, s));
  NSB_CUDA(cudaMemcpyAsync(wq_d, S->w_h.data(), sizeof(double) * lx, cudaMemcpyHostToDevice, s));
  NSB_CUDA(cudaMemcpyAsync(xyz_d, x, nb, cudaMemcpyHostToDevice, s));
  NSB_CUDA(cudaMemcpyAsync(xyz_d + npts, y, nb, cudaMemcpyHostToDevice, s));
  if (dim == 3) NSB_CUDA(cudaMemcpyAsync(xyz_d + 2 * npts, z, nb, cudaMemcpyHostToDevice, s));
  if (dim == 3)
    geom3d_kernel<<<blocks_for(npts), 256, 0, s>>>(xyz_d, xyz_d + npts, xyz_d + 2 * npts, S->D_d, wq_d, lx,
                                                   npts, S->g_d, S->bm1_d, S->jac_d, S->rst_d);
  else
    geom2d_kernel<<<blocks_for(npts), 256, 0, s>>>(xyz_d, xyz_d + npts, S->D_d, wq_d, lx, npts, S->g_d,
                                                   S->bm1_d, S->jac_d, S->rst_d);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());

  // gather-scatter lists (host plan shared with the CPU-testable entry point nsb_host_gs_plan)
  S->glo_h.assign(glo_num, glo_num + npts);
  std::vector<int64_t> off;
  std::vector<int32_t> idx;
  std::vector<double> vm;
  // slab pipeline of the fused operator: slabs sized so that u and w of three fields of one slab (and its share
  // of the streamed geometric factors) stay in the 126 MB L2 until the slab's gather-scatter has run
  S->nslab = 1;
  int64_t slab_elems = 0;
  if (ctx->ax_slab_mb > 0.0) {
    const double per_elem = 6.0 * 8.0 * (double)nloc;   // u + w, three fields
    int64_t want = (int64_t)(ctx->ax_slab_mb * 1e6 / per_elem);
    want = std::max<int64_t>(want, 4 * (int64_t)ctx->num_sms);
    const int64_t ns = (nel + want - 1) / want;
    if (ns >= 2) {
      S->nslab = (int)std::min<int64_t>(ns, 256);
      slab_elems = (nel + S->nslab - 1) / S->nslab;
      S->nslab = (int)((nel + slab_elems - 1) / slab_elems);
    }
  }
  S->slab_e0.assign(S->nslab + 1, 0);
  for (int q = 0; q <= S->nslab; ++q) S->slab_e0[q] = S->nslab == 1 ? (q ? nel : 0) : std::min<int64_t>(nel, q * slab_elems);
  NSB_CHECK(nsb::gs_plan(dim, lx, nel, glo_num, mask, off, idx, S->node_gid, &vm, slab_elems, &S->node_slab));
  const int64_t ngrp = (int64_t)S->node_gid.size();
  const int64_t w0 = (int64_t)idx.size();
  S->nshared = ngrp;
  S->n_local = ngrp;
  nsb::compute_slab_ends(S);
  for (int q = 0; q < S->nslab; ++q) {
    cudaEvent_t e;
    NSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    S->slab_ev.push_back(e);
  }
  NSB_CUDA(cudaEventCreateWithFlags(&S->ev_c, cudaEventDisableTiming));
  S->gs_nnz = w0;
  S->gs_off_h = off;
  S->gs_idx_h = idx;
  NSB_CUDA(cudaEventCreateWithFlags(&S->ev_a, cudaEventDisableTiming));
  NSB_CUDA(cudaEventCreateWithFlags(&S->ev_b, cudaEventDisableTiming));
  NSB_CUDA(cudaMalloc(&S->gs_off_d, sizeof(int64_t) * (ngrp + 1)));
  NSB_CUDA(cudaMalloc(&S->gs_idx_d, sizeof(int32_t) * (w0 > 0 ? w0 : 1)));

  NSB_CUDA(cudaMemcpyAsync(S->gs_off_d, off.data(), sizeof(int64_t) * (ngrp + 1), cudaMemcpyHostToDevice, s));
  NSB_CUDA(cudaMemcpyAsync(S->gs_idx_d, idx.data(), sizeof(int32_t) * w0, cudaMemcpyHostToDevice, s));
  NSB_CUDA(cudaMemcpyAsync(S->vmult_d, vm.data(), nb, cudaMemcpyHostToDevice, s));
  if (mask) {
    NSB_CUDA(cudaMemcpyAsync(S->mask_d, mask, nb, cudaMemcpyHostToDevice, s));
  } else {
    std::vector<double> ones(npts, 1.0);
    NSB_CUDA(cudaMemcpyAsync(S->mask_d, ones.data(), nb, cudaMemcpyHostToDevice, s));
    NSB_CUDA(cudaStreamSynchronize(s));
  }
  NSB_CUDA(cudaStreamSynchronize(s));
  cudaFree(wq_d);
  cudaFree(xyz_d);
  *out = S;
  if (ctx->nranks == 1) return nsb_sem_setup_exchange(S);
  return NSB_OK;
}

// binvm1 = 1/dssum(bm1) and bmask = binvm1*mask need the (possibly inter-rank) dssum
static int finish_assembled(nsb_sem_t S) {
  nsb_context_t ctx = S->ctx;
  cudaStream_t s = ctx->stream;
  const size_t nb = sizeof(double) * S->npts;
  NSB_CUDA(cudaMemcpyAsync(S->binv_d, S->bm1_d, nb, cudaMemcpyDeviceToDevice, s));
  NSB_CHECK(launch_gs(S, S->binv_d, 1, 0, 0, nullptr, 0, 0, nullptr));
  recip_kernel<<<ctx->num_sms * 8, 256, 0, s>>>(S->binv_d, S->npts);
  mul3_kernel<<<ctx->num_sms * 8, 256, 0, s>>>(S->bmask_d, S->binv_d, S->mask_d, S->npts);
  ctx->launches += 2;
  if (S->nshared > 0) {
    if (!S->bnode_d) NSB_CUDA(cudaMalloc(&S->bnode_d, sizeof(double) * S->nshared));
    bnode_kernel<<<blocks_for(S->nshared), 256, 0, s>>>(S->bmask_d, S->gs_off_d, S->gs_idx_d, S->nshared, S->bnode_d);
    ctx->launches++;
  }
  NSB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1) {
    // vmult = 1/global multiplicity
    std::vector<double> ones(S->npts, 1.0);
    NSB_CUDA(cudaMemcpyAsync(S->vmult_d, ones.data(), nb, cudaMemcpyHostToDevice, s));
    NSB_CHECK(launch_gs(S, S->vmult_d, 1, 0, 0, nullptr, 0, 0, nullptr));
    recip_kernel<<<ctx->num_sms * 8, 256, 0, s>>>(S->vmult_d, S->npts);
    ctx->launches++;
    NSB_CUDA(cudaStreamSynchronize(s));
  }
  NSB_CUDA(cudaStreamSynchronize(s));
  return NSB_OK;
}

extern "C" int nsb_sem_setup_exchange(nsb_sem_t S) {
  NSB_REQUIRE(S, "nsb_sem_setup_exchange: NULL");
  if (S->ctx->nranks > 1) NSB_CHECK(nsb::exchange_setup(S));
  S->exchange_ready = true;
  return finish_assembled(S);
}

extern "C" int nsb_sem_destroy(nsb_sem_t S) {
  if (!S) return NSB_OK;
  cudaSetDevice(S->ctx->device);
  cudaStreamSynchronize(S->ctx->stream);
  cudaStreamSynchronize(S->ctx->copy_stream);
  cudaStreamSynchronize(S->ctx->gs_stream);
  clear_step_graphs(S->ctx);
  if (S->halo_flag_off >= 0 && --S->ctx->halo_users <= 0) {   // last mesh gone: the halo area is free again
    S->ctx->halo_users = 0;
    S->ctx->halo_used = 0;
  }
  if (S->hx_seq_d) cudaFree(S->hx_seq_d);
  if (S->hx_ticket_d) cudaFree(S->hx_ticket_d);
  if (S->ifc_poff_d) cudaFree(S->ifc_poff_d);
  if (S->ifc_pent_d) cudaFree(S->ifc_pent_d);
  for (auto &P : S->peers) {
    cudaFree(P.idx_d);
    cudaFree(P.send_d);
    cudaFree(P.recv_d);
  }
  double *ptrs[] = {S->g_d, S->bm1_d, S->jac_d, S->binv_d, S->vmult_d, S->mask_d, S->bmask_d,
                    S->rst_d, S->D_d, S->node_sum_d};
  for (double *p : ptrs)
    if (p) cudaFree(p);
  if (S->gs_off_d) cudaFree(S->gs_off_d);
  if (S->gs_idx_d) cudaFree(S->gs_idx_d);
  if (S->bnode_d) cudaFree(S->bnode_d);
  if (S->diagA_d) cudaFree(S->diagA_d);
  for (double *q : {S->J_d, S->Dg_d, S->rxf_d, S->cfine_d[0], S->cfine_d[1]})
    if (q) cudaFree(q);
  if (S->pcg_d) cudaFree(S->pcg_d);
  ns_free(S);
  if (S->c0_scratch_d) cudaFree(S->c0_scratch_d);
  if (S->c0_l2u_d) cudaFree(S->c0_l2u_d);
  if (S->ev_a) cudaEventDestroy(S->ev_a);
  if (S->ev_b) cudaEventDestroy(S->ev_b);
  if (S->ev_c) cudaEventDestroy(S->ev_c);
  for (auto e : S->slab_ev) cudaEventDestroy(e);
  delete S;
  return NSB_OK;
}

extern "C" int64_t nsb_sem_npts(nsb_sem_t S) { return S ? S->npts : -1; }

extern "C" int nsb_sem_get(nsb_sem_t S, int which, double *out) {
  NSB_REQUIRE(S && out, "nsb_sem_get: NULL argument");
  const double *src = nullptr;
  switch (which) {
    case 0: src = S->bm1_d; break;
    case 1: src = S->jac_d; break;
    case 2: src = S->binv_d; break;
    case 3: src = S->vmult_d; break;
    case 4: src = S->mask_d; break;
    default:
      if (which >= 10 && which < 16) {
        int gi = which - 10;
        if (S->dim == 2) {
          NSB_REQUIRE(gi == 0 || gi == 1 || gi == 3, "nsb_sem_get: 2-D has G1, G2, G4 only");
          gi = gi == 3 ? 2 : gi;
        }
        src = S->g_d + (size_t)gi * S->npts;
      }
  }
  NSB_REQUIRE(src, "nsb_sem_get: unknown selector %d", which);
  cudaSetDevice(S->ctx->device);
  NSB_CUDA(cudaMemcpyAsync(out, src, sizeof(double) * S->npts, cudaMemcpyDeviceToHost, S->ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(S->ctx->stream));
  return NSB_OK;
}

extern "C" int nsb_sem_axhelm(nsb_sem_t S, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout,
                              int field, double h1, double h2) {
  double *u, *w;
  NSB_CHECK(field_ptr(S, bin, cin, field, &u, "nsb_sem_axhelm"));
  NSB_CHECK(field_ptr(S, bout, cout, field, &w, "nsb_sem_axhelm"));
  NSB_REQUIRE(u != w, "nsb_sem_axhelm: in-place application is not supported");
  return launch_axhelm(S, u, w, 1, 0, h1, h2, nullptr, 0, 0, 0, nullptr);
}

extern "C" int nsb_sem_dssum(nsb_sem_t S, nsb_basis_t B, int col, int field) {
  double *v;
  NSB_CHECK(field_ptr(S, B, col, field, &v, "nsb_sem_dssum"));
  return launch_gs(S, v, 1, 0, 0, nullptr, 0, 0, nullptr);
}

extern "C" int nsb_sem_col2(nsb_sem_t S, nsb_basis_t B, int col, int field, int which) {
  double *v;
  NSB_CHECK(field_ptr(S, B, col, field, &v, "nsb_sem_col2"));
  const double *c = which == 0 ? S->bm1_d : which == 2 ? S->binv_d : which == 3 ? S->vmult_d
                    : which == 4 ? S->mask_d : nullptr;
  NSB_REQUIRE(c, "nsb_sem_col2: unknown selector %d", which);
  cudaSetDevice(S->ctx->device);
  col2_kernel<<<S->ctx->num_sms * 8, 256, 0, S->ctx->stream>>>(v, c, S->npts);
  S->ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

extern "C" int nsb_sem_ax(nsb_sem_t S, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout,
                          int field, double h1, double h2) {
  NSB_CHECK(nsb_sem_axhelm(S, bin, cin, bout, cout, field, h1, h2));
  NSB_CHECK(nsb_sem_dssum(S, bout, cout, field));
  return nsb_sem_col2(S, bout, cout, field, 4);
}

// ------------------------------------------------------------------------------------------------
// Jacobi-preconditioned conjugate gradients for the Helmholtz problem (h1 A + h2 B) x = rhs:
// [UPSTREAM-RECALL] Nek5000 hmholtz.f `cggo` + `setprec`, the solver nek_advance runs for every
// velocity component of every time step.  It is the loop
// { axhelm -> dssum -> mask -> glsc3 } built from the kernels above; inner products carry the
// inverse multiplicity (`mult` = vmult) so that duplicated nodes count once.
// ------------------------------------------------------------------------------------------------
namespace {

// diagonal of h1 A + h2 B per local point (no cross terms, like setprec away from deformed boundaries)
__global__ void helm_diag_kernel(const double *__restrict__ g, const double *__restrict__ bm1,
                                 const double *__restrict__ D, int dim, int lx, int64_t npts, double h1, double h2,
                                 double *__restrict__ diag) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int nloc = 1;
  for (int a = 0; a < dim; ++a) nloc *= lx;
  const int64_t e0 = p / nloc * nloc;
  const int loc = (int)(p - e0);
  const int i = loc % lx, j = (loc / lx) % lx, k = dim == 3 ? loc / (lx * lx) : 0;
  double s = 0.0;
  for (int q = 0; q < lx; ++q) {
    const double di = D[q + lx * i], dj = D[q + lx * j];           // D(q,i), D(q,j)
    s += di * di * g[0 * npts + e0 + q + lx * (j + lx * k)];
    s += dj * dj * g[1 * npts + e0 + i + lx * (q + lx * k)];
    if (dim == 3) {
      const double dk = D[q + lx * k];
      s += dk * dk * g[2 * npts + e0 + i + lx * (j + lx * q)];
    }
  }
  diag[p] = h1 * s + h2 * bm1[p];
}

// out = a x + b y
__global__ void axpby_fields_kernel(double *__restrict__ out, const double *__restrict__ x, const double *__restrict__ y,
                                    double a, double b, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) out[t] = a * x[t] + b * y[t];
}

// Scalars of the batched solver (up to kPcgFields systems at once).  They LIVE ON THE DEVICE: alpha, beta and
// the stopping decisions are computed by one-thread-per-system kernels between the sweeps, so an iteration is
// a fixed sequence of launches without a host round trip; the host polls the state every few iterations.
constexpr int kPcgFields = 3;
struct PcgState {
  double rtz1[kPcgFields], rtz2[kPcgFields], r0[kPcgFields], rn[kPcgFields], alpha[kPcgFields], beta[kPcgFields];
  int done[kPcgFields], itf[kPcgFields];
  int it, all_done;
};

// r = mask f ; partial[f] += r (d r) mult     (z = d r is never stored)
__global__ void __launch_bounds__(256) pcg_init_kernel(const double *__restrict__ f, int64_t fs_f, double *__restrict__ r,
                                                       const double *__restrict__ d, const double *__restrict__ mask,
                                                       const double *__restrict__ mult, int64_t n, int nf,
                                                       double *__restrict__ partial) {
  const int fi = blockIdx.y;
  const double *ff = f + (int64_t)fi * fs_f;
  double *rr = r + (int64_t)fi * n;
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const double rv = ff[t] * mask[t];        // the right-hand side lives in the masked space
    const double zv = d[t] * rv;
    rr[t] = rv;
    acc = fma(rv * zv, mult[t], acc);
  }
  acc = block_reduce_sum<256>(acc);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.x * nf + fi] = acc;
}

// partial[f] = sum w p mult   (kernels that cannot fold the product into their epilogue)
__global__ void __launch_bounds__(256) pcg_dot_kernel(const double *__restrict__ w, const double *__restrict__ p,
                                                      const double *__restrict__ mult, int64_t n, int nf,
                                                      double *__restrict__ partial) {
  const int fi = blockIdx.y;
  const double *ww = w + (int64_t)fi * n, *pp = p + (int64_t)fi * n;
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    acc = fma(ww[t] * pp[t], mult[t], acc);
  acc = block_reduce_sum<256>(acc);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.x * nf + fi] = acc;
}

// x += alpha p ; r -= alpha w ; partial[f] = r (d r) mult -- all systems of a point in one thread, so d and mult
// are read once per point; per system the arithmetic (and its order) does not depend on how many run side by side
template <int NF>
__global__ void __launch_bounds__(256) pcg_update_kernel(double *__restrict__ x, int64_t fs_x, double *__restrict__ r,
                                                         const double *__restrict__ p, const double *__restrict__ w,
                                                         const double *__restrict__ d, const double *__restrict__ mult,
                                                         const PcgState *__restrict__ st, int64_t n,
                                                         double *__restrict__ partial) {
  double al[NF], acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    al[f] = st->alpha[f];
    acc[f] = 0.0;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const double dv = ld_stream1(d + t), mv = ld_stream1(mult + t);
    double xv[NF], rv[NF], pv[NF], wv[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      xv[f] = x[(int64_t)f * fs_x + t];
      rv[f] = r[(int64_t)f * n + t];
      pv[f] = ld_stream1(p + (int64_t)f * n + t);
      wv[f] = ld_stream1(w + (int64_t)f * n + t);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      x[(int64_t)f * fs_x + t] = fma(al[f], pv[f], xv[f]);
      const double rn = fma(-al[f], wv[f], rv[f]);
      r[(int64_t)f * n + t] = rn;
      const double zv = dv * rn;
      acc[f] = fma(rn * zv, mv, acc[f]);
    }
  }
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const double s = block_reduce_sum<256>(acc[f]);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * NF + f] = s;
  }
}

// p = d r + beta p
template <int NF>
__global__ void __launch_bounds__(256) pcg_p_kernel(double *__restrict__ p, const double *__restrict__ r,
                                                    const double *__restrict__ d, const PcgState *__restrict__ st,
                                                    int64_t n) {
  double be[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) be[f] = st->beta[f];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const double dv = ld_stream1(d + t);
    double pv[NF], rv[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      pv[f] = p[(int64_t)f * n + t];
      rv[f] = ld_stream1(r + (int64_t)f * n + t);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) p[(int64_t)f * n + t] = fma(be[f], pv[f], dv * rv[f]);
  }
}

// ---- the scalar recurrences of cggo, one thread per system ----
__global__ void pcg_begin_kernel(PcgState *st, const double *__restrict__ red, int nf, int maxit) {
  const int g = threadIdx.x;
  if (g < kPcgFields) {
    st->rtz1[g] = g < nf ? red[g] : 0.0;
    st->rtz2[g] = 1.0;
    st->r0[g] = -1.0;
    st->rn[g] = 0.0;
    st->alpha[g] = 0.0;
    st->beta[g] = 0.0;
    st->done[g] = g < nf ? 0 : 1;
    st->itf[g] = maxit;
  }
  if (g == 0) {
    st->it = 1;
    st->all_done = 0;
  }
}

// after rho = (w, p): alpha = rtz1 / rho; a non-positive rho ends the system (round-off level / singular)
__global__ void pcg_alpha_kernel(PcgState *st, const double *__restrict__ red, int nf) {
  const int g = threadIdx.x;
  if (g >= kPcgFields) return;
  double al = 0.0;
  if (g < nf && !st->done[g]) {
    const double rho = red[g];
    if (!(rho > 0.0)) {
      st->done[g] = 1;
      st->itf[g] = st->it;
    } else {
      al = st->rtz1[g] / rho;
    }
  }
  st->alpha[g] = al;
}

// after rtz_new = (r, d r): shift, stopping test on sqrt((r, D r)) relative to the first one, beta of the next
// iteration (0 for a finished system: it is frozen, its x no longer changes)
__global__ void pcg_conv_kernel(PcgState *st, const double *__restrict__ red, int nf, double tol) {
  __shared__ int s_done[kPcgFields];
  const int g = threadIdx.x;
  if (g < kPcgFields) {
    if (g < nf && !st->done[g]) {
      st->rtz2[g] = st->rtz1[g];
      st->rtz1[g] = red[g];
      st->rn[g] = sqrt(fabs(st->rtz1[g]));
      if (st->r0[g] < 0.0) st->r0[g] = sqrt(fabs(st->rtz2[g]));
      if (st->rn[g] <= tol * st->r0[g]) {
        st->done[g] = 1;
        st->itf[g] = st->it;
      }
    }
    st->beta[g] = (g < nf && !st->done[g]) ? st->rtz1[g] / st->rtz2[g] : 0.0;
    s_done[g] = st->done[g];
  }
  __syncthreads();
  if (g == 0) {
    int all = 1;
    for (int q = 0; q < kPcgFields; ++q) all = all && s_done[q];
    st->all_done = all;
    st->it += 1;
  }
}

// partial[rows][nf] -> red[nf] on the device (summed over ranks)
int reduce_scalars_d(nsb_context_t ctx, int rows, int nf, double *red_d) {
  reduce_partials_kernel<<<1, 32 * kPcgFields, 0, ctx->stream>>>(ctx->partial_d, rows, nf, nf, red_d, 0, nullptr);
  ctx->launches++;
  NSB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1) NSB_CHECK(allreduce_sum_d(ctx, red_d, nf));
  return NSB_OK;
}

}  // namespace

// nf independent systems (h1 A + h2 B) x_f = rhs_f solved side by side: the matrix-vector product reads
// the geometric factors once for all of them (one axhelm + one gather-scatter launch per iteration),
// every system keeps its own alpha / beta / convergence test and is frozen (alpha = beta = 0) once converged,
// so each follows the iteration sequence it would follow alone.
// Nek masks w after the dssum; that pass is skipped here: d carries the mask, so z, p and x stay zero on
// the masked nodes whatever r collects there, and every inner product has an exact zero factor there.
// Per iteration: p = d r + beta p | w = H p with (w, p) accumulated in the axhelm epilogue | dssum |
// x += alpha p, r -= alpha w, (r, d r) -- four sweeps, z never stored, no host synchronisation; the host looks
// at the device-side state every kPcgPoll iterations (a converged system idles with alpha = beta = 0 meanwhile).
extern "C" int nsb_sem_hmholtz_vec(nsb_sem_t S, nsb_basis_t brhs, int crhs, nsb_basis_t bx, int cx, int field0,
                                   int nf, double h1, double h2, double tol, int maxit, int *iters, double *res) {
  NSB_REQUIRE(S && brhs && bx, "nsb_sem_hmholtz: NULL argument");
  NSB_REQUIRE(nf >= 1 && nf <= kPcgFields, "nsb_sem_hmholtz: %d systems at once (1..%d)", nf, kPcgFields);
  double *f, *x;
  for (int g = field0; g < field0 + nf; ++g) {
    NSB_CHECK(field_ptr(S, brhs, crhs, g, &f, "nsb_sem_hmholtz"));
    NSB_CHECK(field_ptr(S, bx, cx, g, &x, "nsb_sem_hmholtz"));
  }
  NSB_CHECK(field_ptr(S, brhs, crhs, field0, &f, "nsb_sem_hmholtz"));
  NSB_CHECK(field_ptr(S, bx, cx, field0, &x, "nsb_sem_hmholtz"));
  NSB_REQUIRE(f != x && maxit >= 1, "nsb_sem_hmholtz: bad argument");
  NSB_REQUIRE(S->exchange_ready, "nsb_sem_hmholtz: call nsb_sem_setup_exchange first");
  const nsb_layout_t Lf = brhs->lay, Lx = bx->lay;
  const int64_t fs_f = nf > 1 ? Lf->off[field0 + 1] - Lf->off[field0] : 0;
  const int64_t fs_x = nf > 1 ? Lx->off[field0 + 1] - Lx->off[field0] : 0;
  for (int g = field0 + 1; g < field0 + nf; ++g)
    NSB_REQUIRE(Lf->off[g] - Lf->off[g - 1] == fs_f && Lx->off[g] - Lx->off[g - 1] == fs_x,
                "nsb_sem_hmholtz: fields are not equally spaced");
  nsb_context_t ctx = S->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int64_t n = S->npts;
  const size_t nb = sizeof(double) * n;
  constexpr int kPcgPoll = 4;
  if (!S->pcg_d) NSB_CUDA(cudaMalloc(&S->pcg_d, nb * (3 * kPcgFields + 1) + sizeof(PcgState) + 64));
  double *r = S->pcg_d, *p = r + kPcgFields * n, *w = p + kPcgFields * n, *d = w + kPcgFields * n;
  PcgState *state = reinterpret_cast<PcgState *>(d + n);
  double *red = ctx->hvec_d + 3 * (kMaxK + 8) + 16;
  PcgState *hstate = reinterpret_cast<PcgState *>(ctx->hpin + 3 * (kMaxK + 8) + 32);   // pinned mirror
  static_assert(sizeof(PcgState) <= sizeof(double) * (kMaxK + 8 - 32), "PcgState must fit the pinned scratch");
  const int grid = ctx->num_sms * 8;
  const dim3 g2(grid, nf);
  NSB_CHECK(ensure_partial(ctx, std::max(grid, ctx->num_sms * 8)));
  // setprec: d = mask / dssum(h1 diag(A) + h2 bm1); diag(A) depends on the mesh only and is kept
  if (!S->diagA_d) {
    NSB_CUDA(cudaMalloc(&S->diagA_d, nb));
    helm_diag_kernel<<<blocks_for(n), 256, 0, st>>>(S->g_d, S->bm1_d, S->D_d, S->dim, S->lx, n, 1.0, 0.0, S->diagA_d);
    ctx->launches++;
  }
  axpby_fields_kernel<<<grid, 256, 0, st>>>(d, S->diagA_d, S->bm1_d, h1, h2, n);
  ctx->launches++;
  NSB_CHECK(launch_gs(S, d, 1, 0, 0, nullptr, 0, 0, nullptr));
  recip_kernel<<<grid, 256, 0, st>>>(d, n);
  col2_kernel<<<grid, 256, 0, st>>>(d, S->mask_d, n);
  ctx->launches += 2;
  for (int g = 0; g < nf; ++g) NSB_CUDA(cudaMemsetAsync(x + g * fs_x, 0, nb, st));
  NSB_CUDA(cudaMemsetAsync(p, 0, nb * nf, st));
  pcg_init_kernel<<<g2, 256, 0, st>>>(f, fs_f, r, d, S->mask_d, S->vmult_d, n, nf, ctx->partial_d);
  ctx->launches++;
  NSB_CHECK(reduce_scalars_d(ctx, grid, nf, red));
  pcg_begin_kernel<<<1, 32, 0, st>>>(state, red, nf, maxit);
  ctx->launches++;
  auto poll = [&]() -> int {
    NSB_CUDA(cudaMemcpyAsync(hstate, state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
    NSB_CUDA(cudaStreamSynchronize(st));
    return check_dev_err(ctx);
  };
  for (int it = 1; it <= maxit; ++it) {
    if (nf == 1) pcg_p_kernel<1><<<grid, 256, 0, st>>>(p, r, d, state, n);               // p = d r + beta p
    else if (nf == 2) pcg_p_kernel<2><<<grid, 256, 0, st>>>(p, r, d, state, n);
    else pcg_p_kernel<3><<<grid, 256, 0, st>>>(p, r, d, state, n);
    ctx->launches++;
    S->ax_dotp = ctx->partial_d;                                                         // w = H p and (w_raw, p)
    S->ax_dot_rows = 0;
    int rc = launch_axhelm(S, p, w, nf, n, h1, h2, nullptr, 0, 0, 0, nullptr);
    S->ax_dotp = nullptr;
    NSB_CHECK(rc);
    const int dot_rows = S->ax_dot_rows;
    if (dot_rows > 0) NSB_CHECK(reduce_scalars_d(ctx, dot_rows, nf, red));               // before the dssum reuses nothing of it
    NSB_CHECK(launch_gs(S, w, nf, n, 0, nullptr, 0, 0, nullptr));                        // dssum
    if (dot_rows == 0) {
      pcg_dot_kernel<<<g2, 256, 0, st>>>(w, p, S->vmult_d, n, nf, ctx->partial_d);       // rho = (w, p)_mult
      ctx->launches++;
      NSB_CHECK(reduce_scalars_d(ctx, grid, nf, red));
    }
    pcg_alpha_kernel<<<1, 32, 0, st>>>(state, red, nf);
    if (nf == 1) pcg_update_kernel<1><<<grid, 256, 0, st>>>(x, fs_x, r, p, w, d, S->vmult_d, state, n, ctx->partial_d);
    else if (nf == 2) pcg_update_kernel<2><<<grid, 256, 0, st>>>(x, fs_x, r, p, w, d, S->vmult_d, state, n, ctx->partial_d);
    else pcg_update_kernel<3><<<grid, 256, 0, st>>>(x, fs_x, r, p, w, d, S->vmult_d, state, n, ctx->partial_d);
    ctx->launches += 2;
    NSB_CHECK(reduce_scalars_d(ctx, grid, nf, red));
    pcg_conv_kernel<<<1, 32, 0, st>>>(state, red, nf, tol);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    if (it % kPcgPoll == 0 || it == maxit) {
      NSB_CHECK(poll());
      if (hstate->all_done) break;
    }
  }
  NSB_CHECK(poll());
  for (int g = 0; g < nf; ++g) {
    if (iters) iters[g] = hstate->itf[g];
    if (res) res[g] = hstate->r0[g] > 0.0 ? hstate->rn[g] / hstate->r0[g] : 0.0;
  }
  return NSB_OK;
}

extern "C" int nsb_sem_hmholtz(nsb_sem_t S, nsb_basis_t brhs, int crhs, nsb_basis_t bx, int cx, int field,
                               double h1, double h2, double tol, int maxit, int *iters, double *res) {
  return nsb_sem_hmholtz_vec(S, brhs, crhs, bx, cx, field, 1, h1, h2, tol, maxit, iters, res);
}

// ------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------
extern "C" int nsb_op_create_sem(nsb_sem_t S, int nfields_apply, double alpha, double beta, double h1,
                                 double h2, const double *cx, const double *cy, const double *cz,
                                 nsb_op_t *out) {
  NSB_REQUIRE(S && out && nfields_apply >= 1, "nsb_op_create_sem: bad argument");
  NSB_REQUIRE(S->exchange_ready, "nsb_op_create_sem: call nsb_sem_setup_exchange first");
  nsb_op_t op = new nsb_op_s();
  op->kind = 0;
  op->sem = S;
  op->nfields_apply = nfields_apply;
  op->alpha = alpha;
  op->beta = beta;
  op->h1 = h1;
  op->h2 = h2;
  if (cx) {
    NSB_REQUIRE(cy && (S->dim == 2 || cz), "nsb_op_create_sem: incomplete convecting velocity");
    nsb_context_t ctx = S->ctx;
    cudaSetDevice(ctx->device);
    const size_t nb = sizeof(double) * S->npts;
    double *c_in = nullptr, *wq_d = nullptr;
    NSB_CUDA(cudaMalloc(&c_in, nb * 3));
    NSB_CUDA(cudaMalloc(&wq_d, sizeof(double) * S->lx));
    NSB_CUDA(cudaMalloc(&op->c_d, nb * S->dim));
    NSB_CUDA(cudaMemcpyAsync(c_in, cx, nb, cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaMemcpyAsync(c_in + S->npts, cy, nb, cudaMemcpyHostToDevice, ctx->stream));
    if (S->dim == 3) NSB_CUDA(cudaMemcpyAsync(c_in + 2 * S->npts, cz, nb, cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaMemcpyAsync(wq_d, S->w_h.data(), sizeof(double) * S->lx, cudaMemcpyHostToDevice, ctx->stream));
    conv_coeff_kernel<<<blocks_for(S->npts), 256, 0, ctx->stream>>>(
        S->rst_d, c_in, c_in + S->npts, c_in + 2 * S->npts, wq_d, S->dim, S->lx, S->npts, op->c_d);
    ctx->launches++;
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(c_in);
    cudaFree(wq_d);
  }
  *out = op;
  return NSB_OK;
}

extern "C" int nsb_op_create_host(nsb_layout_t L, nsb_host_matvec_fn fn, void *user, nsb_op_t *out) {
  NSB_REQUIRE(L && fn && out, "nsb_op_create_host: NULL argument");
  nsb_op_t op = new nsb_op_s();
  op->kind = 1;
  op->lay = L;
  op->fn = fn;
  op->user = user;
  op->hin.assign(L->nfields, nullptr);
  op->hout.assign(L->nfields, nullptr);
  cudaSetDevice(L->ctx->device);
  for (int f = 0; f < L->nfields; ++f) {
    const size_t nb = sizeof(double) * (size_t)(L->hlen[f] > 0 ? L->hlen[f] : 1);   // host-side (element-local) size
    NSB_CUDA(cudaMallocHost((void **)&op->hin[f], nb));
    NSB_CUDA(cudaMallocHost((void **)&op->hout[f], nb));
  }
  *out = op;
  return NSB_OK;
}

// out = outer(inner(in)): the reference's composite maps -- transient_growth_map = adjoint(forward(q))
// (core/matvec.f90:478-495), newton_linearized_map etc. are built this way from the basic solvers.
// A linear host operator (M(a x) = a M(x), %time included) lets the Arnoldi loop hand the un-normalised vector
// over while its last sweep is still running and scale the result on the device (nsb_arnoldi).
extern "C" int nsb_op_set_linear(nsb_op_t op, int linear) {
  NSB_REQUIRE(op, "nsb_op_set_linear: NULL operator");
  op->linear = linear != 0;
  return NSB_OK;
}

extern "C" int nsb_op_create_compose(nsb_layout_t layout, nsb_op_t outer, nsb_op_t inner, nsb_op_t *out) {
  NSB_REQUIRE(layout && outer && inner && out, "nsb_op_create_compose: NULL argument");
  nsb_op_t op = new nsb_op_s();
  op->kind = 2;
  op->lay = layout;
  op->outer = outer;
  op->inner = inner;
  int r = nsb_basis_create(layout, 1, &op->tmp);
  if (r != NSB_OK) {
    delete op;
    return r;
  }
  *out = op;
  return NSB_OK;
}

// out = alpha A(in) + beta B(in), NULL = the identity: LightKrylov's axpby_linop / identity_linop as the reference
// uses them for the resolvent's S = I - exp(TL) (core/linear_operators.f90:364-403) [UPSTREAM-RECALL for the type
// itself], and the legacy maps built from the basic solvers with k_sub2 / k_cmult: newton_linearized_map
// = exp(TL) - I (core/matvec.f90:520-541) and ts_force_sensitivity_map = I - exp(TL)+ (core/matvec.f90:499-516).
// Every field of the vector and %time take part, like k_sub2 (core/krylov_subspace.f90:117-128).
extern "C" int nsb_op_create_axpby(nsb_layout_t layout, nsb_op_t A, nsb_op_t B, double alpha, double beta,
                                   nsb_op_t *out) {
  NSB_REQUIRE(layout && out, "nsb_op_create_axpby: NULL argument");
  nsb_op_t op = new nsb_op_s();
  op->kind = 5;
  op->lay = layout;
  op->outer = A;
  op->inner = B;
  op->alpha = alpha;
  op->beta = beta;
  if (B) {
    int r = nsb_basis_create(layout, 1, &op->tmp);
    if (r != NSB_OK) {
      delete op;
      return r;
    }
  }
  *out = op;
  return NSB_OK;
}

// forward_finite_difference_map (core/matvec.f90:246-379): the linearised forward map approximated by finite
// differences of a NONLINEAR map F (the reference's nonlinear Nek stepper; here any operator handle, typically a host
// callback) about the base state X:
//     f = (1/eps0) sum_i coef_i F(X + amp_i eps0 q),   eps0 = 1e-6 |X|   (k_norm, :276-277)
// order 2: amp = (1, -1), coef = (1, -1)/2 ; order 4: amp = (1, -1, 2, -2), coef = (8, -8, -1, 1)/12  (:279-289).
// Same order of operations as the reference: pert = amp q (:322-323), X + pert (:326-327), work = coef F (:367-368),
// f += work (:369), f *= 1/eps0 (:374).  The base state stays owned by the caller and is read at every application
// (newton_krylov updates it between linear solves).
extern "C" int nsb_op_create_frechet_fd(nsb_layout_t layout, nsb_op_t F, nsb_basis_t base, int col_base, int order,
                                        nsb_op_t *out) {
  NSB_REQUIRE(layout && F && base && out, "nsb_op_create_frechet_fd: NULL argument");
  NSB_REQUIRE(order == 2 || order == 4, "nsb_op_create_frechet_fd: findiff_order = %d (2 or 4)", order);
  NSB_REQUIRE(base->lay == layout, "nsb_op_create_frechet_fd: base state has another layout");
  NSB_REQUIRE(col_base >= 0 && col_base < base->ncols, "nsb_op_create_frechet_fd: column out of range");
  nsb_op_t op = new nsb_op_s();
  op->kind = 6;
  op->lay = layout;
  op->outer = F;
  op->base_b = base;
  op->base_c = col_base;
  op->fd_order = order;
  op->h1 = 1e-6;   // epsilon_base (core/main.f90:16)
  int r = nsb_basis_create(layout, 2, &op->tmp);
  if (r != NSB_OK) {
    delete op;
    return r;
  }
  *out = op;
  return NSB_OK;
}

// epsilon_base of the reference (core/main.f90:16, default 1e-6; core/linear_operators.f90:192): eps0 = epsilon_base |X|
extern "C" int nsb_op_frechet_set_epsilon(nsb_op_t op, double epsilon_base) {
  NSB_REQUIRE(op && op->kind == 6, "nsb_op_frechet_set_epsilon: not a finite-difference operator");
  NSB_REQUIRE(epsilon_base > 0.0, "nsb_op_frechet_set_epsilon: epsilon_base = %g", epsilon_base);
  op->h1 = epsilon_base;
  return NSB_OK;
}

namespace {
int frechet_fd_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  static const double amp2[2] = {1.0, -1.0}, coef2[2] = {0.5, -0.5};
  static const double amp4[4] = {1.0, -1.0, 2.0, -2.0}, coef4[4] = {8.0 / 12.0, -8.0 / 12.0, -1.0 / 12.0, 1.0 / 12.0};
  const double *amp = op->fd_order == 2 ? amp2 : amp4, *coef = op->fd_order == 2 ? coef2 : coef4;
  double xnorm = 0.0;
  NSB_CHECK(nsb_vec_norm(op->base_b, op->base_c, &xnorm));
  NSB_REQUIRE(xnorm > 0.0, "forward_finite_difference_map: the base state has zero norm");
  const double eps0 = op->h1 * xnorm;   // op->h1 = epsilon_base
  NSB_CHECK(nsb_vec_zero(bout, cout));
  for (int i = 0; i < op->fd_order; ++i) {
    NSB_CHECK(nsb_vec_copy(op->tmp, 0, bin, cin));
    NSB_CHECK(nsb_vec_scal(op->tmp, 0, amp[i] * eps0));
    NSB_CHECK(nsb_vec_add2(op->tmp, 0, op->base_b, op->base_c));
    NSB_CHECK(nsb_op_apply(op->outer, op->tmp, 0, op->tmp, 1));
    NSB_CHECK(nsb_vec_scal(op->tmp, 1, coef[i]));
    NSB_CHECK(nsb_vec_add2(bout, cout, op->tmp, 1));
  }
  return nsb_vec_scal(bout, cout, 1.0 / eps0);
}
}  // namespace

extern "C" int nsb_op_destroy(nsb_op_t op) {
  if (!op) return NSB_OK;
  if (op->sem) clear_step_graphs(op->sem->ctx);
  else if (op->lay) clear_step_graphs(op->lay->ctx);
  if (op->tmp) nsb_basis_destroy(op->tmp);
  if (op->c_d) cudaFree(op->c_d);
  for (double *p : op->hin) cudaFreeHost(p);
  for (double *p : op->hout) cudaFreeHost(p);
  delete op;
  return NSB_OK;
}

extern "C" int nsb_op_count(nsb_op_t op, int64_t *n) {
  NSB_REQUIRE(op && n, "nsb_op_count: NULL argument");
  *n = op->napply;
  return NSB_OK;
}

extern "C" int nsb_op_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout) {
  NSB_REQUIRE(op && bin && bout, "nsb_op_apply: NULL argument");
  NSB_REQUIRE(cin >= 0 && cin < bin->ncols && cout >= 0 && cout < bout->ncols,
              "nsb_op_apply: column out of range");
  NSB_REQUIRE(bin->lay == bout->lay, "nsb_op_apply: different layouts");
  NSB_REQUIRE(!(bin == bout && cin == cout), "nsb_op_apply: in-place application is not supported");
  op->napply++;
  if (op->kind == 3) return nsb::stepper_apply(op, bin, cin, bout, cout);
  if (op->kind == 4) return nsb::ns_stepper_apply(op, bin, cin, bout, cout);
  if (op->kind == 2) {
    NSB_REQUIRE(bin->lay == op->lay, "nsb_op_apply: composite operator built for another layout");
    NSB_CHECK(nsb_op_apply(op->inner, bin, cin, op->tmp, 0));
    return nsb_op_apply(op->outer, op->tmp, 0, bout, cout);
  }
  if (op->kind == 6) {
    NSB_REQUIRE(bin->lay == op->lay, "nsb_op_apply: finite-difference operator built for another layout");
    return frechet_fd_apply(op, bin, cin, bout, cout);
  }
  if (op->kind == 5) {
    NSB_REQUIRE(bin->lay == op->lay, "nsb_op_apply: axpby operator built for another layout");
    if (op->outer) NSB_CHECK(nsb_op_apply(op->outer, bin, cin, bout, cout));
    else NSB_CHECK(nsb_vec_copy(bout, cout, bin, cin));
    if (!op->inner) return nsb_vec_axpby(bout, cout, op->alpha, bin, cin, op->beta, 0);
    NSB_CHECK(nsb_op_apply(op->inner, bin, cin, op->tmp, 0));
    return nsb_vec_axpby(bout, cout, op->alpha, op->tmp, 0, op->beta, 0);
  }
  if (op->kind == 1) {
    NSB_REQUIRE(bin->lay == op->lay, "nsb_op_apply: host operator built for another layout");
    nsb_layout_t L = bin->lay;
    std::vector<const double *> pin(L->nfields);
    std::vector<double *> pout(L->nfields), pdl(L->nfields);
    for (int f = 0; f < L->nfields; ++f) {
      pin[f] = op->hin[f];
      pdl[f] = op->hin[f];
      pout[f] = op->hout[f];
    }
    double tin = 0.0, tout = 0.0;
    NSB_CHECK(nsb_vec_download(bin, cin, pdl.data(), &tin));
    int r = op->fn(op->user, pin.data(), tin, pout.data(), &tout);
    if (r != 0) {
      set_error("nsb_op_apply: host matvec callback returned %d", r);
      return NSB_EINVAL;
    }
    return nsb_vec_upload(bout, cout, pout.data(), tout);
  }
  nsb_sem_t S = op->sem;
  nsb_layout_t L = bin->lay;
  NSB_REQUIRE(op->nfields_apply <= L->nfields, "nsb_op_apply: operator covers %d fields, layout has %d",
              op->nfields_apply, L->nfields);
  if (L->c0_sem) {
    NSB_REQUIRE(L->c0_sem == S, "nsb_op_apply: the C0 layout was built on another mesh");
    return c0_apply_sem(op, bin, cin, bout, cout);
  }
  // the applied fields are equally long, hence equally spaced inside a column: one batched
  // launch per kernel (grid.y = field) and one interface exchange for all components
  const int nfa = op->nfields_apply;
  const int64_t fstride = nfa > 1 ? L->off[1] - L->off[0] : 0;
  bool uniform = nfa <= S->ns_fields;
  for (int f = 1; f < nfa; ++f) uniform = uniform && (L->off[f] - L->off[f - 1] == fstride);
  const int step = uniform ? nfa : 1;
  for (int f = 0; f < nfa; f += step) {
    double *u, *w;
    for (int g = f; g < f + step; ++g) {
      NSB_CHECK(field_ptr(S, bin, cin, g, &u, "nsb_op_apply"));
      NSB_CHECK(field_ptr(S, bout, cout, g, &w, "nsb_op_apply"));
    }
    NSB_CHECK(field_ptr(S, bin, cin, f, &u, "nsb_op_apply"));
    NSB_CHECK(field_ptr(S, bout, cout, f, &w, "nsb_op_apply"));
    if (S->nslab > 1 && !L->ctx->prof && S->nshared > 0) {
      NSB_CHECK(launch_ax_gs_pipelined(S, u, w, step, fstride, op->h1, op->h2, op->c_d, op->alpha, op->beta));
    } else {   // one slab, or per-kernel profiling (the roofline table keeps axhelm and gather-scatter apart)
      NSB_CHECK(launch_axhelm(S, u, w, step, fstride, op->h1, op->h2, op->c_d, 1, op->alpha, op->beta, S->bmask_d));
      NSB_CHECK(launch_gs(S, w, step, fstride, 1, u, op->alpha, op->beta, S->bmask_d));
    }
  }
  // the rows outside the operator are carried through unchanged: %time and the remaining fields
  // (pressure, scalars)
  cudaStream_t s = L->ctx->stream;
  NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->time_row, bin->col(cin) + L->time_row, sizeof(double),
                           cudaMemcpyDeviceToDevice, s));
  for (int f = op->nfields_apply; f < L->nfields; ++f)
    if (L->len[f] > 0)
      NSB_CUDA(cudaMemcpyAsync(bout->col(cout) + L->off[f], bin->col(cin) + L->off[f], sizeof(double) * L->len[f],
                               cudaMemcpyDeviceToDevice, s));
  return NSB_OK;
}
