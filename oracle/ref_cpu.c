/*
 * ref_cpu.c -- C restatement of the reference's CPU path for the Arnoldi hot loop.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA kernels and the "port" CPU baseline that
 * bench.py times beside the GPU path.  The reference itself (Fortran 90 + Nek5000 + MPI) cannot be
 * built in this image (no Fortran compiler, Nek5000/LightKrylov not vendored), so this file
 * restates, loop for loop:
 *   - update_hessenberg_matrix: core/krylov_decomposition.f90:150-186 with the
 *     k_copy / k_dot / k_cmult / k_sub2 sweep structure (core/krylov_subspace.f90:26-161)
 *   - k_dot = sum over components of glsc3(a, bm1s, b)  (core/krylov_subspace.f90:40-47)
 *   - [UPSTREAM-RECALL, Nek5000 not in /root/reference] glsc3 (math.f), axhelm with mxm-ordered
 *     sum factorisation (hmholtz.f), dssum as gather-scatter (gslib), col2 mask.
 * OpenMP threads stand in for MPI ranks (points / elements are split across threads).
 * Parity: unpinned by reference tests (it has none); checked against oracle/sem.py + krylov.py,
 * which are pinned to the reference's field-file fixtures.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* the launcher may have pinned OMP_NUM_THREADS to 1 (torchrun does for N > 1 ranks) */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* math.f glsc3: tmp += a(i)*b(i)*mult(i); then gop(+) -- the gop is the OpenMP reduction */
double ref_glsc3(const double *a, const double *b, const double *mult, int64_t n) {
  double tmp = 0.0;
#pragma omp parallel for reduction(+ : tmp) schedule(static)
  for (int64_t i = 0; i < n; ++i) tmp += a[i] * b[i] * mult[i];
  return tmp;
}

static void ref_copy(double *a, const double *b, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) a[i] = b[i];
}
static void ref_cmult(double *a, double c, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) a[i] *= c;
}
static void ref_sub2(double *a, const double *b, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) a[i] -= b[i];
}

/* k_dot over ncomp velocity-like components of npts points each (same bm1s for each) */
double ref_k_dot(const double *p, const double *q, const double *bm1s, int64_t npts, int ncomp) {
  double alpha = 0.0;
  for (int c = 0; c < ncomp; ++c) alpha += ref_glsc3(p + c * npts, bm1s, q + c * npts, npts);
  return alpha;
}

/* update_hessenberg_matrix (core/krylov_decomposition.f90:150-186).
 * Q: k vectors of length n = ncomp*npts, contiguous with stride ldq; f in/out; wrk scratch;
 * h[0..k] = column k of H. */
void ref_update_hessenberg(const double *Q, int64_t ldq, double *f, double *wrk, const double *bm1s,
                           int64_t npts, int ncomp, int k, double *h) {
  const int64_t n = npts * ncomp;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < k; ++i) {
      ref_copy(wrk, Q + (int64_t)i * ldq, n);                   /* k_copy(wrk, q(i)) */
      double alpha = ref_k_dot(f, wrk, bm1s, npts, ncomp);      /* k_dot */
      ref_cmult(wrk, alpha, n);                                 /* k_cmult */
      ref_sub2(f, wrk, n);                                      /* k_sub2 */
      if (pass == 0) h[i] = alpha; else h[i] += alpha;
    }
  double alpha = sqrt(ref_k_dot(f, f, bm1s, npts, ncomp));      /* k_normalize */
  ref_cmult(f, 1.0 / alpha, n);
  h[k] = alpha;
}

/* mxm-ordered element kernels: c(i,j) = sum_l a(i,l) b(l,j), l ascending (mxm.f) */
static inline void mxm(const double *a, int n1, const double *b, int n2, double *c, int n3) {
  for (int j = 0; j < n3; ++j)
    for (int i = 0; i < n1; ++i) {
      double s = 0.0;
      for (int l = 0; l < n2; ++l) s += a[i + n1 * l] * b[l + n2 * j];
      c[i + n1 * j] = s;
    }
}

/* hmholtz.f axhelm, 3-D, deformed branch, constant h1/h2.  D = dxm1 (column-major, D[i + lx*j]),
 * Dt = its transpose.  g = [6][npts].  Elements are independent -> parallel over elements. */
void ref_axhelm3d(double *au, const double *u, const double *g, const double *bm1, const double *D,
                  int lx, int64_t nel, double h1, double h2) {
  const int n2 = lx * lx, n3 = lx * lx * lx;
  const int64_t npts = nel * n3;
  double Dt[32 * 32];
  for (int i = 0; i < lx; ++i)
    for (int j = 0; j < lx; ++j) Dt[i + lx * j] = D[j + lx * i];
#pragma omp parallel
  {
    double *dudr = (double *)malloc(sizeof(double) * n3 * 6);
    double *duds = dudr + n3, *dudt = duds + n3, *t1 = dudt + n3, *t2 = t1 + n3, *t3 = t2 + n3;
#pragma omp for schedule(static)
    for (int64_t e = 0; e < nel; ++e) {
      const double *ue = u + e * n3;
      double *we = au + e * n3;
      mxm(D, lx, ue, lx, dudr, n2);                                        /* dudr = D u       */
      for (int iz = 0; iz < lx; ++iz) mxm(ue + iz * n2, lx, Dt, lx, duds + iz * n2, lx); /* u Dt */
      mxm(ue, n2, Dt, lx, dudt, lx);
      for (int p = 0; p < n3; ++p) {
        const int64_t q = e * n3 + p;
        const double g1 = g[q], g2 = g[npts + q], g3 = g[2 * npts + q], g4 = g[3 * npts + q],
                     g5 = g[4 * npts + q], g6 = g[5 * npts + q];
        const double ur = dudr[p], us = duds[p], ut = dudt[p];
        t1[p] = h1 * (g1 * ur + g4 * us + g5 * ut);
        t2[p] = h1 * (g2 * us + g4 * ur + g6 * ut);
        t3[p] = h1 * (g3 * ut + g5 * ur + g6 * us);
      }
      mxm(Dt, lx, t1, lx, dudr, n2);                                       /* tm1 = Dt tmp1    */
      for (int iz = 0; iz < lx; ++iz) mxm(t2 + iz * n2, lx, D, lx, duds + iz * n2, lx);
      mxm(t3, n2, D, lx, dudt, lx);
      for (int p = 0; p < n3; ++p) {
        double v = dudr[p] + duds[p] + dudt[p];
        if (h2 != 0.0) v += h2 * bm1[e * n3 + p] * ue[p];                   /* addcol4 */
        we[p] = v;
      }
    }
    free(dudr);
  }
}

/* dssum as gather-scatter over unique nodes (CSR lists built by the caller from glo_num) */
void ref_dssum(double *u, const int64_t *off, const int32_t *idx, int64_t nnodes) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < nnodes; ++n) {
    double s = 0.0;
    for (int64_t q = off[n]; q < off[n + 1]; ++q) s += u[idx[q]];
    for (int64_t q = off[n]; q < off[n + 1]; ++q) u[idx[q]] = s;
  }
}

static void ref_col2(double *a, const double *b, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) a[i] *= b[i];
}

/* synthetic matvec of SURVEY.md section 8d per component:
 *   f = alpha q + beta * binvm1 * mask * dssum(axhelm(q))   (ax = axhelm + dssum + col2 mask) */
void ref_matvec(double *f, const double *q, double *tmp, const double *g, const double *bm1,
                const double *binv, const double *mask, const double *D, int lx, int64_t nel,
                int ncomp, const int64_t *off, const int32_t *idx, int64_t nnodes, double h1,
                double h2, double alpha, double beta) {
  const int64_t npts = nel * lx * lx * lx;
  for (int c = 0; c < ncomp; ++c) {
    ref_axhelm3d(tmp, q + c * npts, g, bm1, D, lx, nel, h1, h2);
    ref_dssum(tmp, off, idx, nnodes);
    ref_col2(tmp, mask, npts);
    ref_col2(tmp, binv, npts);
    double *fc = f + c * npts;
    const double *qc = q + c * npts;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < npts; ++i) fc[i] = alpha * qc[i] + beta * tmp[i];
  }
}

/* arnoldi_factorization (core/krylov_decomposition.f90:68-96), steps mstart..mend (0-based).
 * Q has ldq >= n and at least mend+2 columns; H column-major (ldh). */
void ref_arnoldi(double *Q, int64_t ldq, double *H, int ldh, int mstart, int mend, double *wrk,
                 double *tmp, const double *bm1s, const double *g, const double *bm1,
                 const double *binv, const double *mask, const double *D, int lx, int64_t nel,
                 int ncomp, const int64_t *off, const int32_t *idx, int64_t nnodes, double h1,
                 double h2, double alpha, double beta) {
  const int64_t npts = nel * lx * lx * lx;
  for (int m = mstart; m <= mend; ++m) {
    double *f = Q + (int64_t)(m + 1) * ldq;
    ref_matvec(f, Q + (int64_t)m * ldq, tmp, g, bm1, binv, mask, D, lx, nel, ncomp, off, idx, nnodes,
               h1, h2, alpha, beta);
    ref_update_hessenberg(Q, ldq, f, wrk, bm1s, npts, ncomp, m + 1, H + (int64_t)m * ldh);
  }
}
