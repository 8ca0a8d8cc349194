/* A compiled C host that drives the library the way the reference's Fortran host does (through the C ABI
 * only, operator on the HOST as a callback -- the reference's nek_advance lives there):
 *
 *   nsb_init -> nsb_layout_create / nsb_layout_set_weight -> nsb_basis_create -> nsb_op_create_host
 *   -> nsb_arnoldi (arnoldi_factorization, core/krylov_decomposition.f90:2-99)
 *   -> nsb_ts_gmres (ts_gmres, core/newton_krylov.f90:170-299) with LAPACK injected through nsb_set_lapack,
 *      exactly as a Fortran host passes c_funloc(dgeev) ... (core/lapack_wrapper.f90)
 *
 * It prints H, the GMRES solution's BM1 norm and three of its entries; tests/test_c_host.py compares them with
 * tests/golden/c_host_arnoldi.json (written by tests/golden/make_c_host_golden.py from the numpy oracle).
 *
 * usage: arnoldi_host <path to a LAPACK shared library> <symbol prefix, e.g. scipy_>
 */
#include "nekstab_b200.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NPF 5000 /* points per field */
#define NF 2
#define K 12
#define KS 15

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != NSB_OK) {                                                         \
      fprintf(stderr, "%s failed with %d: %s\n", #call, rc_, nsb_last_error()); \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

static long ncalls = 0;

/* out_f[i] = d_i in_f[i] + 0.05 in_f[i+1] - 0.03 in_f[i-1] + 0.02 in_{1-f}[i]   (periodic in i; non-symmetric) */
static int host_matvec(void *user, const double *const *in, double tin, double **out, double *tout) {
  int f, i;
  (void)user;
  for (f = 0; f < NF; ++f)
    for (i = 0; i < NPF; ++i) {
      const double d = 0.9 - 0.8 * ((double)i / NPF);
      out[f][i] = d * in[f][i] + 0.05 * in[f][(i + 1) % NPF] - 0.03 * in[f][(i + NPF - 1) % NPF] + 0.02 * in[1 - f][i];
    }
  *tout = tin;
  ++ncalls;
  return 0;
}

int main(int argc, char **argv) {
  nsb_context_t ctx = 0;
  nsb_layout_t lay = 0;
  nsb_basis_t Q = 0, W = 0;
  nsb_op_t op = 0;
  int64_t len[NF] = {NPF, NPF};
  int in_dot[NF] = {1, 1};
  static double w[NPF], seed[NF][NPF], rhs[NF][NPF], sol[NF][NPF];
  static double H[(K + 1) * K], hist[8];
  const double *wp[NF], *fp[NF];
  double *sp[NF];
  double nrm = 0.0, t = 0.0;
  int i, j, f, calls = 0, nh = 0;
  void *lap, *sym[4];
  const char *names[4] = {"dgeev_", "dgees_", "dtrsen_", "dgels_"};
  char buf[128];

  if (argc < 3) return 64;
  lap = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!lap) {
    fprintf(stderr, "dlopen %s: %s\n", argv[1], dlerror());
    return 3;
  }
  for (i = 0; i < 4; ++i) {
    snprintf(buf, sizeof buf, "%s%s", argv[2], names[i]);
    sym[i] = dlsym(lap, buf);
    if (!sym[i]) {
      fprintf(stderr, "symbol %s missing\n", buf);
      return 3;
    }
  }
  CHECK(nsb_set_lapack(sym[0], sym[1], sym[2], sym[3]));

  CHECK(nsb_init(0, 0, 1, 0, &ctx));
  CHECK(nsb_layout_create(ctx, NF, len, in_dot, 0, &lay));
  for (i = 0; i < NPF; ++i) w[i] = 1.0 + 0.5 * cos(0.01 * i);
  wp[0] = wp[1] = w;
  CHECK(nsb_layout_set_weight(lay, wp));
  CHECK(nsb_basis_create(lay, KS + 2, &Q));
  CHECK(nsb_basis_create(lay, 2, &W));
  CHECK(nsb_op_create_host(lay, host_matvec, 0, &op));
  for (f = 0; f < NF; ++f)
    for (i = 0; i < NPF; ++i) {
      seed[f][i] = sin(0.37 * i + 1.3 * f) + 0.25 * cos(0.011 * i * (f + 1));
      rhs[f][i] = cos(0.21 * i - 0.7 * f);
    }
  fp[0] = seed[0];
  fp[1] = seed[1];
  CHECK(nsb_vec_upload(Q, 0, fp, 0.0));
  CHECK(nsb_vec_normalize(Q, 0, &nrm));
  printf("seed_norm %.17g\n", nrm);
  CHECK(nsb_arnoldi(Q, op, 0, K - 1, NSB_ORTH_CGS2, H, K + 1));
  for (j = 0; j < K; ++j)
    for (i = 0; i < K + 1; ++i) printf("H %d %d %.17g\n", i, j, H[j * (K + 1) + i]);
  printf("matvec_calls_arnoldi %ld\n", ncalls);

  fp[0] = rhs[0];
  fp[1] = rhs[1];
  CHECK(nsb_vec_upload(W, 0, fp, 0.0));
  CHECK(nsb_ts_gmres(Q, op, W, 0, W, 1, 4, KS, 1e-24, NSB_ORTH_CGS2, &calls, hist, &nh));
  CHECK(nsb_vec_norm(W, 1, &nrm));
  sp[0] = sol[0];
  sp[1] = sol[1];
  CHECK(nsb_vec_download(W, 1, sp, &t));
  printf("gmres_calls %d\n", calls);
  printf("gmres_restarts %d\n", nh);
  printf("sol_norm %.17g\n", nrm);
  printf("sol %.17g %.17g %.17g\n", sol[0][0], sol[0][NPF / 2], sol[1][NPF - 1]);
  printf("last_residual %.3e\n", hist[nh - 1]);

  CHECK(nsb_op_destroy(op));
  CHECK(nsb_basis_destroy(W));
  CHECK(nsb_basis_destroy(Q));
  CHECK(nsb_layout_destroy(lay));
  CHECK(nsb_finalize(ctx));
  return 0;
}
