// Krylov checkpoint / restart wire formats of the reference behind the C ABI:
//   HES<session>%04d      Hessenberg matrix, list-directed text, row by row
//                         (writer core/eigensolvers.f90:837-843, restart reader :246-266)
//   KRY<session>0.f%05d   Krylov vectors as Nek5000 field files (outpost2, core/eigensolvers.f90:803-809;
//                         read back by load_files, core/IO.f90:11-72)
// Host code only: files are parsed on the host and the fields go to the device basis through nsb_vec_upload.
// [UPSTREAM-RECALL] Nek5000 field-file layout (prepost.f mfo_write_hdr / ic.f mfi): 132-byte ASCII header
// '#std wdsize nx ny nz nelo nelg time istep fid nfiles rdcode', float32 endian tag 6.54321, int32 global
// element ids, then per group (X, U: ndim components element by element; P, T: one block each).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nsb_internal.h"

using namespace nsb;

namespace {

bool read_file(const char *path, std::vector<unsigned char> &buf) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == buf.size();
}

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

}  // namespace

// core/eigensolvers.f90:837-843: write(67,*) ((H(i,j), j=1,k), i=1,k+1) -- full double precision so that a
// restart reproduces the factorisation.
extern "C" int nsb_hessenberg_write(const char *path, const double *H, int ldh, int k) {
  NSB_REQUIRE(path && H && k >= 1 && ldh >= k + 1, "nsb_hessenberg_write: bad argument");
  FILE *f = fopen(path, "w");
  NSB_REQUIRE(f, "nsb_hessenberg_write: cannot open %s", path);
  int col = 0;
  for (int i = 0; i < k + 1; ++i)
    for (int j = 0; j < k; ++j) {
      fprintf(f, "  %.17g", H[(size_t)j * ldh + i]);
      if (++col % 3 == 0) fputc('\n', f);
    }
  if (col % 3) fputc('\n', f);
  fclose(f);
  return NSB_OK;
}

// core/eigensolvers.f90:246-266: the file holds (mstart+1) x mstart values row by row (any whitespace layout,
// Fortran D exponents accepted); the leading block of the (k_dim+1) x k_dim matrix H is filled, subsampled to
// k_dim columns when k_dim < mstart like the reference.
extern "C" int nsb_hessenberg_read(const char *path, int k_dim, int mstart, double *H, int ldh) {
  NSB_REQUIRE(path && H && k_dim >= 1 && mstart >= 1 && ldh >= k_dim + 1, "nsb_hessenberg_read: bad argument");
  std::vector<unsigned char> buf;
  NSB_REQUIRE(read_file(path, buf), "nsb_hessenberg_read: cannot read %s", path);
  buf.push_back(0);
  for (auto &c : buf)
    if (c == 'D' || c == 'd') c = 'E';
  std::vector<double> vals;
  const char *p = (const char *)buf.data();
  for (;;) {
    char *end = nullptr;
    const double v = strtod(p, &end);
    if (end == p) break;
    vals.push_back(v);
    p = end;
    while (*p == ',') ++p;
  }
  NSB_REQUIRE(vals.size() == (size_t)(mstart + 1) * mstart, "nsb_hessenberg_read: %s holds %zu values, expected %zu",
              path, vals.size(), (size_t)(mstart + 1) * mstart);
  for (int j = 0; j < k_dim; ++j)
    for (int i = 0; i < k_dim + 1; ++i) H[(size_t)j * ldh + i] = 0.0;
  const int r = std::min(mstart + 1, k_dim + 1), c = std::min(mstart, k_dim);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) H[(size_t)j * ldh + i] = vals[(size_t)i * mstart + j];
  return NSB_OK;
}

// One Nek field file -> column `col` of the basis.  Velocity components go to layout fields ufield0 ..
// ufield0 + ndim - 1, pressure to pfield, temperature to tfield (either may be -1: not wanted / not in the
// layout; a pressure block is accepted only when the layout's pressure field has the velocity mesh's size --
// Nek writes pressure on mesh 1 -- pass pfield = -1 for PN-PN-2 layouts); the header's time is returned, %time of
// the column stays zero like load_files leaves it.  lglel (1-based global element ids of this rank's elements, Nek's
// LGLEL; NULL: local element e is the e-th element of the file) selects the rank's elements from the file.
extern "C" int nsb_fld_read_into(nsb_basis_t B, int col, const char *path, const int64_t *lglel, int64_t nel_local,
                                 int ufield0, int pfield, int tfield, double *time_out) {
  NSB_REQUIRE(B && path && nel_local >= 1, "nsb_fld_read_into: bad argument");
  NSB_REQUIRE(col >= 0 && col < B->ncols, "nsb_fld_read_into: column %d out of range", col);
  std::vector<unsigned char> raw;
  NSB_REQUIRE(read_file(path, raw) && raw.size() > 136, "nsb_fld_read_into: cannot read %s", path);
  char hdr[133];
  memcpy(hdr, raw.data(), 132);
  hdr[132] = 0;
  char tag[8] = "", rd[16] = "";
  int wd = 0, nx = 0, ny = 0, nz = 0, istep = 0, fid = 0, nfiles = 0;
  long long nelo = 0, nelg = 0;
  double time = 0.0;
  const int got = sscanf(hdr, "%4s %d %d %d %d %lld %lld %lf %d %d %d %15s", tag, &wd, &nx, &ny, &nz, &nelo, &nelg, &time,
                         &istep, &fid, &nfiles, rd);
  NSB_REQUIRE(got >= 11 && strcmp(tag, "#std") == 0 && (wd == 4 || wd == 8), "nsb_fld_read_into: %s is not a Nek field file", path);
  float etag;
  memcpy(&etag, raw.data() + 132, 4);
  bool swap = false;
  if (std::fabs(etag - 6.54321f) > 1e-5f) {
    uint32_t u;
    memcpy(&u, raw.data() + 132, 4);
    u = bswap32(u);
    memcpy(&etag, &u, 4);
    NSB_REQUIRE(std::fabs(etag - 6.54321f) < 1e-5f, "nsb_fld_read_into: %s: bad endian tag", path);
    swap = true;
  }
  const int ndim = nz > 1 ? 3 : 2;
  const int64_t npt = (int64_t)nx * ny * nz;
  nsb_layout_t L = B->lay;
  // element map of the file
  std::vector<int64_t> pos_of_local(nel_local);
  {
    const unsigned char *em = raw.data() + 136;
    NSB_REQUIRE(raw.size() >= 136 + 4 * (size_t)nelo, "nsb_fld_read_into: %s truncated", path);
    if (!lglel) {
      NSB_REQUIRE(nel_local <= nelo, "nsb_fld_read_into: %s holds %lld elements, %lld wanted", path, nelo, (long long)nel_local);
      for (int64_t e = 0; e < nel_local; ++e) pos_of_local[e] = e;
    } else {
      std::vector<int64_t> where((size_t)nelg + 1, -1);
      for (long long q = 0; q < nelo; ++q) {
        uint32_t u;
        memcpy(&u, em + 4 * q, 4);
        if (swap) u = bswap32(u);
        if ((long long)u >= 1 && (long long)u <= nelg) where[u] = q;
      }
      for (int64_t e = 0; e < nel_local; ++e) {
        NSB_REQUIRE(lglel[e] >= 1 && lglel[e] <= nelg && where[lglel[e]] >= 0,
                    "nsb_fld_read_into: global element %lld is not in %s", (long long)lglel[e], path);
        pos_of_local[e] = where[lglel[e]];
      }
    }
  }
  size_t off = 136 + 4 * (size_t)nelo;
  auto value = [&](size_t byte_off) -> double {
    if (wd == 8) {
      uint64_t u;
      memcpy(&u, raw.data() + byte_off, 8);
      if (swap) u = bswap64(u);
      double d;
      memcpy(&d, &u, 8);
      return d;
    }
    uint32_t u;
    memcpy(&u, raw.data() + byte_off, 4);
    if (swap) u = bswap32(u);
    float fl;
    memcpy(&fl, &u, 4);
    return (double)fl;
  };
  std::vector<std::vector<double>> host(L->nfields);
  std::vector<const double *> ptrs(L->nfields, nullptr);
  auto take_group = [&](int ncomp, const int *dest_fields) -> int {
    const size_t bytes = (size_t)nelo * ncomp * npt * wd;
    NSB_REQUIRE(raw.size() >= off + bytes, "nsb_fld_read_into: %s truncated", path);
    for (int c = 0; c < ncomp; ++c) {
      const int f = dest_fields ? dest_fields[c] : -1;
      if (f < 0) continue;
      NSB_REQUIRE(f < L->nfields && L->hlen[f] == nel_local * npt,
                  "nsb_fld_read_into: layout field %d has %lld entries, file gives %lld", f,
                  (long long)(f < L->nfields ? L->hlen[f] : -1), (long long)(nel_local * npt));
      host[f].resize((size_t)(nel_local * npt));
      for (int64_t e = 0; e < nel_local; ++e) {
        const size_t base = off + ((size_t)pos_of_local[e] * ncomp + c) * npt * wd;
        double *dst = host[f].data() + e * npt;
        for (int64_t q = 0; q < npt; ++q) dst[q] = value(base + (size_t)q * wd);
      }
      ptrs[f] = host[f].data();
    }
    off += bytes;
    return NSB_OK;
  };
  for (const char *c = rd; *c; ++c) {
    if (*c == 'X') {
      NSB_CHECK(take_group(ndim, nullptr));
    } else if (*c == 'U') {
      int dst[3] = {ufield0 >= 0 ? ufield0 : -1, ufield0 >= 0 ? ufield0 + 1 : -1, ufield0 >= 0 ? ufield0 + 2 : -1};
      NSB_CHECK(take_group(ndim, dst));
    } else if (*c == 'P') {
      NSB_CHECK(take_group(1, &pfield));
    } else if (*c == 'T') {
      NSB_CHECK(take_group(1, &tfield));
    }
  }
  if (time_out) *time_out = time;
  // load_files (core/IO.f90:60-68) copies vx, vy, vz, pr, t only: %time of the loaded vector stays zero
  return nsb_vec_upload(B, col, ptrs.data(), 0.0);
}

// Restart of krylov_schur (core/eigensolvers.f90:240-285): H from HES<session><mstart>, Krylov vectors
// 1..mstart+1 from KRY<session>0.f00001 ... into columns 0..mstart of Q.  On return *mstart_next is the 0-based
// index of the next Arnoldi step (the reference's `mstart = mstart + 1`, 1-based).  dir may be NULL (cwd).
extern "C" int nsb_restart_load(nsb_basis_t Q, const char *dir, const char *session, int mstart, int k_dim,
                                const int64_t *lglel, int64_t nel_local, int ufield0, int pfield, int tfield, double *H,
                                int ldh, int *mstart_next) {
  NSB_REQUIRE(Q && session && H && mstart >= 1 && k_dim >= 1, "nsb_restart_load: bad argument");
  NSB_REQUIRE(mstart + 1 <= Q->ncols, "nsb_restart_load: %d vectors to load, basis has %d columns", mstart + 1, Q->ncols);
  const std::string base = (dir && dir[0]) ? std::string(dir) + "/" : std::string();
  char name[64];
  snprintf(name, sizeof name, "HES%s%04d", session, mstart);
  NSB_CHECK(nsb_hessenberg_read((base + name).c_str(), k_dim, mstart, H, ldh));
  for (int i = 1; i <= mstart + 1; ++i) {
    snprintf(name, sizeof name, "KRY%s0.f%05d", session, i);
    NSB_CHECK(nsb_fld_read_into(Q, i - 1, (base + name).c_str(), lglel, nel_local, ufield0, pfield, tfield, nullptr));
  }
  if (mstart_next) *mstart_next = mstart;
  return NSB_OK;
}
