/*
 * pprts_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see pprts_oracle.h for the contract and pinning status).
 *
 * fp64 restatement of the reference's PETSc-free pprts path.  Array layouts are the reference's own
 * (dof fastest, then z, x, y; Fortran column-major) so that every loop can be read next to the Fortran it follows.
 */
#define _GNU_SOURCE
#include "pprts_oracle.h"

#include <fcntl.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

/* ================================================================================================================
 * search / interpolation
 * ================================================================================================================ */

/* src/search.fypp:177-232 (search_sorted_bisection): 1-based "real-numbered" location of val in a sorted array. */
double orc_search_sorted_bisection(const double *arr, int n, double val) {
  int i = 1, j = n, k;
  double inc;
#define A_(x) arr[(x)-1]
  if (A_(i) <= A_(j)) { /* ascending */
    for (;;) {
      k = (i + j) / 2;
      if (val < A_(k)) j = k; else i = k;
      if (i + 1 >= j) {
        inc = (i == j) ? 0.0 : (val - A_(i)) / (A_(j) - A_(i));
        return dmin(dmax(1.0, (double)i + inc), (double)n);
      }
    }
  } else { /* descending */
    for (;;) {
      k = (i + j) / 2;
      if (val > A_(k)) j = k; else i = k;
      if (i + 1 >= j) {
        inc = (i == j) ? 0.0 : (val - A_(j)) / (A_(i) - A_(j));
        return dmin(dmax(1.0, (double)j - inc), (double)n);
      }
    }
  }
#undef A_
}

/* Same algorithm on an fp32 axis with an fp32 sample value (find_real_location as called from
 * src/optprop_LUT.F90:1470-1472), arithmetic carried out in fp64. */
double orc_find_real_location_f32axis(const float *arr, int n, float val) {
  int i = 1, j = n, k;
  double inc, v = (double)val;
#define A_(x) ((double)arr[(x)-1])
  if (A_(i) <= A_(j)) {
    for (;;) {
      k = (i + j) / 2;
      if (v < A_(k)) j = k; else i = k;
      if (i + 1 >= j) {
        inc = (i == j) ? 0.0 : (v - A_(i)) / (A_(j) - A_(i));
        return dmin(dmax(1.0, (double)i + inc), (double)n);
      }
    }
  } else {
    for (;;) {
      k = (i + j) / 2;
      if (v > A_(k)) j = k; else i = k;
      if (i + 1 >= j) {
        inc = (i == j) ? 0.0 : (v - A_(j)) / (A_(i) - A_(j));
        return dmin(dmax(1.0, (double)j - inc), (double)n);
      }
    }
  }
#undef A_
}

/* src/interpolation.F90:63 and :546-556.  interpolation_lattice_snapping = max(1e-3_irealLUT, epsilon) */
static const double ORC_SNAP = (double)1e-3f;
int orc_dim_needs_interpolation(double pti) {
  double frac = pti - (double)(long)pti;
  if (frac < ORC_SNAP) return 0;
  if (frac > 1.0 - ORC_SNAP) return 0;
  return 1;
}

/* src/interpolation.F90:317-360 (interp_vec_bilinear_iterative). pti are 1-based fractional indices,
 * db is [lattice][coef] (Fortran db(:, ofs)), offsets as ndarray_offsets. */
void orc_interp_vec_nd(const double *pti, int ndim, const float *db, int ncoef, const long *offsets, double *out) {
  long ioff_lo[ORC_MAX_DIM], ioff_hi[ORC_MAX_DIM], ofs_base = 1, ofs;
  double wlo[ORC_MAX_DIM], whi[ORC_MAX_DIM], w;
  int Ninterp = 0, d, c;
  long b;
  for (d = 0; d < ndim; ++d) {
    if (orc_dim_needs_interpolation(pti[d])) {
      b = (long)pti[d];
      whi[Ninterp] = pti[d] - (double)b;
      wlo[Ninterp] = 1.0 - whi[Ninterp];
      ioff_lo[Ninterp] = offsets[d] * (b - 1);
      ioff_hi[Ninterp] = offsets[d] * b;
      ++Ninterp;
    } else {
      ofs_base += offsets[d] * (lround(pti[d]) - 1);
    }
  }
  for (c = 0; c < ncoef; ++c) out[c] = 0.0;
  for (b = 0; b < (1L << Ninterp); ++b) {
    ofs = ofs_base;
    w = 1.0;
    for (d = 0; d < Ninterp; ++d) {
      if (b & (1L << d)) { ofs += ioff_hi[d]; w = w * whi[d]; }
      else               { ofs += ioff_lo[d]; w = w * wlo[d]; }
    }
    const float *col = db + (ofs - 1) * (long)ncoef;
    for (c = 0; c < ncoef; ++c) out[c] = fma(w, (double)col[c], out[c]);
  }
}

/* ================================================================================================================
 * closed-form per-cell scalars
 * ================================================================================================================ */

static int approx_d(double a, double b) { /* helper_functions.fypp:1272-1287, default precision 10*epsilon */
  double f = 10.0 * DBL_EPSILON;
  return (a <= b + f) && (a >= b - f);
}

/* src/eddington.F90:167-242 (eddington_coeff_ec); irealeddington = real64 (src/data_parameters.F90:100) */
void orc_eddington_coeff_ec(double dtau, double w0, double g, double mu0, double *t, double *r, double *rdir,
                            double *sdir, double *tdir) {
  double f, g1, g2, g3, g4, alpha1, alpha2, A, beta, e0, e, e2, k_mu0, k_g3, k_g4, k_2_e, dtau_slant;
  f = 0.75 * g;
  g1 = 2.0 - w0 * (1.25 + f);
  g2 = w0 * (0.75 - f);
  g3 = 0.5 - mu0 * f;
  dtau_slant = dmax(dtau / dmax(sqrt(DBL_MIN), mu0), 0.0);
  if (dtau_slant > 1e-6) {
    g4 = 1.0 - g3;
    alpha1 = g1 * g4 + g2 * g3;
    alpha2 = g1 * g3 + g2 * g4;
    A = sqrt(dmax((g1 - g2) * (g1 + g2), 1e-12));
    k_mu0 = A * mu0;
    k_g3 = A * g3;
    k_g4 = A * g4;
    e0 = exp(-dtau_slant);
    *tdir = e0;
    e = exp(-A * dtau);
    e2 = e * e;
    k_2_e = 2 * A * e;
    if (approx_d(k_mu0, 1.0)) k_mu0 = 1 - 10 * DBL_EPSILON;
    beta = 1 / (A + g1 + (A - g1) * e2);
    *r = g2 * (1 - e2) * beta;
    *t = k_2_e * beta;
    beta = w0 * beta / (1 - k_mu0 * k_mu0);
    *sdir = beta * (k_2_e * (g4 + alpha1 * mu0) - e0 * ((1 + k_mu0) * (alpha1 + k_g4) - (1 - k_mu0) * (alpha1 - k_g4) * e2));
    *rdir = beta * ((1 - k_mu0) * (alpha2 + k_g3) - (1 + k_mu0) * (alpha2 - k_g3) * e2 - k_2_e * (g3 - alpha2 * mu0) * e0);
  } else {
    *t = 1.0 - g1 * dtau;
    *r = g2 * dtau;
    *sdir = (1.0 - g3) * (w0 * dtau);
    *rdir = g3 * (w0 * dtau);
    *tdir = 1.0 - dtau_slant;
  }
}

/* src/schwarzschild.F90:36-67 (B_eff): 2-point Gauss-Legendre on (0,1) (dgauss, :173-303, m = 2) */
static double B_eff_mu(double B_far, double B_near, double tau, double mu) {
  double dtau = tau / mu, tm1;
  if (dtau < 1e-3) return (B_far + B_near) * .5;
  tm1 = expm1(-dtau);
  return (-B_near + B_far * (tm1 + 1)) / tm1 + ((B_far - B_near) * mu) / tau;
}
double orc_B_eff(double B_far, double B_near, double tau) {
  static const double pt[2] = {0.21132486540518713, 0.78867513459481287};
  static const double wi[2] = {0.5, 0.5};
  double B = 0;
  int imu;
  for (imu = 0; imu < 2; ++imu) B = B + B_eff_mu(B_far, B_near, tau, pt[imu]) * pt[imu] * wi[imu];
  return B * 2;
}

/* src/helper_functions.fypp:1622-1666 (delta_scale with f = g**2, delta_scale_optprop) */
void orc_delta_scale(double *kabs, double *ksca, double *g) {
  double f = (*g) * (*g), dtau = *kabs + *ksca, w0;
  if (dtau < DBL_EPSILON) return;
  w0 = *ksca / dtau;
  if (*g >= 1.0 - DBL_EPSILON * 10) {
    dtau = dtau * (1.0 - w0);
    w0 = 0;
    *g = 0;
  } else {
    dtau = dtau * (1.0 - w0 * f);
    *g = (*g - f) / (1.0 - f);
    w0 = w0 * (1.0 - f) / (1.0 - f * w0);
  }
  *kabs = dtau * (1.0 - w0);
  *ksca = dtau * w0;
}

/* src/schwarzschild.F90:69-80 (schwarzschild_radiance); the threshold is the default-real literal 1e-3 */
static void schwarzschild_radiance(double tau, double B_near, double B_far, double *L) {
  if (tau > (double)1e-3f) {
    double tm1 = expm1(-tau);
    *L = *L * (tm1 + 1) + (B_far - B_near) - (B_near - (B_far - B_near) / tau) * tm1;
  } else {
    *L = (B_near + B_far) * .5 * tau + *L * (1.0 - tau);
  }
}
/* src/schwarzschild.F90:82-150 (schwarzschild, use_legendre branch) with Nmu = 2: Edn, Eup have ke+1 entries */
void orc_schwarzschild2(int ke, const double *dtau, double albedo, double *Edn, double *Eup, const double *planck,
                        const double *opt_srfc_emission) {
  static const double pt[2] = {0.21132486540518713, 0.78867513459481287};
  static const double wi[2] = {0.5, 0.5};
  int imu, k;
  double Bsrfc = opt_srfc_emission ? *opt_srfc_emission : planck[ke];
  for (k = 0; k <= ke; ++k) { Edn[k] = 0; Eup[k] = 0; }
  for (imu = 0; imu < 2; ++imu) {
    double mu = pt[imu], Ldn = 0;
    for (k = 0; k < ke; ++k) {
      schwarzschild_radiance(dtau[k] / mu, planck[k], planck[k + 1], &Ldn);
      Edn[k + 1] = Edn[k + 1] + Ldn * mu * wi[imu];
    }
  }
  for (imu = 0; imu < 2; ++imu) {
    double mu = pt[imu], Lup = Bsrfc * (1.0 - albedo) + albedo * Edn[ke] * 2;
    Eup[ke] = Eup[ke] + Lup * mu * wi[imu];
    for (k = ke - 1; k >= 0; --k) {
      schwarzschild_radiance(dtau[k] / mu, planck[k + 1], planck[k], &Lup);
      Eup[k] = Eup[k] + Lup * mu * wi[imu];
    }
  }
  for (k = 0; k <= ke; ++k) { Eup[k] = Eup[k] * 2 * ORC_PI; Edn[k] = Edn[k] * 2 * ORC_PI; }
}
/* src/pprts.F90:2125-2198 (adding) */
void orc_adding(int N, double *a11, double *a12, double *a21, double *a22, double *a13, double *a23, double *a33,
                const double *dtau, const double *planck, double *Btop, double *Bbot) {
  double t = a11[0], r = a12[0], tdir = a33[0], rdir = a13[0], sdir = a23[0];
  double rl, tl, Tbot, Ttop, Rbot, Rtop;
  int k;
  for (k = 1; k < N; ++k) { /* reflectivity as seen from the top */
    rl = r; tl = t;
    r = r + (a12[k] * t * t) / (1.0 - r * a12[k]);
    t = t * a11[k] / (1.0 - rl * a12[k]);
    sdir = (a11[k] * sdir + tdir * a13[k] * rl * a11[k]) / (1.0 - rl * a12[k]) + tdir * a23[k];
    rdir = rdir + (tdir * a13[k] + sdir * a12[k]) * tl;
    tdir = tdir * a33[k];
  }
  Ttop = t; Rtop = r;
  a13[N - 1] = rdir; a23[N - 1] = sdir; a33[N - 1] = tdir;
  t = a22[N - 1]; r = a21[N - 1];
  for (k = N - 2; k >= 0; --k) { /* reflectivity as seen from the bottom */
    rl = r; tl = t;
    r = a12[k] + (r * a11[k] * a11[k]) / (1.0 - r * a12[k]);
    t = t * a11[k] / (1.0 - rl * a21[k]);
  }
  (void)tl;
  Tbot = t; Rbot = r;
  a12[N - 1] = Rtop; a22[N - 1] = Ttop; a11[N - 1] = Tbot; a21[N - 1] = Rbot;
  for (k = 0; k < N - 1; ++k) a11[k] = a12[k] = a21[k] = a22[k] = a13[k] = a23[k] = a33[k] = 0.0;
  if (planck) {
    double *Edn = (double *)malloc(sizeof(double) * (N + 1)), *Eup = (double *)malloc(sizeof(double) * (N + 1));
    const double zero = 0.0;
    orc_schwarzschild2(N, dtau, 0.0, Edn, Eup, planck, &zero);
    *Bbot = Edn[N] / ORC_PI;
    *Btop = Eup[0] / ORC_PI;
    free(Edn); free(Eup);
  }
}

/* ================================================================================================================
 * stream tables and symmetry permutations
 * ================================================================================================================ */

/* src/pprts.F90:332-396 */
int orc_streams_init(orc_streams *st, int solver_id) {
  int i;
  memset(st, 0, sizeof(*st));
  st->solver_id = solver_id;
  st->dirtop_adiv = 1;
  st->dirside_adiv = 1;
  switch (solver_id) {
  case 310:
    st->Ndt = 1; st->Nds = 1; st->Nft = 2; st->Nfs = 4;
    st->ft_inward[0] = 0; st->ft_inward[1] = 1;
    for (i = 0; i < 4; ++i) st->fs_inward[i] = i % 2; /* [F,T,F,T] */
    break;
  case 810:
    st->Ndt = 4; st->Nds = 2; st->Nft = 2; st->Nfs = 4;
    st->dirtop_adiv = 4; st->dirside_adiv = 2;
    st->ft_inward[0] = 0; st->ft_inward[1] = 1;
    for (i = 0; i < 4; ++i) st->fs_inward[i] = i % 2;
    break;
  case 324:
    st->Ndt = 1; st->Nds = 1; st->Nft = 8; st->Nfs = 8;
    for (i = 0; i < 8; ++i) st->ft_inward[i] = i % 2;       /* [F,T]x4 */
    for (i = 0; i < 8; ++i) st->fs_inward[i] = (i + 1) % 2; /* [T,F]x4 */
    break;
  default:
    return 1;
  }
  st->Nd = st->Ndt + 2 * st->Nds;
  st->Nf = st->Nft + 2 * st->Nfs;
  return 0;
}

/* dir2diff permutations: src/optprop.F90:1009-1045 (3->10), :1098-1124 (3->24), :1148-1184 (8->10).
 * c is coeff(dst*Nd + src), 0-based; the Fortran block index is dst+1. */
static void copy_block(double *c, const double *bak, int Nd, int dst_new, int dst_old, const int *srcperm) {
  int s;
  for (s = 0; s < Nd; ++s) c[dst_new * Nd + s] = bak[dst_old * Nd + (srcperm ? srcperm[s] : s)];
}
void orc_dir2diff_symmetry(const orc_streams *st, double *c, int le, int ln) {
  double bak[ORC_MAX_DOF * ORC_MAX_DOF];
  int Nd = st->Nd, n = st->Nd * st->Nf, k;
  if (st->solver_id == 310) {
    if (le) { memcpy(bak, c, n * sizeof(double)); copy_block(c, bak, Nd, 2, 3, 0); copy_block(c, bak, Nd, 3, 2, 0);
              copy_block(c, bak, Nd, 4, 5, 0); copy_block(c, bak, Nd, 5, 4, 0); }
    if (ln) { memcpy(bak, c, n * sizeof(double)); copy_block(c, bak, Nd, 6, 7, 0); copy_block(c, bak, Nd, 7, 6, 0);
              copy_block(c, bak, Nd, 8, 9, 0); copy_block(c, bak, Nd, 9, 8, 0); }
  } else if (st->solver_id == 810) {
    static const int pe[8] = {1, 0, 3, 2, 4, 5, 6, 7}, pn[8] = {2, 3, 0, 1, 4, 5, 6, 7};
    if (le) { memcpy(bak, c, n * sizeof(double)); copy_block(c, bak, Nd, 2, 3, pe); copy_block(c, bak, Nd, 3, 2, pe);
              copy_block(c, bak, Nd, 4, 5, pe); copy_block(c, bak, Nd, 5, 4, pe); }
    if (ln) { memcpy(bak, c, n * sizeof(double)); copy_block(c, bak, Nd, 6, 7, pn); copy_block(c, bak, Nd, 7, 6, pn);
              copy_block(c, bak, Nd, 8, 9, pn); copy_block(c, bak, Nd, 9, 8, pn); }
  } else if (st->solver_id == 324) {
    /* pC(src, dst), 1-based dst lists from the Fortran, written 0-based here */
    static const int e_new[8] = {0, 2, 4, 6, 1, 3, 5, 7}, e_old[8] = {2, 0, 6, 4, 3, 1, 7, 5};
    static const int n_new[8] = {0, 2, 4, 6, 1, 3, 5, 7}, n_old[8] = {4, 6, 0, 2, 5, 7, 1, 3};
    if (le) {
      memcpy(bak, c, n * sizeof(double));
      for (k = 0; k < 8; ++k) copy_block(c, bak, Nd, e_new[k], e_old[k], 0);
      for (k = 0; k < 4; ++k) { copy_block(c, bak, Nd, 9 + 2 * k, 8 + 2 * k, 0); copy_block(c, bak, Nd, 8 + 2 * k, 9 + 2 * k, 0); }
    }
    if (ln) {
      memcpy(bak, c, n * sizeof(double));
      for (k = 0; k < 8; ++k) copy_block(c, bak, Nd, n_new[k], n_old[k], 0);
      for (k = 0; k < 4; ++k) { copy_block(c, bak, Nd, 17 + 2 * k, 16 + 2 * k, 0); copy_block(c, bak, Nd, 16 + 2 * k, 17 + 2 * k, 0); }
    }
  }
}
/* dir2dir: none for 3 streams (src/optprop.F90:1256-1266), 8 streams :1268-1301 */
void orc_dir2dir_symmetry(const orc_streams *st, double *c, int le, int ln) {
  double bak[ORC_MAX_DOF * ORC_MAX_DOF];
  int Nd = st->Nd, n = Nd * Nd, k;
  if (Nd != 8) return;
  static const int pe[8] = {1, 0, 3, 2, 4, 5, 6, 7}, pn[8] = {2, 3, 0, 1, 4, 5, 6, 7};
  if (le) { memcpy(bak, c, n * sizeof(double)); for (k = 0; k < 8; ++k) copy_block(c, bak, Nd, k, pe[k], pe); }
  if (ln) { memcpy(bak, c, n * sizeof(double)); for (k = 0; k < 8; ++k) copy_block(c, bak, Nd, k, pn[k], pn); }
}

/* ================================================================================================================
 * LUT files and lookup
 * ================================================================================================================ */

/* src/mmap.F90:129-203 (binary_file_to_mmap / load_mmap_array_2d) */
int orc_mmap4_open(const char *path, const float **data, long *ncoef, long *nlat, void **map, unsigned long *maplen) {
  long pagesize = sysconf(_SC_PAGESIZE);
  size_t nhdr = (size_t)pagesize / sizeof(size_t), i, ndim = 0;
  size_t *header = (size_t *)malloc(nhdr * sizeof(size_t));
  int fd = open(path, O_RDONLY);
  if (fd < 0) { free(header); return 1; }
  if (read(fd, header, nhdr * sizeof(size_t)) != (ssize_t)(nhdr * sizeof(size_t))) { close(fd); free(header); return 2; }
  if (header[0] != sizeof(float)) { close(fd); free(header); return 3; }
  for (i = 3; i < nhdr; ++i) if (header[i] != 0) ++ndim;
  if (ndim != 2) { close(fd); free(header); return 4; }
  *ncoef = (long)header[3];
  *nlat = (long)header[4];
  *maplen = header[2] + (unsigned long)pagesize;
  *map = mmap(0, *maplen, PROT_READ, MAP_PRIVATE | MAP_NORESERVE, fd, 0);
  close(fd);
  free(header);
  if (*map == MAP_FAILED) return 5;
  *data = (const float *)((const char *)(*map) + pagesize);
  return 0;
}

/* src/mmap.F90:63-127 (arr_to_binary_datafile_2d) */
int orc_mmap4_write(const char *path, const float *data, long ncoef, long nlat) {
  long pagesize = sysconf(_SC_PAGESIZE);
  size_t nhdr = (size_t)pagesize / sizeof(size_t);
  size_t *header = (size_t *)calloc(nhdr, sizeof(size_t));
  FILE *f = fopen(path, "wb");
  if (!f) { free(header); return 1; }
  header[0] = sizeof(float);
  header[1] = (size_t)(ncoef * nlat);
  header[2] = (size_t)(ncoef * nlat) * sizeof(float);
  header[3] = (size_t)ncoef;
  header[4] = (size_t)nlat;
  fwrite(header, sizeof(size_t), nhdr, f);
  fwrite(data, sizeof(float), (size_t)(ncoef * nlat), f);
  fclose(f);
  free(header);
  return 0;
}

static void axes_offsets(orc_lut_axes *a) { /* helper_functions.fypp:2432-2437 ndarray_offsets */
  int d;
  a->off[0] = 1;
  for (d = 1; d < a->ndim; ++d) a->off[d] = a->off[d - 1] * a->n[d - 1];
}

int orc_lut_load(orc_lut *lut, int solver_id, const char *pT, const char *pS, const char *pD, const int *ndir,
                 const float *const *axdir) {
  orc_streams st;
  long nc, nl, nlat_dir = 1, nlat_diff = 1;
  int d, rc;
  memset(lut, 0, sizeof(*lut));
  if (orc_streams_init(&st, solver_id)) return 10;
  lut->Nd = st.Nd; lut->Nf = st.Nf;
  lut->dir.ndim = 6; lut->diff.ndim = 4;
  for (d = 0; d < 6; ++d) {
    float *cp = (float *)malloc(sizeof(float) * ndir[d]);
    memcpy(cp, axdir[d], sizeof(float) * ndir[d]);
    lut->dir.n[d] = ndir[d]; lut->dir.ax[d] = cp; nlat_dir *= ndir[d];
    if (d < 4) { lut->diff.n[d] = ndir[d]; lut->diff.ax[d] = cp; nlat_diff *= ndir[d]; }
  }
  axes_offsets(&lut->dir); axes_offsets(&lut->diff);
  if (pT) { if ((rc = orc_mmap4_open(pT, &lut->Tdir, &nc, &nl, &lut->map[0], &lut->maplen[0]))) return 20 + rc;
            if (nc != st.Nd * st.Nd || nl != nlat_dir) return 30; }
  if (pS) { if ((rc = orc_mmap4_open(pS, &lut->Sdir, &nc, &nl, &lut->map[1], &lut->maplen[1]))) return 40 + rc;
            if (nc != st.Nd * st.Nf || nl != nlat_dir) return 50; }
  if (pD) { if ((rc = orc_mmap4_open(pD, &lut->Sdiff, &nc, &nl, &lut->map[2], &lut->maplen[2]))) return 60 + rc;
            if (nc != st.Nf * st.Nf || nl != nlat_diff) return 70; }
  return 0;
}

void orc_lut_free(orc_lut *lut) {
  int d;
  for (d = 0; d < 3; ++d) if (lut->map[d]) munmap(lut->map[d], lut->maplen[d]);
  for (d = 0; d < 6; ++d) free((void *)lut->dir.ax[d]);
  memset(lut, 0, sizeof(*lut));
}

/* src/optprop.F90:550-593 (get_coeff_cube) + src/optprop_LUT.F90:1452-1613 (LUT_get_*) */
int orc_opp_get_coeff(const orc_lut *lut, const orc_streams *st, float tauz, float w0, float g, float aspect_zx,
                      float phi, float theta, int mode, int le, int ln, double *coeff) {
  double pti[ORC_MAX_DIM];
  float sample[ORC_MAX_DIM];
  const orc_lut_axes *ax = (mode == 3) ? &lut->diff : &lut->dir;
  int d;
  float save_aspect = aspect_zx;
  if (aspect_zx < ax->ax[2][0]) save_aspect = ax->ax[2][0]; /* OPP_TINYASPECT_RETCODE: clamp, continue */
  sample[0] = tauz; sample[1] = w0; sample[2] = save_aspect; sample[3] = g; sample[4] = phi; sample[5] = theta;
  for (d = 0; d < ax->ndim; ++d) pti[d] = orc_find_real_location_f32axis(ax->ax[d], ax->n[d], sample[d]);
  if (mode == 1) {
    if (!lut->Tdir) return 1;
    orc_interp_vec_nd(pti, 6, lut->Tdir, st->Nd * st->Nd, ax->off, coeff);
    orc_dir2dir_symmetry(st, coeff, le, ln);
  } else if (mode == 2) {
    if (!lut->Sdir) return 1;
    orc_interp_vec_nd(pti, 6, lut->Sdir, st->Nd * st->Nf, ax->off, coeff);
    orc_dir2diff_symmetry(st, coeff, le, ln);
  } else if (mode == 3) {
    if (!lut->Sdiff) return 1;
    orc_interp_vec_nd(pti, 4, lut->Sdiff, st->Nf * st->Nf, ax->off, coeff);
  } else {
    return 2;
  }
  return 0;
}

/* src/pprts_base.F90:1491-1543 (get_coeff): optical properties -> LUT space (fp32 casts, clamps) */
int orc_get_coeff(const orc_lut *lut, const orc_streams *st, double kabs, double ksca, double g, double dz,
                  double dx, int mode, double phi, double theta, int le, int ln, double *coeff) {
  const orc_lut_axes *ax = (mode == 3) ? &lut->diff : &lut->dir;
  float aspect_zx = (float)(dz / dx);
  float w0 = (float)(ksca / dmax(kabs + ksca, DBL_EPSILON));
  float tauz = (float)((kabs + ksca) * dz);
  aspect_zx = fmaxf(ax->ax[2][0], aspect_zx);
  tauz = fmaxf(ax->ax[0][0], fminf(ax->ax[0][ax->n[0] - 1], tauz));
  w0 = fmaxf(ax->ax[1][0], fminf(ax->ax[1][ax->n[1] - 1], w0));
  return orc_opp_get_coeff(lut, st, tauz, w0, (float)g, aspect_zx, (float)phi, (float)theta, mode, le, ln, coeff);
}

/* ================================================================================================================
 * solver
 * ================================================================================================================ */

#define ORC_UID_MIN (-1000)
#define ORC_UID_MAX (1000)
#define ORC_NHIST 100
#define ORC_NTIME 30

typedef struct {
  int lset, lsolar_rad, lthermal_rad, lchanged, lWm2_dir, lWm2_diff;
  double *edir, *ediff, *abso; /* interior (dof, Nz+1, Nx, Ny) / (Nz,Nx,Ny) */
  double dir_hist[ORC_NHIST], diff_hist[ORC_NHIST];
  double time[ORC_NTIME], maxnorm[ORC_NTIME]; /* error-growth tracking, src/pprts_base.F90:161-162 */
  int Niter_dir, Niter_diff;
  double final_res_dir, final_res_diff;
  double diff_sor_omega;
} orc_solution;

struct orc_solver {
  orc_config cfg;
  orc_streams st;
  const orc_lut *lut;
  int Nz, Nx, Ny, Nz1;
  double dx, dy;
  double *dz;
  int *l1d;
  double unconstrained_fraction;
  /* atmosphere collapse (src/pprts_base.F90:118-122, :1088-1093 atmk): the solver grid starts at atmosphere layer
   * icollapse-1; dz_atm / l1d_atm live on the Nz_atm atmosphere layers, everything else on the Nz solver layers */
  int Nz_atm, icollapse;
  double *dz_atm, *Btop, *Bbot;
  int *l1d_atm;
  /* t_suninfo, src/pprts_base.F90:128-135 */
  double phi, theta, mu, costheta, symmetry_phi, sundir[3];
  int xinc, yinc;
  /* t_atmosphere */
  double *kabs, *ksca, *g, *planck, *Bsrfc, *albedo;
  int have_planck, have_Bsrfc, have_optprop;
  double *a11, *a12, *a13, *a23, *a33;
  double *dir2dir, *dir2diff, *diff2diff;
  double *b;
  double *mm_xx, *mm_xb;          /* ghosted scratch of orc_mat_mult_ediff (kept between calls) */
  double *dir_scale, *diff_scale; /* Wm2_to_W, interior layout */
  /* opt_buildings (t_pprts_buildings, src/buildings.F90:40-69): faces of building cells */
  int nbld, bld_have_planck;
  int *bld_iface;
  double *bld_albedo, *bld_planck;
  orc_solution *sol[ORC_UID_MAX - ORC_UID_MIN + 1];
};

#define FG(a, nd, d, k, i, j) (a)[(d) + (long)(nd) * ((k) + (long)Nz1 * (((i) + 1) + (long)(Nx + 2) * ((j) + 1)))]
#define FI(a, nd, d, k, i, j) (a)[(d) + (long)(nd) * ((k) + (long)Nz1 * ((i) + (long)Nx * (j)))]
#define CF(a, nc, n, k, i, j) (a)[(n) + (long)(nc) * ((k) + (long)Nz * ((i) + (long)Nx * (j)))]
#define A3(a, k, i, j) (a)[(k) + (long)Nz * ((i) + (long)Nx * (j))]
#define P3(a, k, i, j) (a)[(k) + (long)Nz1 * ((i) + (long)Nx * (j))]
#define A2(a, i, j) (a)[(i) + (long)Nx * (j)]
#define LOCALS                                                                                                       \
  const int Nz = s->Nz, Nx = s->Nx, Ny = s->Ny, Nz1 = s->Nz1;                                                          \
  const int Nd = s->st.Nd, Nf = s->st.Nf, Ndt = s->st.Ndt, Nds = s->st.Nds, Nft = s->st.Nft, Nfs = s->st.Nfs;          \
  (void)Nz; (void)Nx; (void)Ny; (void)Nz1; (void)Nd; (void)Nf; (void)Ndt; (void)Nds; (void)Nft; (void)Nfs;

static long ngh(const orc_solver *s, int nd) { return (long)nd * s->Nz1 * (s->Nx + 2) * (s->Ny + 2); }
static long nin(const orc_solver *s, int nd) { return (long)nd * s->Nz1 * s->Nx * s->Ny; }

/* ---- halo helpers for ONE periodic rank (src/pprts_base.F90:1622-1731, src/pprts_explicit.F90:1019-1073) ------ */
static void interior_to_ghosted(const orc_solver *s, int nd, const double *in, double *gh) {
  LOCALS
  int d, k, i, j;
  memset(gh, 0, sizeof(double) * ngh(s, nd));
#pragma omp parallel for schedule(static) private(i, k, d)
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d)
    FG(gh, nd, d, k, i, j) = FI(in, nd, d, k, i, j);
}
static void ghosted_to_interior(const orc_solver *s, int nd, const double *gh, double *out) {
  LOCALS
  int d, k, i, j;
#pragma omp parallel for schedule(static) private(i, k, d)
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d)
    FI(out, nd, d, k, i, j) = FG(gh, nd, d, k, i, j);
}
/* halo_fill_5pt: x ghosts over the full (ghosted) y range, y ghosts over interior x only */
static void halo_fill_5pt(const orc_solver *s, int nd, double *v) {
  LOCALS
  int d, k, i, j;
  for (j = -1; j <= Ny; ++j) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    double se = FG(v, nd, d, k, Nx - 1, j), sw = FG(v, nd, d, k, 0, j);
    FG(v, nd, d, k, -1, j) = se; /* rw: from west neighbour's east edge */
    FG(v, nd, d, k, Nx, j) = sw; /* re */
  }
  for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    double sn = FG(v, nd, d, k, i, Ny - 1), ss = FG(v, nd, d, k, i, 0);
    FG(v, nd, d, k, i, -1) = sn;
    FG(v, nd, d, k, i, Ny) = ss;
  }
}
/* fill_ghost (src/pprts_explicit.F90:1019-1073): like halo_fill_5pt but x ghosts over interior y only */
static void fill_ghost(const orc_solver *s, int nd, double *v) {
  LOCALS
  int d, k, i, j;
  for (j = 0; j < Ny; ++j) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    double se = FG(v, nd, d, k, Nx - 1, j), sw = FG(v, nd, d, k, 0, j);
    FG(v, nd, d, k, -1, j) = se;
    FG(v, nd, d, k, Nx, j) = sw;
  }
  for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    double sn = FG(v, nd, d, k, i, Ny - 1), ss = FG(v, nd, d, k, i, 0);
    FG(v, nd, d, k, i, -1) = sn;
    FG(v, nd, d, k, i, Ny) = ss;
  }
}
/* halo_reduce_5pt: ghost -> owner (ADD), then zero ghosts */
static void halo_reduce_5pt(const orc_solver *s, int nd, double *v) {
  LOCALS
  int d, k, i, j;
  long n = (long)nd * Nz1 * (Ny + 2);
  double *se = (double *)malloc(sizeof(double) * n), *sw = (double *)malloc(sizeof(double) * n);
  long c = 0;
  for (j = -1; j <= Ny; ++j) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d, ++c) {
    se[c] = FG(v, nd, d, k, Nx, j);
    sw[c] = FG(v, nd, d, k, -1, j);
  }
  n = (long)nd * Nz1 * Nx;
  double *sn = (double *)malloc(sizeof(double) * n), *ss = (double *)malloc(sizeof(double) * n);
  c = 0;
  for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d, ++c) {
    sn[c] = FG(v, nd, d, k, i, Ny);
    ss[c] = FG(v, nd, d, k, i, -1);
  }
  c = 0;
  for (j = -1; j <= Ny; ++j) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d, ++c) {
    FG(v, nd, d, k, 0, j) += se[c];      /* rw = west neighbour's east ghost */
    FG(v, nd, d, k, Nx - 1, j) += sw[c]; /* re */
  }
  c = 0;
  for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d, ++c) {
    FG(v, nd, d, k, i, 0) += sn[c];      /* rs = south neighbour's north ghost */
    FG(v, nd, d, k, i, Ny - 1) += ss[c]; /* rn */
  }
  for (j = -1; j <= Ny; ++j) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    FG(v, nd, d, k, -1, j) = 0; FG(v, nd, d, k, Nx, j) = 0;
  }
  for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    FG(v, nd, d, k, i, -1) = 0; FG(v, nd, d, k, i, Ny) = 0;
  }
  free(se); free(sw); free(sn); free(ss);
}

/* ---- init ------------------------------------------------------------------------------------------------------ */

/* helper_functions.fypp:1993-2016 (cartesian_2_spherical) */
static void cartesian_2_spherical(const double sundir[3], double *phi, double *theta) {
  /* norm2() does not underflow for tiny components (tests/test_helper_functions: sundir = [0, -tiny, -1]): hypot */
  double n = hypot(hypot(sundir[0], sundir[1]), sundir[2]);
  double sn[3] = {sundir[0], sundir[1], sundir[2]}, proj[3], pn, az, c;
  if (n > 0) { sn[0] /= n; sn[1] /= n; sn[2] /= n; }
  c = dmax(-1.0, dmin(1.0, -sn[2])); /* dot(sundir_normed, -e_z) */
  *theta = acos(c) * 180.0 / ORC_PI;
  /* vec_proj_on_plane(sundir, -e_z) = v - dot(v,n) n */
  proj[0] = sundir[0]; proj[1] = sundir[1]; proj[2] = 0.0;
  pn = hypot(proj[0], proj[1]);
  if (pn > 0) { proj[0] /= pn; proj[1] /= pn; }
  c = dmax(-1.0, dmin(1.0, -proj[1])); /* angle to -e_y */
  az = acos(c);
  az = az * ((-proj[0]) >= 0 ? 1.0 : -1.0); /* sign(1, dot(proj, -e_x)) */
  *phi = az * 180.0 / ORC_PI;
}

/* src/pprts.F90:1118-1183 (setup_suninfo) */
void orc_set_angles(orc_solver *s, const double sundir[3]) {
  double d2r = ORC_PI / 180.0, sp;
  cartesian_2_spherical(sundir, &s->phi, &s->theta);
  s->mu = dmax(cos(s->theta * d2r), 0.0);
  /* spherical_2_cartesian, helper_functions.fypp:2022-2033 */
  s->sundir[0] = -sin(s->theta * d2r) * sin(s->phi * d2r);
  s->sundir[1] = -sin(s->theta * d2r) * cos(s->phi * d2r);
  s->sundir[2] = -cos(s->theta * d2r);
  if (s->theta >= 90.0) s->theta = -1.0;
  s->costheta = dmax(cos(s->theta * d2r), 0.0);
  sp = acos(cos(s->phi * d2r));
  s->symmetry_phi = dmin(90.0, dmax(0.0, asin(sin(sp)) / d2r));
  s->xinc = (sin(s->phi * d2r) > 0.0) ? 0 : 1;
  s->yinc = (cos(s->phi * d2r) < 0.0) ? 1 : 0;
}

void orc_get_suninfo(orc_solver *s, double *phi, double *theta, double *symmetry_phi, int *xinc, int *yinc) {
  *phi = s->phi; *theta = s->theta; *symmetry_phi = s->symmetry_phi; *xinc = s->xinc; *yinc = s->yinc;
}

/* src/pprts.F90:649-725 (determine_1d_layers), one rank; works on the atmosphere grid (C_one_atm) */
static void determine_1d_layers(orc_solver *s) {
  const int Nz = s->Nz_atm, Nx = s->Nx, Ny = s->Ny; /* A3 below indexes the atmosphere arrays */
  int k, i, j, any, kk, count3d = 0;
  int *l1d = s->l1d_atm;
  const double *dz = s->dz_atm;
  for (k = 0; k < Nz; ++k) l1d[k] = 0;
  any = 0;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) if (A3(dz, Nz - 1, i, j) / s->dx > s->cfg.twostr_ratio) any = 1;
  l1d[Nz - 1] = any;
  for (k = Nz - 2; k >= 0; --k) {
    any = 0;
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) if (A3(dz, k, i, j) / s->dx > s->cfg.twostr_ratio) any = 1;
    if (any) { for (kk = 0; kk <= k; ++kk) l1d[kk] = 1; break; }
  }
  if (s->icollapse > 1) /* :692-706 layers that are collapsed have to be 1-D */
    for (k = 0; k <= s->icollapse - 1; ++k) l1d[k] = 1;
  { /* :708-719: the 1-D block is contiguous from the top */
    int n1d = 0;
    for (k = 0; k < Nz; ++k) n1d += l1d[k];
    for (k = 0; k < n1d; ++k) l1d[k] = 1;
  }
  for (k = 0; k < Nz; ++k) count3d += !l1d[k];
  s->unconstrained_fraction = (double)count3d / (double)Nz;
  for (k = 0; k < s->Nz; ++k) s->l1d[k] = l1d[k + s->icollapse - 1]; /* atmk, src/pprts_base.F90:1088-1093 */
}

/* src/pprts.F90:3901-3987 (gen_scale_dir_flx_vec_arr / gen_scale_diff_flx_vec_arr) */
static void gen_scale_vectors(orc_solver *s) {
  LOCALS
  int k, i, j, d;
  double Az, Ax, Ay;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) {
    for (k = 0; k < Nz1; ++k) {
      Az = s->dx * s->dy / (double)s->st.dirtop_adiv;
      for (d = 0; d < Ndt; ++d) FI(s->dir_scale, Nd, d, k, i, j) = Az;
    }
    for (k = 0; k < Nz; ++k) {
      Ax = s->dy * A3(s->dz, k, i, j) / (double)s->st.dirside_adiv;
      for (d = 0; d < Nds; ++d) FI(s->dir_scale, Nd, Ndt + d, k, i, j) = Ax;
      Ay = s->dx * A3(s->dz, k, i, j) / (double)s->st.dirside_adiv;
      for (d = 0; d < Nds; ++d) FI(s->dir_scale, Nd, Ndt + Nds + d, k, i, j) = Ay;
    }
    for (d = Ndt; d < Nd; ++d) FI(s->dir_scale, Nd, d, Nz, i, j) = 1.0;
    /* diffuse: difftop%area_divider = diffside%area_divider = 1 for every pprts solver */
    Az = s->dx * s->dy;
    for (k = 0; k < Nz1; ++k) for (d = 0; d < Nft; ++d) FI(s->diff_scale, Nf, d, k, i, j) = Az;
    for (k = 0; k < Nz; ++k) {
      Ax = s->dy * A3(s->dz, k, i, j);
      for (d = 0; d < Nfs; ++d) FI(s->diff_scale, Nf, Nft + d, k, i, j) = Ax;
      Ay = s->dx * A3(s->dz, k, i, j);
      for (d = 0; d < Nfs; ++d) FI(s->diff_scale, Nf, Nft + Nfs + d, k, i, j) = Ay;
    }
    for (d = Nft; d < Nf; ++d) FI(s->diff_scale, Nf, d, Nz, i, j) = 1.0;
  }
}

orc_solver *orc_solver_create(const orc_config *cfg, const orc_lut *lut) {
  orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
  long nc;
  s->cfg = *cfg;
  if (orc_streams_init(&s->st, cfg->solver_id)) { free(s); return 0; }
  s->lut = lut;
  s->icollapse = cfg->collapseindex > 1 ? cfg->collapseindex : 1;
  if (s->icollapse > cfg->Nz) { free(s); return 0; }
  s->Nz_atm = cfg->Nz;
  s->Nz = cfg->Nz - s->icollapse + 1; s->Nx = cfg->Nx; s->Ny = cfg->Ny; s->Nz1 = s->Nz + 1;
  s->dx = cfg->dx; s->dy = cfg->dy;
  nc = (long)s->Nz_atm * s->Nx * s->Ny;
  s->dz_atm = (double *)malloc(sizeof(double) * nc);
  memcpy(s->dz_atm, cfg->dz, sizeof(double) * nc);
  nc = (long)s->Nz * s->Nx * s->Ny;
  s->dz = (double *)malloc(sizeof(double) * nc);
  { int k, i, j; for (j = 0; j < s->Ny; ++j) for (i = 0; i < s->Nx; ++i) for (k = 0; k < s->Nz; ++k)
      s->dz[k + (long)s->Nz * (i + (long)s->Nx * j)] = s->dz_atm[(k + s->icollapse - 1) + (long)s->Nz_atm * (i + (long)s->Nx * j)]; }
  s->l1d = (int *)calloc(s->Nz, sizeof(int));
  s->l1d_atm = (int *)calloc(s->Nz_atm, sizeof(int));
  s->Btop = (double *)calloc((long)s->Nx * s->Ny, sizeof(double));
  s->Bbot = (double *)calloc((long)s->Nx * s->Ny, sizeof(double));
  determine_1d_layers(s);
  s->kabs = (double *)calloc(nc, sizeof(double));
  s->ksca = (double *)calloc(nc, sizeof(double));
  s->g = (double *)calloc(nc, sizeof(double));
  s->planck = (double *)calloc((long)s->Nz1 * s->Nx * s->Ny, sizeof(double));
  s->Bsrfc = (double *)calloc((long)s->Nx * s->Ny, sizeof(double));
  s->albedo = (double *)calloc((long)s->Nx * s->Ny, sizeof(double));
  s->a11 = (double *)calloc(nc, sizeof(double)); s->a12 = (double *)calloc(nc, sizeof(double));
  s->a13 = (double *)calloc(nc, sizeof(double)); s->a23 = (double *)calloc(nc, sizeof(double));
  s->a33 = (double *)calloc(nc, sizeof(double));
  s->dir2dir = (double *)calloc(nc * s->st.Nd * s->st.Nd, sizeof(double));
  s->dir2diff = (double *)calloc(nc * s->st.Nd * s->st.Nf, sizeof(double));
  s->diff2diff = (double *)calloc(nc * s->st.Nf * s->st.Nf, sizeof(double));
  s->b = (double *)calloc(nin(s, s->st.Nf), sizeof(double));
  s->dir_scale = (double *)calloc(nin(s, s->st.Nd), sizeof(double));
  s->diff_scale = (double *)calloc(nin(s, s->st.Nf), sizeof(double));
  gen_scale_vectors(s);
  orc_set_angles(s, cfg->sundir);
#ifdef _OPENMP
  if (cfg->nthreads > 0) omp_set_num_threads(cfg->nthreads);
#endif
  return s;
}

static void free_solution(orc_solution *p) {
  if (!p) return;
  free(p->edir); free(p->ediff); free(p->abso); free(p);
}

void orc_solver_destroy(orc_solver *s) {
  int u;
  if (!s) return;
  for (u = 0; u <= ORC_UID_MAX - ORC_UID_MIN; ++u) free_solution(s->sol[u]);
  free(s->dz_atm); free(s->l1d_atm); free(s->Btop); free(s->Bbot);
  free(s->dz); free(s->l1d); free(s->kabs); free(s->ksca); free(s->g); free(s->planck); free(s->Bsrfc);
  free(s->albedo); free(s->a11); free(s->a12); free(s->a13); free(s->a23); free(s->a33);
  free(s->dir2dir); free(s->dir2diff); free(s->diff2diff); free(s->b); free(s->dir_scale); free(s->diff_scale);
  free(s->bld_iface); free(s->bld_albedo); free(s->bld_planck);
  free(s->mm_xx); free(s->mm_xb);
  free(s);
}

/* ---- set_optical_properties (src/pprts.F90:1764-2076) --------------------------------------------------------- */
int orc_set_optical_properties(orc_solver *s, double albedo, const double *kabs_in, const double *ksca_in,
                               const double *g_in, const double *planck_in, const double *planck_srfc,
                               const double *albedo_2d, int ldelta_scaling) {
  LOCALS
  const int Nza = s->Nz_atm, ic = s->icollapse, off = ic - 1; /* atmk(k) = k + icollapse - 1 */
  const long nca = (long)Nza * Nx * Ny;
  long n;
  int k, i, j;
  /* atmosphere-sized working copies (atm%kabs ... live on C_one_atm); the solver-grid arrays are the layers from
   * atmk(0) on.  Inputs that are not given keep their previous values (:1805-1860) -- only for icollapse == 1 */
  double *kabs = (double *)malloc(sizeof(double) * nca), *ksca = (double *)malloc(sizeof(double) * nca),
         *g = (double *)malloc(sizeof(double) * nca);
  int rc = 0;
#define ATM(a, k, i, j) (a)[(k) + (long)Nza * ((i) + (long)Nx * (j))]
#define PATM(a, k, i, j) (a)[(k) + (long)(Nza + 1) * ((i) + (long)Nx * (j))]
  for (n = 0; n < (long)Nx * Ny; ++n) s->albedo[n] = albedo_2d ? albedo_2d[n] : albedo;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nza; ++k) {
    const int ks = k - off; /* solver layer, < 0 inside the collapsed block */
    ATM(kabs, k, i, j) = kabs_in ? ATM(kabs_in, k, i, j) : (ks >= 0 ? A3(s->kabs, ks, i, j) : 0.0);
    ATM(ksca, k, i, j) = ksca_in ? ATM(ksca_in, k, i, j) : (ks >= 0 ? A3(s->ksca, ks, i, j) : 0.0);
    ATM(g, k, i, j) = g_in ? ATM(g_in, k, i, j) : (ks >= 0 ? A3(s->g, ks, i, j) : 0.0);
  }
  s->have_planck = planck_in != 0;
  if (planck_in)
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k)
      P3(s->planck, k, i, j) = PATM(planck_in, k + off, i, j);
  s->have_Bsrfc = planck_srfc != 0;
  if (planck_srfc) memcpy(s->Bsrfc, planck_srfc, sizeof(double) * (long)Nx * Ny);
  if (ldelta_scaling) /* :1912-1914 */
    for (n = 0; n < nca; ++n) orc_delta_scale(&kabs[n], &ksca[n], &g[n]);
  for (n = 0; n < nca; ++n) /* :1929-1935 */
    if (kabs[n] < 0 || ksca[n] < 0 || g[n] < 0 || isnan(kabs[n]) || isnan(ksca[n]) || isnan(g[n])) { rc = 1; goto done; }
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
    A3(s->kabs, k, i, j) = ATM(kabs, k + off, i, j);
    A3(s->ksca, k, i, j) = ATM(ksca, k + off, i, j);
    A3(s->g, k, i, j) = ATM(g, k + off, i, j);
  }
  /* :1962-2007 Eddington coefficients in 1-D layers; a21 = a12, a22 = a11 (:2072-2073) */
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
    if (s->l1d[k]) {
      double ext = dmax(DBL_MIN, A3(s->kabs, k, i, j) + A3(s->ksca, k, i, j));
      orc_eddington_coeff_ec(A3(s->dz, k, i, j) * ext, A3(s->ksca, k, i, j) / ext, A3(s->g, k, i, j), s->costheta,
                             &A3(s->a11, k, i, j), &A3(s->a12, k, i, j), &A3(s->a13, k, i, j),
                             &A3(s->a23, k, i, j), &A3(s->a33, k, i, j));
    }
  }
  if (ic > 1) { /* handle_atm_collapse :2080-2123: the top `ic` atmosphere layers become solver layer 0 */
    double *c[7], *dtau = (double *)malloc(sizeof(double) * ic), *pl = (double *)malloc(sizeof(double) * (ic + 1));
    int q;
    for (q = 0; q < 7; ++q) c[q] = (double *)malloc(sizeof(double) * ic);
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) {
      for (k = 0; k < ic; ++k) {
        double ext = dmax(DBL_MIN, ATM(kabs, k, i, j) + ATM(ksca, k, i, j));
        orc_eddington_coeff_ec(ATM(s->dz_atm, k, i, j) * ext, ATM(ksca, k, i, j) / ext, ATM(g, k, i, j), s->costheta,
                               &c[0][k], &c[1][k], &c[4][k], &c[5][k], &c[6][k]);
        c[2][k] = c[1][k]; c[3][k] = c[0][k]; /* a21 = a12, a22 = a11 */
        dtau[k] = ATM(s->dz_atm, k, i, j) * ATM(kabs, k, i, j);
        if (planck_in) pl[k] = PATM(planck_in, k, i, j);
      }
      if (planck_in) pl[ic] = PATM(planck_in, ic, i, j);
      orc_adding(ic, c[0], c[1], c[2], c[3], c[4], c[5], c[6], planck_in ? dtau : 0, planck_in ? pl : 0,
                 &A2(s->Btop, i, j), &A2(s->Bbot, i, j));
      A3(s->a11, 0, i, j) = c[0][ic - 1]; A3(s->a12, 0, i, j) = c[1][ic - 1];
      A3(s->a13, 0, i, j) = c[4][ic - 1]; A3(s->a23, 0, i, j) = c[5][ic - 1]; A3(s->a33, 0, i, j) = c[6][ic - 1];
    }
    for (q = 0; q < 7; ++q) free(c[q]);
    free(dtau); free(pl);
  }
  s->have_optprop = 1;
done:
  free(kabs); free(ksca); free(g);
#undef ATM
#undef PATM
  return rc;
}

/* ---- geometric direct-beam coefficients (src/geometric_coeffs.F90; -pprts_geometric_coeffs, src/pprts.F90:3142-3149)
 * dir2dir of a 3-stream box with arbitrary (sloped) top and bottom faces from pure geometry: every source face is
 * projected along the sun onto the opposite faces, the lit areas are weighted with Beer-Lambert along the beam.
 * vertices: a,b,c,d (upper four in the reference's numbering, z of the lower face first) ... h, 3 coordinates each. */
typedef double gv3[3];
static void g_sub(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static double g_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double g_norm(const double *a) { return sqrt(g_dot(a, a)); }
static void g_axpy(const double *o, double c, const double *d, double *out) { out[0] = o[0] + c * d[0]; out[1] = o[1] + c * d[1]; out[2] = o[2] + c * d[2]; }
static void g_copy(const double *a, double *o) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
/* compute_normal_3d, helper_functions.fypp:1960-1974 */
static void g_normal(const double *p1, const double *p2, const double *p3, double *n) {
  gv3 U, V; double l;
  g_sub(p2, p1, U); g_sub(p3, p1, V);
  n[0] = U[1] * V[2] - U[2] * V[1]; n[1] = U[2] * V[0] - U[0] * V[2]; n[2] = U[0] * V[1] - U[1] * V[0];
  l = g_norm(n); n[0] /= l; n[1] /= l; n[2] /= l;
}
/* hit_plane, intersection.fypp:79-91 */
static double g_hit_plane(const double *p_loc, const double *p_dir, const double *po, const double *pn) {
  double discr = g_dot(p_dir, pn); gv3 d;
  if (discr <= DBL_EPSILON && discr > -DBL_EPSILON) return sqrt(DBL_MAX);
  g_sub(po, p_loc, d);
  return g_dot(d, pn) / discr;
}
/* triangle_area_by_vertices / quadrangle_area_by_vertices, helper_functions.fypp:2116-2136, :3087-3092 (Heron) */
static double g_tri(const double *v1, const double *v2, const double *v3) {
  gv3 d; double e1, e2, e3, p2;
  g_sub(v2, v1, d); e1 = g_norm(d); g_sub(v3, v2, d); e2 = g_norm(d); g_sub(v1, v3, d); e3 = g_norm(d);
  p2 = (e1 + e2 + e3) / 2;
  return sqrt(dmax(p2 * (p2 - e1) * (p2 - e2) * (p2 - e3), 0.0));
}
static double g_quad(const double *v1, const double *v2, const double *v3, const double *v4) { return g_tri(v1, v2, v3) + g_tri(v1, v3, v4); }
/* line_intersection_3d, intersection.fypp:395-434 */
static int g_line_intersection(const double *p1, const double *dir1, const double *p3, const double *dir2, double *c1, double *c2) {
  gv3 a, b, p13; double na, nb, nd, den, d2121, d4343, d4321, d1343, d1321;
  const double limit = sqrt(DBL_EPSILON);
  g_copy(dir1, a); g_copy(dir2, b); /* p2 - p1 = direction1, p4 - p3 = direction2 */
  na = g_norm(a); nb = g_norm(b);
  nd = g_dot(a, b) / (na * nb);
  den = 1.0 * 1.0 - nd * nd; /* nd_mnop(p2,p1,p2,p1) = nd_mnop(p4,p3,p4,p3) = 1 */
  if (den <= limit && den >= -limit) { *c1 = -DBL_MAX; *c2 = -DBL_MAX; return 1; }
  g_sub(p1, p3, p13);
  d2121 = g_dot(a, a); d4343 = g_dot(b, b); d4321 = g_dot(b, a); d1343 = g_dot(p13, b); d1321 = g_dot(p13, a);
  den = d2121 * d4343 - d4321 * d4321;
  *c1 = (d1343 * d4321 - d1321 * d4343) / den;
  *c2 = (d1343 + *c1 * d4321) / d4343;
  return 0;
}
/* project_points :292-333 */
static void g_project_points(const double *sundir, const double *origin, const double *normal, double *v1, double *v2, double *v3, double *v4) {
  const double big = 1e5;
  double *vs[4]; int q, m;
  vs[0] = v1; vs[1] = v2; vs[2] = v3; vs[3] = v4;
  if (fabs(g_dot(sundir, normal)) < sqrt(DBL_EPSILON)) {
    gv3 sp;
    for (m = 0; m < 3; ++m) sp[m] = -(fabs(normal[m]) - 1.0) * sundir[m];
    for (q = 0; q < 4; ++q) {
      double h = g_hit_plane(vs[q], normal, origin, normal);
      for (m = 0; m < 3; ++m) vs[q][m] = vs[q][m] + h * normal[m] - big * sp[m];
    }
  } else {
    for (q = 0; q < 4; ++q) {
      double h = g_hit_plane(vs[q], sundir, origin, normal);
      for (m = 0; m < 3; ++m) vs[q][m] = vs[q][m] + h * sundir[m];
    }
  }
}
/* rearange_projection :394-404 */
static void g_rearange_projection(const double *direction1, const double *origin2, const double *direction2,
                                  const double *origin3, const double *direction3, double *origin1) {
  double c21, c22, c31, c32, c;
  g_line_intersection(origin1, direction1, origin2, direction2, &c21, &c22);
  g_line_intersection(origin1, direction1, origin3, direction3, &c31, &c32);
  c = dmin(dmax(dmax(c21, c31), 0.0), 1.0);
  g_axpy(origin1, c, direction1, origin1);
}
/* rearange_projections :372-392 */
static void g_rearange_projections(const double *f1, const double *f2, const double *f3, const double *f4,
                                   double *v1, double *v2, double *v3, double *v4) {
  gv3 d, d2, d3;
  g_sub(f1, v1, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f4, f3, d2); g_sub(f3, f2, d3); g_rearange_projection(d, f3, d2, f2, d3, v1); }
  g_sub(f2, v2, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f4, f3, d2); g_sub(f1, f4, d3); g_rearange_projection(d, f3, d2, f4, d3, v2); }
  g_sub(f3, v3, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f1, f2, d2); g_sub(f1, f4, d3); g_rearange_projection(d, f2, d2, f4, d3, v3); }
  g_sub(f4, v4, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f1, f2, d2); g_sub(f3, f2, d3); g_rearange_projection(d, f2, d2, f2, d3, v4); }
}
/* proj_var_to_edges :349-370 (the fourth point is moved along c3 - c2 although intersected along c4 - c1, as there) */
static void g_proj_var_to_edges(const double *c1, const double *c2, const double *c3, const double *c4, const double *v,
                                double *p1, double *p2, double *p3, double *p4) {
  gv3 d, e; double c, t;
  g_sub(c4, c1, d); g_sub(c4, c3, e); g_line_intersection(v, d, c3, e, &c, &t); g_axpy(v, c, d, p1);
  g_sub(c2, c1, d); g_sub(c3, c2, e); g_line_intersection(v, d, c2, e, &c, &t); g_axpy(v, c, d, p2);
  g_sub(c1, c2, d); g_sub(c1, c4, e); g_line_intersection(v, d, c4, e, &c, &t); g_axpy(v, c, d, p3);
  g_sub(c4, c1, d); g_sub(c1, c2, e); g_line_intersection(v, d, c2, e, &c, &t); g_sub(c3, c2, d); g_axpy(v, c, d, p4);
}
static double g_ext(double beta_s) { double x = dmax(beta_s, 10 * DBL_MIN); return expm1(-x) / (-x); } /* :276-283 */
static double g_num_dst(double s0, double l0, double h0, double c_ext, int n) { /* :248-274 */
  double dl, dh, ds, acc = 0.0; int i;
  if (n <= 0) n = 10;
  dl = l0 / n; dh = h0 / n; ds = s0 / n;
  for (i = 1; i <= n; ++i) { double j = i - 0.5; acc += dl * (j * dh) * g_ext(j * ds * c_ext); }
  return acc;
}
/* gomtrc_coeffs :135-246; slice is 0-based here */
static void g_gomtrc(const double *sundir, const double *f1, const double *f2, const double *f3, const double *f4, const double *f5,
                     const double *v1, const double *v2, const double *v3, const double *v4, double c_ext,
                     const int slice[3], double *coeffs, double sun_up_down, int nint) {
  const double small = sqrt(DBL_EPSILON);
  gv3 n0, normal, pl, pb, pt, pr, d, d2, hs;
  double s, cos_src_trgt, area1, area2, area3, aq, at, tot, sum;
  g_normal(f1, f2, f3, n0);
  s = g_hit_plane(f1, sundir, f5, n0); hs[0] = s * sundir[0]; hs[1] = s * sundir[1]; hs[2] = s * sundir[2]; s = g_norm(hs);
  g_normal(f1, f2, f5, normal);
  g_sub(f1, f2, d); g_sub(f1, f4, d2);
  cos_src_trgt = cos(acos(g_dot(d, d2) / (g_norm(d) * g_norm(d2))) - ORC_PI / 2);
#define G_S(v, f, nrm) do { double h_ = g_hit_plane(v, sundir, f, nrm); hs[0] = h_ * sundir[0]; hs[1] = h_ * sundir[1]; hs[2] = h_ * sundir[2]; s = g_norm(hs); } while (0)
#define G_LEN(a, b) (g_sub(a, b, d), g_norm(d))
  g_sub(v1, f1, d);
  if (g_norm(d) > small) {
    g_proj_var_to_edges(f1, f2, f3, f4, v1, pl, pb, pt, pr);
    area1 = g_quad(v1, pb, f3, pl) * exp(-c_ext * s);
    G_S(v1, f1, normal);
    aq = g_quad(v1, pr, f2, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f1, pr) * cos_src_trgt, G_LEN(pr, v1), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v1, f1, normal);
    aq = g_quad(v1, pl, f4, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f1, pt) * cos_src_trgt, G_LEN(pt, v1), c_ext, nint);
  } else if ((g_sub(v2, f2, d), g_norm(d)) > small) {
    g_proj_var_to_edges(f2, f1, f4, f3, v2, pl, pt, pb, pr);
    area1 = g_quad(v2, pl, f4, pt) * exp(-c_ext * s);
    G_S(v2, f2, normal);
    aq = g_quad(v2, pr, f1, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f2, pr) * cos_src_trgt, G_LEN(pr, v2), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v2, f2, normal);
    aq = g_quad(v2, pl, f3, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f2, pb) * cos_src_trgt, G_LEN(pb, v2), c_ext, nint);
  } else if ((g_sub(v3, f3, d), g_norm(d)) > small) {
    g_proj_var_to_edges(f4, f3, f2, f1, v3, pr, pb, pt, pl);
    area1 = g_quad(v3, pt, f1, pr) * exp(-c_ext * s);
    G_S(v3, f3, normal);
    aq = g_quad(v3, pl, f4, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f3, pl) * cos_src_trgt, G_LEN(pl, v3), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v3, f3, normal);
    aq = g_quad(v3, pr, f2, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f3, pb) * cos_src_trgt, G_LEN(pb, v3), c_ext, nint);
  } else if ((g_sub(v4, f4, d), g_norm(d)) > small) {
    g_proj_var_to_edges(f4, f3, f2, f1, v4, pr, pb, pt, pl);
    area1 = g_quad(v4, pr, f2, pb) * exp(-c_ext * s);
    G_S(v4, f4, normal);
    aq = g_quad(v4, pl, f3, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f4, pl) * cos_src_trgt, G_LEN(pl, v4), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v4, f4, normal);
    aq = g_quad(v4, pr, f1, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f4, pt) * cos_src_trgt, G_LEN(pt, v4), c_ext, nint);
  } else {
    area1 = g_quad(f1, f2, f3, f4) * exp(-c_ext * s);
    area2 = 0.0; aq = 0.0; at = 0.0;
  }
#undef G_S
#undef G_LEN
  area3 = aq + at;
  area1 = area1 - sun_up_down * area3;
  area3 = area3 + sun_up_down * area3;
  coeffs[slice[0]] = dmax(area1, 0.0); coeffs[slice[1]] = dmax(area2, 0.0); coeffs[slice[2]] = dmax(area3, 0.0);
  tot = g_quad(f1, f2, f3, f4);
  sum = coeffs[slice[0]] + coeffs[slice[1]] + coeffs[slice[2]];
  tot = dmax(tot, sum);
  coeffs[slice[0]] /= tot; coeffs[slice[1]] /= tot; coeffs[slice[2]] /= tot;
}
/* dir2dir3_geometric_coeffs :34-133.  Returns 1 if the extinction coefficient is too small (CHKERR there). */
int orc_dir2dir3_geometric_coeffs(const double *vertices, const double *sundir, double c_ext, int num_intervals, double *coeffs) {
  const double *a = vertices, *b = vertices + 3, *c = vertices + 6, *d = vertices + 9, *e = vertices + 12,
               *f = vertices + 15, *g = vertices + 18, *h = vertices + 21;
  static const int sl_z[3] = {0, 6, 3}, sl_x[3] = {4, 7, 1}, sl_y[3] = {8, 5, 2};
  gv3 n, ms, p1, p2, p3, p4;
  double sun_up_down;
  if (c_ext < DBL_MIN) return 1;
  g_normal(e, f, h, n);
  ms[0] = -sundir[0]; ms[1] = -sundir[1]; ms[2] = -sundir[2];
  sun_up_down = dmin(g_dot(ms, n) >= 0.0 ? 1.0 : -1.0, 0.0);
  /* src z */
  g_copy(h, p1); g_copy(g, p2); g_copy(e, p3); g_copy(f, p4);
  g_normal(d, b, a, n);
  g_project_points(sundir, d, n, p1, p2, p3, p4);
  g_rearange_projections(d, c, a, b, p1, p2, p3, p4);
  g_gomtrc(sundir, d, c, a, b, g, p1, p2, p3, p4, c_ext, sl_z, coeffs, sun_up_down, num_intervals);
  /* src x */
  g_copy(h, p1); g_copy(d, p2); g_copy(b, p3); g_copy(f, p4);
  g_normal(c, a, e, n);
  g_project_points(sundir, a, n, p1, p2, p3, p4);
  g_rearange_projections(g, c, a, e, p1, p2, p3, p4);
  g_gomtrc(sundir, g, c, a, e, d, p1, p2, p3, p4, c_ext, sl_x, coeffs, sun_up_down, num_intervals);
  /* src y */
  g_copy(f, p1); g_copy(b, p2); g_copy(a, p3); g_copy(e, p4);
  g_normal(c, d, h, n);
  g_project_points(sundir, d, n, p1, p2, p3, p4);
  g_rearange_projections(h, d, c, g, p1, p2, p3, p4);
  g_gomtrc(sundir, h, d, c, g, b, p1, p2, p3, p4, c_ext, sl_y, coeffs, sun_up_down, num_intervals);
  return 0;
}

/* test hooks for the geometry helpers above (tests/test_intersection/test_intersection.F90) */
double orc_hit_plane(const double *p_loc, const double *p_dir, const double *po, const double *pn) { return g_hit_plane(p_loc, p_dir, po, pn); }
void orc_compute_normal_3d(const double *p1, const double *p2, const double *p3, double *n) { g_normal(p1, p2, p3, n); }
int orc_line_intersection_3d(const double *p1, const double *d1, const double *p3, const double *d2, double *c1, double *c2) {
  return g_line_intersection(p1, d1, p3, d2, c1, c2);
}
double orc_triangle_area_by_vertices(const double *v1, const double *v2, const double *v3) { return g_tri(v1, v2, v3); }

/* ---- buildings (src/buildings.F90) ----------------------------------------------------------------------------------
 * A building is a set of faces of grid cells: iface(m) is the 1-based flattened index into (6, Nz, Nx, Ny), first axis
 * fastest (faceidx_by_cell_plus_offset, src/buildings.F90:215-222; ndarray_offsets helper_functions.fypp:2432-2437),
 * face ids PPRTS_TOP_FACE = 1 (z+0), BOT = 2 (z+1), LEFT = 3 (x+0), RIGHT = 4 (x+1), REAR = 5 (y+0), FRONT = 6 (y+1)
 * (src/boxmc_geometry.F90:45-51).  nfaces = 0 removes them (opt_buildings not present). */
enum { ORC_TOP_FACE = 1, ORC_BOT_FACE, ORC_LEFT_FACE, ORC_RIGHT_FACE, ORC_REAR_FACE, ORC_FRONT_FACE };
int orc_set_buildings(orc_solver *s, int nfaces, const int *iface, const double *albedo, const double *planck) {
  int m;
  free(s->bld_iface); free(s->bld_albedo); free(s->bld_planck);
  s->bld_iface = 0; s->bld_albedo = s->bld_planck = 0; s->nbld = 0; s->bld_have_planck = 0;
  if (nfaces <= 0) return 0;
  for (m = 0; m < nfaces; ++m) if (iface[m] < 1 || (long)iface[m] > 6L * s->Nz * s->Nx * s->Ny) return 1;
  s->nbld = nfaces;
  s->bld_iface = (int *)malloc(sizeof(int) * nfaces); memcpy(s->bld_iface, iface, sizeof(int) * nfaces);
  s->bld_albedo = (double *)malloc(sizeof(double) * nfaces); memcpy(s->bld_albedo, albedo, sizeof(double) * nfaces);
  if (planck) {
    s->bld_planck = (double *)malloc(sizeof(double) * nfaces); memcpy(s->bld_planck, planck, sizeof(double) * nfaces);
    s->bld_have_planck = 1;
  }
  return 0;
}
/* ind_1d_to_nd (helper_functions.fypp:2439-2454) followed by idx(2:4) - 1: face id (1-based), cell (k,i,j) 0-based */
static void bld_idx(const orc_solver *s, int m, int *face, int *k, int *i, int *j) {
  long r = (long)s->bld_iface[m] - 1;
  *face = (int)(r % 6) + 1; r /= 6;
  *k = (int)(r % s->Nz); r /= s->Nz;
  *i = (int)(r % s->Nx); r /= s->Nx;
  *j = (int)r;
}
/* alloc_coeff_dir2dir::set_buildings_coeff (src/pprts.F90:3194-3212): a building cell passes no direct radiation */
static void set_buildings_coeff_dir2dir(orc_solver *s) {
  LOCALS
  int m, face, k, i, j, n;
  for (m = 0; m < s->nbld; ++m) {
    bld_idx(s, m, &face, &k, &i, &j);
    for (n = 0; n < Nd * Nd; ++n) CF(s->dir2dir, Nd * Nd, n, k, i, j) = 0.0;
  }
}
/* alloc_coeff_diff2diff::set_buildings_coeff (src/pprts.F90:3579-3677): every stream leaving the cell through a
 * building face receives nothing but the Lambertian reflection albedo/streams of the streams entering through that
 * face.  v(src,dst) = c[dst*Nf + src]. */
static void set_buildings_coeff_diff2diff(orc_solver *s) {
  LOCALS
  int m, face, k, i, j, isrc, idst, src;
  for (m = 0; m < s->nbld; ++m) {
    double *c;
    bld_idx(s, m, &face, &k, &i, &j);
    c = &CF(s->diff2diff, Nf * Nf, 0, k, i, j);
    if (face == ORC_TOP_FACE || face == ORC_BOT_FACE) {
      const int out_is_inward = face == ORC_BOT_FACE; /* top: eup leaves, bottom: edn leaves */
      for (idst = 0; idst < Nft; ++idst) if (s->st.ft_inward[idst] == out_is_inward) {
        for (src = 0; src < Nf; ++src) c[idst * Nf + src] = 0.0;
        for (isrc = 0; isrc < Nft; ++isrc) if (s->st.ft_inward[isrc] != out_is_inward)
          c[idst * Nf + isrc] = s->bld_albedo[m] / (double)(Nft / 2);
      }
    } else {
      const int off = (face == ORC_LEFT_FACE || face == ORC_RIGHT_FACE) ? Nft : Nft + Nfs;
      const int out_is_inward = face == ORC_RIGHT_FACE || face == ORC_FRONT_FACE;
      for (idst = 0; idst < Nfs; ++idst) if (s->st.fs_inward[idst] == out_is_inward) {
        for (src = 0; src < Nf; ++src) c[(off + idst) * Nf + src] = 0.0;
        for (isrc = 0; isrc < Nfs; ++isrc) if (s->st.fs_inward[isrc] != out_is_inward)
          c[(off + idst) * Nf + off + isrc] = s->bld_albedo[m] / (double)(Nfs / 2);
      }
    }
  }
}

/* ---- coefficients (src/pprts.F90:3088-3494) ---------------------------------------------------------------------- */
static void store_coeff(const orc_solver *s, const double *v, double *dst, int n) {
  int c;
  if (s->cfg.coeff_fp32) for (c = 0; c < n; ++c) dst[c] = (double)(float)v[c];
  else for (c = 0; c < n; ++c) dst[c] = v[c];
}
/* determine_vertex_heights (src/pprts.F90:726-755): height of vertex (k, i, j), k an atmosphere level, i in -1..Nx,
 * j in -1..Ny: minus the layer thicknesses above it in the column it belongs to, clamped into the local domain */
static double vert_height(const orc_solver *s, int katm, int i, int j) {
  const int Nza = s->Nz_atm, Nx = s->Nx;
  const int ci = i < 0 ? 0 : (i > s->Nx - 1 ? s->Nx - 1 : i), cj = j < 0 ? 0 : (j > s->Ny - 1 ? s->Ny - 1 : j);
  double h = 0.0;
  int k;
  for (k = 0; k < katm; ++k) h = h - s->dz_atm[k + (long)Nza * (ci + (long)Nx * cj)];
  return h;
}
/* setup_default_unit_cube_geometry(dx, dy, -1) + init_vertices (src/pprts.F90:3057-3085, :3130-3143) with
 * ltop_bottom_faces_planar = ltop_bottom_planes_parallel = .true. (their defaults under -pprts_geometric_coeffs)
 * unless cfg.geometric_coeffs carries bit 2 (-pprts_top_bottom_faces_planar no) / bit 4
 * (-pprts_top_bottom_planes_parallel no), then the z shift of :3143 */
static void cell_vertices(const orc_solver *s, int k, int i, int j, double *v) {
  const int ka = k + s->icollapse - 1; /* atmk(k) */
  const double dz = s->dz_atm[ka + (long)s->Nz_atm * (i + (long)s->Nx * j)];
  double zmin;
  int q;
  for (q = 0; q < 24; ++q) v[q] = 0.0;
  v[3] = s->dx; v[7] = s->dy; v[9] = s->dx; v[10] = s->dy;
  v[15] = s->dx; v[19] = s->dy; v[21] = s->dx; v[22] = s->dy;
  v[2] = vert_height(s, ka + 1, i, j);       /* hs(2,1,1) */
  v[5] = vert_height(s, ka + 1, i + 1, j);   /* hs(2,2,1) */
  v[8] = vert_height(s, ka + 1, i, j + 1);   /* hs(2,1,2) */
  v[11] = vert_height(s, ka + 1, i + 1, j + 1); /* hs(2,2,2) */
  v[14] = vert_height(s, ka, i, j);          /* hs(1,1,1) */
  v[17] = vert_height(s, ka, i + 1, j);      /* hs(1,2,1) */
  v[20] = vert_height(s, ka, i, j + 1);      /* hs(1,1,2) */
  v[23] = vert_height(s, ka, i + 1, j + 1);  /* hs(1,2,2) */
  if (!(s->cfg.geometric_coeffs & 2)) {      /* planar faces */
    v[11] = v[8] + (v[5] - v[2]);
    v[23] = v[20] + (v[17] - v[14]);
  }
  if (!(s->cfg.geometric_coeffs & 4)) for (q = 0; q < 4; ++q) v[14 + 3 * q] = v[2 + 3 * q] + dz; /* parallel planes */
  zmin = v[2];
  for (q = 1; q < 8; ++q) zmin = dmin(zmin, v[2 + 3 * q]);
  for (q = 0; q < 8; ++q) v[2 + 3 * q] -= zmin;
}
static int alloc_coeffs(orc_solver *s, int lsolar) {
  LOCALS
  long col;
  int bad = 0;
  if (s->cfg.geometric_coeffs && lsolar && Nd != 3) return 1; /* 'only implemented for 3 direct streams', :3124-3125 */
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (col = 0; col < (long)Nx * Ny; ++col) {
    int i = (int)(col % Nx), j = (int)(col / Nx), k;
    double v[ORC_MAX_DOF * ORC_MAX_DOF];
    for (k = 0; k < Nz; ++k) {
      if (s->l1d[k]) continue;
      double ka = A3(s->kabs, k, i, j), ks = A3(s->ksca, k, i, j), gg = A3(s->g, k, i, j), dz = A3(s->dz, k, i, j);
      if (lsolar && s->cfg.geometric_coeffs) {
        /* -pprts_geometric_coeffs (:3142-3149) with -pprts_conserve_lut_atm_abso (its default then, :3272-3273, :3324-3351):
         * dir2dir from the geometry of the cell, dir2diff = the LUT's distribution over the diffuse streams scaled so
         * that every source loses to absorption exactly what the LUT says */
        double T_LUT[9], S_LUT[3 * ORC_MAX_DOF], T_G[9], vert[24];
        int src, dst;
        orc_get_coeff(s->lut, &s->st, ka, ks, gg, dz, s->dx, 1, s->symmetry_phi, s->theta, s->xinc == 0, s->yinc == 0, v);
        store_coeff(s, v, T_LUT, Nd * Nd);
        orc_get_coeff(s->lut, &s->st, ka, ks, gg, dz, s->dx, 2, s->symmetry_phi, s->theta, s->xinc == 0, s->yinc == 0, v);
        store_coeff(s, v, S_LUT, Nd * Nf);
        cell_vertices(s, k, i, j, vert);
        if (orc_dir2dir3_geometric_coeffs(vert, s->sundir, ka + ks, 0, T_G)) { bad |= 2; continue; }
        for (src = 0; src < Nd; ++src) {
          double Sn = 0, Tn = 0, Tg = 0;
          for (dst = 0; dst < Nf; ++dst) Sn += S_LUT[dst * Nd + src];
          for (dst = 0; dst < Nd; ++dst) { Tn += T_LUT[dst * Nd + src]; Tg += T_G[dst * Nd + src]; }
          for (dst = 0; dst < Nf; ++dst) {
            double sg = (Sn <= DBL_EPSILON) ? 0.0 : S_LUT[dst * Nd + src] / Sn;
            CF(s->dir2diff, Nd * Nf, dst * Nd + src, k, i, j) = sg * (1.0 - (1.0 - Tn - Sn) - Tg);
          }
        }
        for (dst = 0; dst < Nd * Nd; ++dst) CF(s->dir2dir, Nd * Nd, dst, k, i, j) = T_G[dst];
      } else if (lsolar) {
        orc_get_coeff(s->lut, &s->st, ka, ks, gg, dz, s->dx, 1, s->symmetry_phi, s->theta, s->xinc == 0, s->yinc == 0, v);
        store_coeff(s, v, &CF(s->dir2dir, Nd * Nd, 0, k, i, j), Nd * Nd);
        orc_get_coeff(s->lut, &s->st, ka, ks, gg, dz, s->dx, 2, s->symmetry_phi, s->theta, s->xinc == 0, s->yinc == 0, v);
        store_coeff(s, v, &CF(s->dir2diff, Nd * Nf, 0, k, i, j), Nd * Nf);
      }
      orc_get_coeff(s->lut, &s->st, ka, ks, gg, dz, s->dx, 3, 0, 0, 0, 0, v);
      store_coeff(s, v, &CF(s->diff2diff, Nf * Nf, 0, k, i, j), Nf * Nf);
    }
  }
  if (bad) return bad;
  if (s->nbld) { /* :3188-3190, :3492-3494 */
    if (lsolar) set_buildings_coeff_dir2dir(s);
    set_buildings_coeff_diff2diff(s);
  }
  return 0;
}

/* ---- scale_flx (src/pprts.F90:3682-3741) ------------------------------------------------------------------------ */
void orc_scale_flx_arr(orc_solver *s, double *arr, int ldir, int lWm2) {
  long n, N = nin(s, ldir ? s->st.Nd : s->st.Nf);
  const double *sc = ldir ? s->dir_scale : s->diff_scale;
  for (n = 0; n < N; ++n) {
    if (lWm2) arr[n] = arr[n] * ((sc[n] != 0.0) ? 1.0 / sc[n] : 0.0);
    else arr[n] = arr[n] * sc[n];
  }
}
static void scale_flx(orc_solver *s, orc_solution *sol, int lWm2) {
  if (sol->lsolar_rad && sol->lWm2_dir != lWm2) { orc_scale_flx_arr(s, sol->edir, 1, lWm2); sol->lWm2_dir = lWm2; }
  if (sol->lWm2_diff != lWm2) { orc_scale_flx_arr(s, sol->ediff, 0, lWm2); sol->lWm2_diff = lWm2; }
}

/* src/pprts_base.F90:1097-1142 (determine_ksp_tolerances, non-PETSc branch) */
static void determine_ksp_tolerances(const orc_solver *s, double *rtol, double *atol) {
  *rtol = 1e-5;
  *atol = dmax(1e-8, 1e-4 * (double)s->Nx * (double)s->Ny * (double)s->Nz1 * s->unconstrained_fraction);
}

/* ---- direct beam (src/pprts_explicit.F90:60-459) ----------------------------------------------------------------- */
static void exchange_direct_boundary(const orc_solver *s, int lsun_north, int lsun_east, double *x) {
  LOCALS
  int d, k, i, j;
  /* one periodic rank is its own neighbour and owns both domain edges: with -pprts_open_bc the send buffer of the edge
   * is zero (:262-263, :270-271, :287-288, :295-296) */
  const int lopen = s->cfg.open_bc;
  for (d = Ndt; d < Ndt + Nds; ++d) for (j = 0; j < Ny; ++j) for (k = 0; k < Nz1; ++k) {
    if (lsun_east) FG(x, Nd, d, k, Nx, j) = lopen ? 0.0 : FG(x, Nd, d, k, 0, j);  /* send xs west, recv from east into xe+1 */
    else           FG(x, Nd, d, k, 0, j) = lopen ? 0.0 : FG(x, Nd, d, k, Nx, j);  /* send xe+1 east, recv from west into xs */
  }
  for (d = Ndt + Nds; d < Nd; ++d) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) {
    if (lsun_north) FG(x, Nd, d, k, i, Ny) = lopen ? 0.0 : FG(x, Nd, d, k, i, 0);
    else            FG(x, Nd, d, k, i, 0) = lopen ? 0.0 : FG(x, Nd, d, k, i, Ny);
  }
}

/* single_column_solve (src/pprts_base.F90:1318-1458): the direct beam of ONE column that is its own horizontal
 * neighbour, for unit incidence on every top stream.  The reference assembles  A = -I + T  over (dof, level) and calls
 * KSPSolve with PETSc's defaults; A is block lower triangular in the levels with dense (side x side) diagonal blocks, so
 * ILU(0) is the exact factorisation and the solve is a direct one.  Here: per 3-D layer the 2 Nds side unknowns from
 * (I - M_ss) x_s = M_st top(k) by Gaussian elimination, then top(k+1); a 1-D layer has top(k+1) = exp(-kabs dz / cos
 * theta) top(k) (:1355-1365 -- kabs alone, as written there) and no side flux.  col = (Nd, Nz1), sides of layer k at level k. */
static void single_column_solve(const orc_solver *s, int i, int j, double *col) {
  LOCALS
  const int ns = 2 * Nds;
  int k, p, q, r, src, dst;
  double v[ORC_MAX_DOF * ORC_MAX_DOF], c[ORC_MAX_DOF * ORC_MAX_DOF];
  double M[ORC_MAX_DOF][ORC_MAX_DOF + 1];
  memset(col, 0, sizeof(double) * Nd * Nz1);
  for (src = 0; src < Ndt; ++src) col[src] = 1.0;
  for (k = 0; k < Nz; ++k) {
    double *lev = col + (long)k * Nd, *nxt = lev + Nd;
    if (s->l1d[k]) {
      const double vv = exp(-(A3(s->kabs, k, i, j) * A3(s->dz, k, i, j) / s->costheta));
      for (src = 0; src < Ndt; ++src) nxt[src] = vv * lev[src];
      continue;
    }
    orc_get_coeff(s->lut, &s->st, A3(s->kabs, k, i, j), A3(s->ksca, k, i, j), A3(s->g, k, i, j), A3(s->dz, k, i, j), s->dx, 1,
                  s->symmetry_phi, s->theta, s->xinc == 0, s->yinc == 0, v);
    store_coeff(s, v, c, Nd * Nd); /* cdir2dir(src, dst) = c[dst * Nd + src] */
    for (p = 0; p < ns; ++p) {
      double rhs = 0;
      for (q = 0; q < ns; ++q) M[p][q] = (p == q ? 1.0 : 0.0) - c[(Ndt + p) * Nd + Ndt + q];
      for (src = 0; src < Ndt; ++src) rhs += c[(Ndt + p) * Nd + src] * lev[src];
      M[p][ns] = rhs;
    }
    for (p = 0; p < ns; ++p) { /* no pivoting: I - M_ss is diagonally dominant by columns (energy conservation) */
      for (r = p + 1; r < ns; ++r) {
        const double f = M[r][p] / M[p][p];
        for (q = p; q <= ns; ++q) M[r][q] -= f * M[p][q];
      }
    }
    for (p = ns - 1; p >= 0; --p) {
      double acc = M[p][ns];
      for (q = p + 1; q < ns; ++q) acc -= M[p][q] * lev[Ndt + q];
      lev[Ndt + p] = acc / M[p][p];
    }
    for (dst = 0; dst < Ndt; ++dst) {
      double acc = 0;
      for (src = 0; src < Nd; ++src) acc += c[dst * Nd + src] * lev[src];
      nxt[dst] = acc;
    }
  }
}

/* set_open_bc (src/pprts_base.F90:1189-1316): side inflow through the sunward domain edges = the side fluxes of the
 * single-column solution of the edge column, scaled like the TOA incidence.  lb is ghosted; after the reference's
 * LocalToGlobal(ADD) + GlobalToLocal round trip the value sits both in the ghost face and in its periodic owner. */
static void set_open_bc(const orc_solver *s, double fac, double *lb) {
  LOCALS
  const int lsun_north = s->yinc == 0, lsun_east = s->xinc == 0;
  double *col = (double *)malloc(sizeof(double) * Nd * Nz1);
  int i, j, k, src;
  j = lsun_north ? Ny - 1 : 0;
  for (i = 0; i < Nx; ++i) {
    single_column_solve(s, i, j, col);
    for (k = 0; k < Nz; ++k) for (src = 0; src < Nds; ++src) {
      const int ioff = Ndt + Nds + src;
      const double val = fac * col[(long)k * Nd + ioff];
      FG(lb, Nd, ioff, k, i, j + 1 - s->yinc) = val;
      if (lsun_north) FG(lb, Nd, ioff, k, i, 0) = val;
    }
  }
  i = lsun_east ? Nx - 1 : 0;
  for (j = 0; j < Ny; ++j) {
    single_column_solve(s, i, j, col);
    for (k = 0; k < Nz; ++k) for (src = 0; src < Nds; ++src) {
      const int ioff = Ndt + src;
      const double val = fac * col[(long)k * Nd + ioff];
      FG(lb, Nd, ioff, k, i + 1 - s->xinc, j) = val;
      if (lsun_east) FG(lb, Nd, ioff, k, 0, j) = val;
    }
  }
  free(col);
}

static void explicit_edir_forward_sweep(const orc_solver *s, const int dxr[3], const int dyr[3], const double *b,
                                        double *x) {
  LOCALS
  const int xinc = s->xinc, yinc = s->yinc;
  int i, j, k, idst, isrc, src, dst;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (dst = 0; dst < Ndt; ++dst)
    FG(x, Nd, dst, 0, i, j) = FG(b, Nd, dst, 0, i, j);
  if (s->cfg.open_bc) { /* :358-378: the sunward edge faces take the boundary inflow */
    const int jb = yinc == 0 ? Ny : 0, ib = xinc == 0 ? Nx : 0;
    for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (dst = Ndt + Nds; dst < Nd; ++dst)
      FG(x, Nd, dst, k, i, jb) = FG(b, Nd, dst, k, i, jb);
    for (j = 0; j < Ny; ++j) for (k = 0; k < Nz1; ++k) for (dst = Ndt; dst < Ndt + Nds; ++dst)
      FG(x, Nd, dst, k, ib, j) = FG(b, Nd, dst, k, ib, j);
  }
  for (k = 0; k < Nz; ++k) {
    if (s->l1d[k]) {
      for (j = dyr[0]; j != dyr[1] + dyr[2]; j += dyr[2]) for (i = dxr[0]; i != dxr[1] + dxr[2]; i += dxr[2])
        for (idst = 0; idst < Ndt; ++idst) FG(x, Nd, idst, k + 1, i, j) = FG(x, Nd, idst, k, i, j) * A3(s->a33, k, i, j);
    } else {
      for (j = dyr[0]; j != dyr[1] + dyr[2]; j += dyr[2]) for (i = dxr[0]; i != dxr[1] + dxr[2]; i += dxr[2]) {
        const double *v = &CF(s->dir2dir, Nd * Nd, 0, k, i, j); /* v(src,dst) = v[dst*Nd+src] */
        double acc;
        dst = 0;
        for (idst = 0; idst < Ndt; ++idst, ++dst) {
          acc = 0; src = 0;
          for (isrc = 0; isrc < Ndt; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i, j) * v[dst * Nd + src];
          for (isrc = 0; isrc < Nds; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i + 1 - xinc, j) * v[dst * Nd + src];
          for (isrc = 0; isrc < Nds; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i, j + 1 - yinc) * v[dst * Nd + src];
          FG(x, Nd, dst, k + 1, i, j) = acc;
        }
        for (idst = 0; idst < Nds; ++idst, ++dst) {
          acc = 0; src = 0;
          for (isrc = 0; isrc < Ndt; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i, j) * v[dst * Nd + src];
          for (isrc = 0; isrc < Nds; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i + 1 - xinc, j) * v[dst * Nd + src];
          for (isrc = 0; isrc < Nds; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i, j + 1 - yinc) * v[dst * Nd + src];
          FG(x, Nd, dst, k, i + xinc, j) = acc;
        }
        for (idst = 0; idst < Nds; ++idst, ++dst) {
          acc = 0; src = 0;
          for (isrc = 0; isrc < Ndt; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i, j) * v[dst * Nd + src];
          for (isrc = 0; isrc < Nds; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i + 1 - xinc, j) * v[dst * Nd + src];
          for (isrc = 0; isrc < Nds; ++isrc, ++src) acc = acc + FG(x, Nd, src, k, i, j + 1 - yinc) * v[dst * Nd + src];
          FG(x, Nd, dst, k, i, j + yinc) = acc;
        }
      }
    }
  }
}

static double norm2_diff_and_copy(const orc_solver *s, int nd, double *xg, const double *v0) {
  LOCALS
  int d, k, i, j;
  double acc = 0;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (d = 0; d < nd; ++d) {
    double df = FI(xg, nd, d, k, i, j) - FG(v0, nd, d, k, i, j);
    acc += df * df;
    FI(xg, nd, d, k, i, j) = FG(v0, nd, d, k, i, j);
  }
  return sqrt(acc);
}

static int explicit_edir(orc_solver *s, double edirTOA, orc_solution *sol) {
  LOCALS
  int maxiter = s->cfg.maxit_dir > 0 ? s->cfg.maxit_dir : 1000, iter, i, j, d, converged = 0;
  double rtol_d, atol_d, rtol, atol, res, res1 = 0, rel, fac;
  double *lb = (double *)calloc(ngh(s, Nd), sizeof(double));
  double *v0 = (double *)calloc(ngh(s, Nd), sizeof(double));
  int dxr[3] = {0, Nx - 1, 1}, dyr[3] = {0, Ny - 1, 1};
  const int lsun_north = s->yinc == 0, lsun_east = s->xinc == 0;
  determine_ksp_tolerances(s, &rtol_d, &atol_d);
  rtol = s->cfg.rtol >= 0 ? s->cfg.rtol : rtol_d;
  atol = s->cfg.atol >= 0 ? s->cfg.atol : atol_d;
  if (sol->dir_hist[0] < 0) { /* -ksp_complete_initial_run, src/pprts_explicit.F90:114-121 */
    maxiter = maxiter + 2;
    rtol = dmin(rtol, rtol_d);
    atol = dmin(atol, atol_d);
  }
  if (lsun_east) { dxr[0] = Nx - 1; dxr[1] = 0; dxr[2] = -1; }  /* -explicit_edir_permute (default on) */
  if (lsun_north) { dyr[0] = Ny - 1; dyr[1] = 0; dyr[2] = -1; }
  /* setup_incSolar, src/pprts_base.F90:1146-1181 */
  fac = edirTOA * s->dx * s->dy / (double)s->st.dirtop_adiv;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (d = 0; d < Ndt; ++d) FG(lb, Nd, d, 0, i, j) = fac;
  if (s->cfg.open_bc) set_open_bc(s, fac, lb);
  interior_to_ghosted(s, Nd, sol->edir, v0);
  exchange_direct_boundary(s, lsun_north, lsun_east, v0);
  sol->Niter_dir = 0;
  for (iter = 1; iter <= maxiter; ++iter) {
    explicit_edir_forward_sweep(s, dxr, dyr, lb, v0);
    exchange_direct_boundary(s, lsun_north, lsun_east, v0);
    res = dmax(DBL_MIN, norm2_diff_and_copy(s, Nd, sol->edir, v0));
    if (iter == 1) res1 = res;
    rel = (res1 <= sqrt(DBL_MIN)) ? 0.0 : res / res1;
    sol->dir_hist[(iter < ORC_NHIST ? iter : ORC_NHIST) - 1] = res;
    sol->Niter_dir = iter;
    sol->final_res_dir = res;
    if (res < atol || rel < rtol) { converged = 1; break; }
  }
  ghosted_to_interior(s, Nd, v0, sol->edir);
  sol->lchanged = 1;
  sol->lWm2_dir = 0;
  free(lb); free(v0);
  return converged ? 0 : 1;
}

/* ---- source term (src/pprts.F90:4641-4987) ----------------------------------------------------------------------- */
static void setup_b(orc_solver *s, orc_solution *sol) {
  LOCALS
  const double eps = DBL_EPSILON;
  double *xsrc = (double *)calloc(ngh(s, Nf), sizeof(double));
  double *xedir = 0;
  int k, i, j, src, dst, idof, idofdst, idiff, iside;
  if (sol->lsolar_rad) { /* set_solar_source :4684-4846 */
    const int xinc = s->xinc, yinc = s->yinc;
    xedir = (double *)calloc(ngh(s, Nd), sizeof(double));
    interior_to_ghosted(s, Nd, sol->edir, xedir);
    halo_fill_5pt(s, Nd, xedir);
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
      int any = 0;
      for (src = 0; src < Nd; ++src) if (FG(xedir, Nd, src, k, i, j) > eps) any = 1;
      if (!any) continue;
      if (s->l1d[k]) {
        for (src = 0; src < Ndt; ++src) for (idiff = 0; idiff < Nft; ++idiff) {
          if (s->st.ft_inward[idiff])
            FG(xsrc, Nf, idiff, k + 1, i, j) += FG(xedir, Nd, src, k, i, j) * A3(s->a23, k, i, j) / (double)(Nft / 2);
          else
            FG(xsrc, Nf, idiff, k, i, j) += FG(xedir, Nd, src, k, i, j) * A3(s->a13, k, i, j) / (double)(Nft / 2);
        }
      } else {
        const double *c = &CF(s->dir2diff, Nd * Nf, 0, k, i, j); /* dir2diff(dst*Nd + src) */
        double solrad;
        dst = 0;
        for (idofdst = 0; idofdst < Nft; ++idofdst, ++dst) {
          double *tgt = s->st.ft_inward[idofdst] ? &FG(xsrc, Nf, dst, k + 1, i, j) : &FG(xsrc, Nf, dst, k, i, j);
          src = 0;
          for (idof = 0; idof < Ndt; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i, j); *tgt += solrad * c[dst * Nd + src]; }
          for (idof = 0; idof < Nds; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i + 1 - xinc, j); *tgt += solrad * c[dst * Nd + src]; }
          for (idof = 0; idof < Nds; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i, j + 1 - yinc); *tgt += solrad * c[dst * Nd + src]; }
        }
        for (idofdst = 0; idofdst < Nfs; ++idofdst, ++dst) {
          double *tgt = s->st.fs_inward[idofdst] ? &FG(xsrc, Nf, dst, k, i + 1, j) : &FG(xsrc, Nf, dst, k, i, j);
          src = 0;
          for (idof = 0; idof < Ndt; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i, j); *tgt += solrad * c[dst * Nd + src]; }
          for (idof = 0; idof < Nds; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i + 1 - xinc, j); *tgt += solrad * c[dst * Nd + src]; }
          for (idof = 0; idof < Nds; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i, j + 1 - yinc); *tgt += solrad * c[dst * Nd + src]; }
        }
        for (idofdst = 0; idofdst < Nfs; ++idofdst, ++dst) {
          double *tgt = s->st.fs_inward[idofdst] ? &FG(xsrc, Nf, dst, k, i, j + 1) : &FG(xsrc, Nf, dst, k, i, j);
          src = 0;
          for (idof = 0; idof < Ndt; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i, j); *tgt += solrad * c[dst * Nd + src]; }
          for (idof = 0; idof < Nds; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i + 1 - xinc, j); *tgt += solrad * c[dst * Nd + src]; }
          for (idof = 0; idof < Nds; ++idof, ++src) { solrad = FG(xedir, Nd, src, k, i, j + 1 - yinc); *tgt += solrad * c[dst * Nd + src]; }
        }
      }
    }
    /* ground albedo reflecting direct radiation :4830-4843 */
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (dst = 0; dst < Nft; ++dst) if (!s->st.ft_inward[dst])
      for (src = 0; src < Ndt; ++src)
        FG(xsrc, Nf, dst, Nz, i, j) += FG(xedir, Nd, src, Nz, i, j) * A2(s->albedo, i, j) / (double)(Nft / 2);
  }
  if (sol->lthermal_rad) { /* set_thermal_source :4848-4987 */
    const double Az = s->dx * s->dy; /* difftop%area_divider = 1 */
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) {
      for (k = 0; k < Nz; ++k) {
        double b0 = P3(s->planck, k, i, j), b1 = P3(s->planck, k + 1, i, j);
        double tauz = A3(s->kabs, k, i, j) * A3(s->dz, k, i, j), btop, bbot, bfac, emis;
        if (s->l1d[k]) {
          bfac = ORC_PI * Az / (double)(Nft / 2);
          if (s->icollapse > 1 && k == 0) { /* :4875-4877 */
            btop = A2(s->Btop, i, j) * bfac;
            bbot = A2(s->Bbot, i, j) * bfac;
          } else {
            emis = 1.0 - A3(s->a11, k, i, j) - A3(s->a12, k, i, j);
            emis = dmax(0.0, dmin(1.0, emis));
            btop = orc_B_eff(b1, b0, tauz) * bfac * emis;
            bbot = orc_B_eff(b0, b1, tauz) * bfac * emis;
          }
          for (src = 0; src < Nft; ++src) {
            if (s->st.ft_inward[src]) FG(xsrc, Nf, src, k + 1, i, j) += bbot;
            else FG(xsrc, Nf, src, k, i, j) += btop;
          }
        } else {
          const double *c = &CF(s->diff2diff, Nf * Nf, 0, k, i, j); /* diff2diff(src,dst) = c[dst*Nf+src] */
          double Ax = s->dy * A3(s->dz, k, i, j), Ay = s->dx * A3(s->dz, k, i, j), sum;
          btop = orc_B_eff(b1, b0, tauz);
          bbot = orc_B_eff(b0, b1, tauz);
          src = 0;
          bfac = ORC_PI * Az / (double)(Nft / 2);
          for (iside = 1; iside <= Nft; ++iside, ++src) {
            sum = 0; for (dst = 0; dst < Nf; ++dst) sum += c[dst * Nf + src];
            emis = dmax(0.0, dmin(1.0, 1.0 - sum));
            if (!s->st.ft_inward[iside - 1]) FG(xsrc, Nf, src, k, i, j) += btop * bfac * emis;
            else FG(xsrc, Nf, src, k + 1, i, j) += bbot * bfac * emis;
          }
          bfac = ORC_PI * Ax / (double)(Nfs / 2);
          for (iside = 1; iside <= Nfs; ++iside, ++src) {
            sum = 0; for (dst = 0; dst < Nf; ++dst) sum += c[dst * Nf + src];
            emis = dmax(0.0, dmin(1.0, 1.0 - sum));
            emis = (iside > Nfs / 2) ? btop * emis : bbot * emis;
            if (!s->st.fs_inward[iside - 1]) FG(xsrc, Nf, src, k, i, j) += emis * bfac;
            else FG(xsrc, Nf, src, k, i + 1, j) += emis * bfac;
          }
          bfac = ORC_PI * Ay / (double)(Nfs / 2);
          for (iside = 1; iside <= Nfs; ++iside, ++src) {
            sum = 0; for (dst = 0; dst < Nf; ++dst) sum += c[dst * Nf + src];
            emis = dmax(0.0, dmin(1.0, 1.0 - sum));
            emis = (iside > Nfs / 2) ? btop * emis : bbot * emis;
            if (!s->st.fs_inward[iside - 1]) FG(xsrc, Nf, src, k, i, j) += emis * bfac;
            else FG(xsrc, Nf, src, k, i, j + 1) += emis * bfac;
          }
        }
      }
      /* thermal emission at the surface :4958-4985 */
      for (src = 0; src < Nft; ++src) if (!s->st.ft_inward[src]) {
        if (s->have_Bsrfc) {
          double emis = dmax(0.0, dmin(1.0, 1.0 - A2(s->albedo, i, j)));
          FG(xsrc, Nf, src, Nz, i, j) += A2(s->Bsrfc, i, j) * Az * emis * ORC_PI / (double)(Nft / 2);
        } else {
          FG(xsrc, Nf, src, Nz, i, j) += P3(s->planck, Nz, i, j) * Az * (1.0 - A2(s->albedo, i, j)) * ORC_PI / (double)(Nft / 2);
        }
      }
    }
  }
  if (s->nbld) { /* :4669-4672 */
    int m, face;
    if (sol->lsolar_rad) { /* set_buildings_reflection :4989-5080: the direct beam hitting a building face leaves as
                              albedo/streams in every diffuse stream leaving through that face */
      for (m = 0; m < s->nbld; ++m) {
        const double ag = s->bld_albedo[m];
        int kk, ii, jj, off, out_is_inward, s0, ns;
        bld_idx(s, m, &face, &k, &i, &j);
        kk = k; ii = i; jj = j;
        if (face == ORC_BOT_FACE) kk = k + 1;
        if (face == ORC_RIGHT_FACE) ii = i + 1;
        if (face == ORC_FRONT_FACE) jj = j + 1;
        if (face == ORC_TOP_FACE || face == ORC_BOT_FACE) {
          out_is_inward = face == ORC_BOT_FACE;
          for (idiff = 0; idiff < Nft; ++idiff) if (s->st.ft_inward[idiff] == out_is_inward) FG(xsrc, Nf, idiff, kk, ii, jj) = 0.0;
          for (src = 0; src < Ndt; ++src) for (idiff = 0; idiff < Nft; ++idiff) if (s->st.ft_inward[idiff] == out_is_inward)
            FG(xsrc, Nf, idiff, kk, ii, jj) += FG(xedir, Nd, src, kk, ii, jj) * ag / (double)(Nft / 2);
        } else {
          const int lx = face == ORC_LEFT_FACE || face == ORC_RIGHT_FACE;
          off = lx ? Nft : Nft + Nfs;
          s0 = lx ? Ndt : Ndt + Nds; ns = Nds;
          out_is_inward = face == ORC_RIGHT_FACE || face == ORC_FRONT_FACE;
          for (idiff = 0; idiff < Nfs; ++idiff) if (s->st.fs_inward[idiff] == out_is_inward) FG(xsrc, Nf, off + idiff, kk, ii, jj) = 0.0;
          for (src = s0; src < s0 + ns; ++src) for (idiff = 0; idiff < Nfs; ++idiff) if (s->st.fs_inward[idiff] == out_is_inward)
            FG(xsrc, Nf, off + idiff, kk, ii, jj) += FG(xedir, Nd, src, kk, ii, jj) * ag / (double)(Nfs / 2);
        }
      }
    }
    if (sol->lthermal_rad && s->bld_have_planck) { /* set_buildings_emission :5082-5145 */
      const double Az = s->dx * s->dy;
      for (m = 0; m < s->nbld; ++m) {
        double emis, Ax, Ay;
        int kk, ii, jj, off, out_is_inward;
        bld_idx(s, m, &face, &k, &i, &j);
        Ax = s->dy * A3(s->dz, k, i, j); Ay = s->dx * A3(s->dz, k, i, j); /* diffside%area_divider = 1 */
        emis = ORC_PI * s->bld_planck[m] * (1.0 - s->bld_albedo[m]);
        kk = k; ii = i; jj = j;
        if (face == ORC_BOT_FACE) kk = k + 1;
        if (face == ORC_RIGHT_FACE) ii = i + 1;
        if (face == ORC_FRONT_FACE) jj = j + 1;
        if (face == ORC_TOP_FACE || face == ORC_BOT_FACE) {
          out_is_inward = face == ORC_BOT_FACE;
          for (idiff = 0; idiff < Nft; ++idiff) if (s->st.ft_inward[idiff] == out_is_inward)
            FG(xsrc, Nf, idiff, kk, ii, jj) = Az * emis / (double)(Nft / 2);
        } else {
          const int lx = face == ORC_LEFT_FACE || face == ORC_RIGHT_FACE;
          off = lx ? Nft : Nft + Nfs;
          out_is_inward = face == ORC_RIGHT_FACE || face == ORC_FRONT_FACE;
          for (idiff = 0; idiff < Nfs; ++idiff) if (s->st.fs_inward[idiff] == out_is_inward)
            FG(xsrc, Nf, off + idiff, kk, ii, jj) = (lx ? Ax : Ay) * emis / (double)(Nfs / 2);
        }
      }
    }
  }
  free(xedir);
  halo_reduce_5pt(s, Nf, xsrc);
  ghosted_to_interior(s, Nf, xsrc, s->b);
  free(xsrc);
}

/* ---- diffuse operator and solvers ------------------------------------------------------------------------------ */
static int inv_dof(const orc_solver *s, int dof) { /* src/pprts_explicit.F90:1001-1014 */
  int inc = s->st.ft_inward[0] ? 1 : -1;
  return s->st.ft_inward[dof] ? dof + inc : dof - inc;
}

/* src/pprts_shell.F90:366-541 (op_mat_mult_ediff) */
void orc_mat_mult_ediff(orc_solver *s, const double *x, double *y) {
  LOCALS
  double *xx, *xb;
  long n, N = nin(s, Nf);
  int colour, jj;
  if (!s->mm_xx) { s->mm_xx = (double *)malloc(sizeof(double) * ngh(s, Nf)); s->mm_xb = (double *)malloc(sizeof(double) * ngh(s, Nf)); }
  xx = s->mm_xx; xb = s->mm_xb;
  memset(xb, 0, sizeof(double) * ngh(s, Nf));
  interior_to_ghosted(s, Nf, x, xx);
  halo_fill_5pt(s, Nf, xx); /* DMGlobalToLocal */
  /* The reference's cell loop scatter-adds into the cell's own faces and the faces i+1 / j+1 / k+1; every entry of xb
   * has exactly ONE producing cell, so the result does not depend on the order of the cells.  Rows j of one parity
   * never write the same entries (row j touches rows j and j+1 of the ghosted array), so the cell loop runs as two
   * OpenMP-parallel half sweeps over j -- bitwise identical to the serial loop. */
  for (colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
  for (jj = colour; jj < Ny; jj += 2) {
  int i;
  for (i = 0; i < Nx; ++i) {
    int j = jj, k, idst, isrc, dst, src, mdst, msrc;
    for (k = 0; k < Nz; ++k) {
      if (s->l1d[k]) {
        for (idst = 0; idst < Nft; ++idst) {
          if (s->st.ft_inward[idst]) {
            FG(xb, Nf, idst, k + 1, i, j) -= FG(xx, Nf, idst, k, i, j) * A3(s->a11, k, i, j);
            FG(xb, Nf, idst, k + 1, i, j) -= FG(xx, Nf, inv_dof(s, idst), k + 1, i, j) * A3(s->a12, k, i, j);
          } else {
            FG(xb, Nf, idst, k, i, j) -= FG(xx, Nf, idst, k + 1, i, j) * A3(s->a11, k, i, j);
            FG(xb, Nf, idst, k, i, j) -= FG(xx, Nf, inv_dof(s, idst), k, i, j) * A3(s->a12, k, i, j);
          }
        }
      } else {
        const double *v = &CF(s->diff2diff, Nf * Nf, 0, k, i, j);
        double *tgt;
        dst = 0;
        for (idst = 0; idst < Nft + 2 * Nfs; ++idst, ++dst) {
          if (idst < Nft) { mdst = s->st.ft_inward[idst] ? k + 1 : k; tgt = &FG(xb, Nf, dst, mdst, i, j); }
          else if (idst < Nft + Nfs) { mdst = s->st.fs_inward[idst - Nft] ? i + 1 : i; tgt = &FG(xb, Nf, dst, k, mdst, j); }
          else { mdst = s->st.fs_inward[idst - Nft - Nfs] ? j + 1 : j; tgt = &FG(xb, Nf, dst, k, i, mdst); }
          src = 0;
          for (isrc = 0; isrc < Nft; ++isrc, ++src) {
            msrc = s->st.ft_inward[isrc] ? k : k + 1;
            *tgt -= FG(xx, Nf, src, msrc, i, j) * v[dst * Nf + src];
          }
          for (isrc = 0; isrc < Nfs; ++isrc, ++src) {
            msrc = s->st.fs_inward[isrc] ? i : i + 1;
            *tgt -= FG(xx, Nf, src, k, msrc, j) * v[dst * Nf + src];
          }
          for (isrc = 0; isrc < Nfs; ++isrc, ++src) {
            msrc = s->st.fs_inward[isrc] ? j : j + 1;
            *tgt -= FG(xx, Nf, src, k, i, msrc) * v[dst * Nf + src];
          }
        }
      }
    }
    for (idst = 0; idst < Nft; ++idst) if (!s->st.ft_inward[idst]) /* albedo row :500-505 */
      FG(xb, Nf, idst, Nz, i, j) -= FG(xx, Nf, inv_dof(s, idst), Nz, i, j) * A2(s->albedo, i, j);
  }
  }
  }
  halo_reduce_5pt(s, Nf, xb); /* DMLocalToGlobal(ADD_VALUES) */
  ghosted_to_interior(s, Nf, xb, y);
#pragma omp parallel for schedule(static)
  for (n = 0; n < N; ++n) y[n] += x[n]; /* VecAXPY: diagonal */
}

/* src/pprts_explicit.F90:715-848 (exchange_diffuse_boundary), one periodic rank */
static void exchange_diffuse_boundary(const orc_solver *s, double *v0) {
  LOCALS
  int k, i, j, idof, dof;
  for (j = 0; j < Ny; ++j) for (k = 0; k < Nz1; ++k) for (idof = 0; idof < Nfs; ++idof) {
    dof = Nft + idof;
    if (s->st.fs_inward[idof]) FG(v0, Nf, dof, k, 0, j) = FG(v0, Nf, dof, k, Nx, j);   /* send_e(xe+1) -> neighbour xs */
    else                       FG(v0, Nf, dof, k, Nx, j) = FG(v0, Nf, dof, k, 0, j);   /* send_w(xs) -> neighbour xe+1 */
  }
  for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) for (idof = 0; idof < Nfs; ++idof) {
    dof = Nft + Nfs + idof;
    if (s->st.fs_inward[idof]) FG(v0, Nf, dof, k, i, 0) = FG(v0, Nf, dof, k, i, Ny);
    else                       FG(v0, Nf, dof, k, i, Ny) = FG(v0, Nf, dof, k, i, 0);
  }
}

/* src/pprts_explicit.F90:849-1015 (explicit_ediff_sor_sweep) */
static void explicit_ediff_sor_sweep(const orc_solver *s, const int dxr[3], const int dyr[3], const int dzr[3],
                                     double omega, const double *xb, double *x0) {
  LOCALS
  int i, j, k, idst, isrc, src, dst, msrc, mdst;
  double sigma;
  if (dzr[2] < 0) {
    for (j = dyr[0]; j != dyr[1] + dyr[2]; j += dyr[2]) for (i = dxr[0]; i != dxr[1] + dxr[2]; i += dxr[2])
      for (idst = 0; idst < Nft; ++idst) if (!s->st.ft_inward[idst])
        FG(x0, Nf, idst, Nz, i, j) = FG(xb, Nf, idst, Nz, i, j) + FG(x0, Nf, inv_dof(s, idst), Nz, i, j) * A2(s->albedo, i, j);
  }
  for (j = dyr[0]; j != dyr[1] + dyr[2]; j += dyr[2]) for (i = dxr[0]; i != dxr[1] + dxr[2]; i += dxr[2])
    for (k = dzr[0]; k != dzr[1] + dzr[2]; k += dzr[2]) {
      if (s->l1d[k]) {
        for (idst = 0; idst < Nft; ++idst) {
          if (s->st.ft_inward[idst])
            FG(x0, Nf, idst, k + 1, i, j) = FG(xb, Nf, idst, k + 1, i, j) + FG(x0, Nf, idst, k, i, j) * A3(s->a11, k, i, j) +
                                            FG(x0, Nf, inv_dof(s, idst), k + 1, i, j) * A3(s->a12, k, i, j);
          else
            FG(x0, Nf, idst, k, i, j) = FG(xb, Nf, idst, k, i, j) + FG(x0, Nf, idst, k + 1, i, j) * A3(s->a11, k, i, j) +
                                        FG(x0, Nf, inv_dof(s, idst), k, i, j) * A3(s->a12, k, i, j);
        }
      } else {
        const double *v = &CF(s->diff2diff, Nf * Nf, 0, k, i, j);
        dst = 0;
        for (idst = 0; idst < Nft + 2 * Nfs; ++idst, ++dst) {
          double *px; const double *pb;
          if (idst < Nft) { mdst = s->st.ft_inward[idst] ? k + 1 : k; px = &FG(x0, Nf, dst, mdst, i, j); pb = &FG(xb, Nf, dst, mdst, i, j); }
          else if (idst < Nft + Nfs) { mdst = s->st.fs_inward[idst - Nft] ? i + 1 : i; px = &FG(x0, Nf, dst, k, mdst, j); pb = &FG(xb, Nf, dst, k, mdst, j); }
          else { mdst = s->st.fs_inward[idst - Nft - Nfs] ? j + 1 : j; px = &FG(x0, Nf, dst, k, i, mdst); pb = &FG(xb, Nf, dst, k, i, mdst); }
          sigma = 0; src = 0;
          for (isrc = 0; isrc < Nft; ++isrc, ++src) { msrc = s->st.ft_inward[isrc] ? k : k + 1; sigma = sigma + FG(x0, Nf, src, msrc, i, j) * v[dst * Nf + src]; }
          for (isrc = 0; isrc < Nfs; ++isrc, ++src) { msrc = s->st.fs_inward[isrc] ? i : i + 1; sigma = sigma + FG(x0, Nf, src, k, msrc, j) * v[dst * Nf + src]; }
          for (isrc = 0; isrc < Nfs; ++isrc, ++src) { msrc = s->st.fs_inward[isrc] ? j : j + 1; sigma = sigma + FG(x0, Nf, src, k, i, msrc) * v[dst * Nf + src]; }
          *px = (1.0 - omega) * (*px) + omega * (*pb + sigma);
        }
      }
    }
  if (dzr[2] > 0) {
    for (j = dyr[0]; j != dyr[1] + dyr[2]; j += dyr[2]) for (i = dxr[0]; i != dxr[1] + dxr[2]; i += dxr[2])
      for (idst = 0; idst < Nft; ++idst) if (!s->st.ft_inward[idst])
        FG(x0, Nf, idst, Nz, i, j) = FG(xb, Nf, idst, Nz, i, j) + FG(x0, Nf, inv_dof(s, idst), Nz, i, j) * A2(s->albedo, i, j);
  }
}

/* src/pprts_explicit.F90:461-713 (explicit_ediff): alternating-direction SOR with adaptive omega */
static int explicit_ediff(orc_solver *s, orc_solution *sol) {
  LOCALS
  int maxiter = s->cfg.maxit_diff > 0 ? s->cfg.maxit_diff : 10000, iter, converged = 0;
  double rtol_d, atol_d, rtol, atol, rel;
  double omega_adaptive = 1.0, omega_min = 1.0, omega_max = 1.25, omega_dir = 1.0, omega_step = 0.1 * 0.5;
  double omega_save, log_rate, log_rate_prev = 0.0, best = DBL_MAX;
  int iter_at_best = 1, frozen = 0;
  const int stagnation_window = 50;
  double *residual, *v0, *lvb;
  determine_ksp_tolerances(s, &rtol_d, &atol_d);
  rtol = s->cfg.rtol >= 0 ? s->cfg.rtol : rtol_d;
  atol = s->cfg.atol >= 0 ? s->cfg.atol : atol_d;
  if (sol->diff_sor_omega > 0.0) omega_adaptive = sol->diff_sor_omega;
  omega_save = omega_adaptive;
  if (sol->diff_hist[0] < 0) { maxiter = maxiter + 2; rtol = dmin(rtol, rtol_d); atol = dmin(atol, atol_d); }
  residual = (double *)calloc(maxiter + 1, sizeof(double));
  v0 = (double *)calloc(ngh(s, Nf), sizeof(double));
  lvb = (double *)calloc(ngh(s, Nf), sizeof(double));
  interior_to_ghosted(s, Nf, sol->ediff, v0);
  fill_ghost(s, Nf, v0);
  interior_to_ghosted(s, Nf, s->b, lvb);
  fill_ghost(s, Nf, lvb);
  sol->Niter_diff = 0;
  for (iter = 1; iter <= maxiter; ++iter) {
    if ((iter + 1) % 2 == 0) {
      int dxr[3] = {0, Nx - 1, 1}, dyr[3] = {0, Ny - 1, 1}, dzr[3] = {0, Nz - 1, 1};
      explicit_ediff_sor_sweep(s, dxr, dyr, dzr, omega_adaptive, lvb, v0);
    } else {
      int dxr[3] = {Nx - 1, 0, -1}, dyr[3] = {Ny - 1, 0, -1}, dzr[3] = {Nz - 1, 0, -1};
      explicit_ediff_sor_sweep(s, dxr, dyr, dzr, omega_adaptive, lvb, v0);
    }
    exchange_diffuse_boundary(s, v0);
    residual[iter] = norm2_diff_and_copy(s, Nf, sol->ediff, v0);
    rel = (residual[1] <= sqrt(DBL_MIN)) ? 0.0 : residual[iter] / residual[1];
    sol->diff_hist[(iter < ORC_NHIST ? iter : ORC_NHIST) - 1] = residual[iter];
    sol->Niter_diff = iter;
    sol->final_res_diff = residual[iter];
    if (residual[iter] < atol || rel < rtol) { converged = 1; break; }
    if (residual[iter] < best) { best = residual[iter]; iter_at_best = iter; }
    if (iter >= 3) {
      if (!frozen && omega_adaptive > omega_min && (iter - iter_at_best) > stagnation_window) {
        omega_adaptive = omega_min; omega_save = omega_adaptive; frozen = 1;
      }
      if (!frozen && residual[iter] > 0.0 && residual[iter - 2] > 0.0) {
        log_rate = 0.5 * log(residual[iter] / residual[iter - 2]);
        if (log_rate < log_rate_prev) omega_step = dmin(omega_step * 1.3, omega_max - omega_min);
        else { omega_dir = -omega_dir; omega_step = dmax(omega_step * 0.5, 0.01); }
        log_rate_prev = log_rate;
        omega_adaptive = dmin(dmax(omega_adaptive + omega_dir * omega_step, omega_min), omega_max);
        omega_save = omega_adaptive;
      }
    }
  }
  sol->diff_sor_omega = omega_save;
  ghosted_to_interior(s, Nf, v0, sol->ediff);
  sol->lchanged = 1;
  sol->lWm2_diff = 0;
  free(residual); free(v0); free(lvb);
  return converged ? 0 : 1;
}

/* Unpreconditioned BiCGStab on the shell operator with the convergence test of MyKSPConverged
 * (src/pprts.F90:4437-4486: r/r0 <= rtol or r <= atol, checked from iteration 1; nonzero initial guess :4343).
 * PETSc's KSPFBCGS+ILU itself is third-party and not restated; any converged Krylov method yields the same x. */
/* blocked sum: fixed blocks of 65536 entries added in index order, so the value does not depend on the thread count */
static double dot_(const double *a, const double *b, long n) {
  const long B = 65536, nb = (n + B - 1) / B;
  double *part = (double *)malloc(sizeof(double) * (nb > 0 ? nb : 1)), s = 0;
  long ib;
#pragma omp parallel for schedule(static)
  for (ib = 0; ib < nb; ++ib) {
    long i, e = (ib + 1) * B < n ? (ib + 1) * B : n;
    double acc = 0;
    for (i = ib * B; i < e; ++i) acc += a[i] * b[i];
    part[ib] = acc;
  }
  for (ib = 0; ib < nb; ++ib) s += part[ib];
  free(part);
  return s;
}
static int bicgstab_ediff(orc_solver *s, orc_solution *sol) {
  long n = nin(s, s->st.Nf), i;
  int maxit = s->cfg.maxit_diff > 0 ? s->cfg.maxit_diff : 1000, it, converged = 0;
  double rtol_d, atol_d, rtol, atol, rho = 1, alpha = 1, omega = 1, rho1, beta, r0n, rn, tt;
  double *x = sol->ediff, *r = (double *)malloc(sizeof(double) * n), *rh = (double *)malloc(sizeof(double) * n);
  double *p = (double *)calloc(n, sizeof(double)), *v = (double *)calloc(n, sizeof(double));
  double *sv = (double *)malloc(sizeof(double) * n), *t = (double *)malloc(sizeof(double) * n);
  determine_ksp_tolerances(s, &rtol_d, &atol_d);
  rtol = s->cfg.rtol >= 0 ? s->cfg.rtol : rtol_d;
  atol = s->cfg.atol >= 0 ? s->cfg.atol : atol_d;
  if (sol->diff_hist[0] < 0) { rtol = dmin(rtol, rtol_d); atol = dmin(atol, atol_d); }
  orc_mat_mult_ediff(s, x, r);
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) { r[i] = s->b[i] - r[i]; rh[i] = r[i]; }
  r0n = dmax(DBL_MIN, sqrt(dot_(r, r, n)));
  sol->diff_hist[0] = r0n;
  sol->Niter_diff = 0;
  for (it = 1; it <= maxit; ++it) {
    rho1 = dot_(rh, r, n);
    beta = (rho1 / rho) * (alpha / omega);
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    orc_mat_mult_ediff(s, p, v);
    alpha = rho1 / dot_(rh, v, n);
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) sv[i] = r[i] - alpha * v[i];
    orc_mat_mult_ediff(s, sv, t);
    tt = dot_(t, t, n);
    omega = tt > 0 ? dot_(t, sv, n) / tt : 0.0;
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) { x[i] += alpha * p[i] + omega * sv[i]; r[i] = sv[i] - omega * t[i]; }
    rho = rho1;
    rn = sqrt(dot_(r, r, n));
    sol->diff_hist[(it < ORC_NHIST ? it : ORC_NHIST - 1)] = rn;
    sol->Niter_diff = it;
    sol->final_res_diff = rn;
    if (rn / r0n <= rtol || rn <= atol) { converged = 1; break; }
    if (omega == 0.0 || isnan(rn)) break;
  }
  sol->lchanged = 1;
  sol->lWm2_diff = 0;
  free(r); free(rh); free(p); free(v); free(sv); free(t);
  return converged ? 0 : 1;
}

/* ---- absorption (src/pprts.F90:5152-5504) ----------------------------------------------------------------------- */
static void calc_flx_div(orc_solver *s, orc_solution *sol) {
  LOCALS
  int i, j, k, isrc, idof, msrc, src, dst;
  double cdiv, sum;
  double *ledir = 0, *lediff = (double *)calloc(ngh(s, Nf), sizeof(double));
  const int xinc = s->xinc, yinc = s->yinc;
  memset(sol->abso, 0, sizeof(double) * (long)Nz * Nx * Ny);
  scale_flx(s, sol, 0);
  if (sol->lsolar_rad) {
    ledir = (double *)calloc(ngh(s, Nd), sizeof(double));
    interior_to_ghosted(s, Nd, sol->edir, ledir);
    halo_fill_5pt(s, Nd, ledir);
  }
  interior_to_ghosted(s, Nf, sol->ediff, lediff);
  halo_fill_5pt(s, Nf, lediff);
  if (s->cfg.by_coeff_divergence) {
    if (sol->lsolar_rad) {
      for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
        if (s->l1d[k]) {
          for (isrc = 0; isrc < Ndt; ++isrc) {
            cdiv = A3(s->kabs, k, i, j) * A3(s->dz, k, i, j) / s->costheta;
            A3(sol->abso, k, i, j) += FG(ledir, Nd, isrc, k, i, j) * (-expm1(-cdiv));
          }
        } else {
          const double *c1 = &CF(s->dir2dir, Nd * Nd, 0, k, i, j), *c2 = &CF(s->dir2diff, Nd * Nf, 0, k, i, j);
          idof = 0;
          for (isrc = 0; isrc < Ndt + 2 * Nds; ++isrc, ++idof) {
            double sum1 = 0, sum2 = 0, e;
            for (dst = 0; dst < Nd; ++dst) sum1 += c1[dst * Nd + idof];
            for (dst = 0; dst < Nf; ++dst) sum2 += c2[dst * Nd + idof];
            cdiv = 1.0 - sum1 - sum2;
            if (isrc < Ndt) e = FG(ledir, Nd, idof, k, i, j);
            else if (isrc < Ndt + Nds) e = FG(ledir, Nd, idof, k, i + 1 - xinc, j);
            else e = FG(ledir, Nd, idof, k, i, j + 1 - yinc);
            A3(sol->abso, k, i, j) += e * cdiv;
          }
        }
      }
    }
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
      if (!s->l1d[k]) {
        const double *c = &CF(s->diff2diff, Nf * Nf, 0, k, i, j);
        idof = 0;
        for (isrc = 0; isrc < Nft + 2 * Nfs; ++isrc, ++idof) {
          double e;
          sum = 0; for (dst = 0; dst < Nf; ++dst) sum += c[dst * Nf + idof];
          cdiv = 1.0 - sum;
          if (isrc < Nft) { msrc = s->st.ft_inward[isrc] ? k : k + 1; e = FG(lediff, Nf, idof, msrc, i, j); }
          else if (isrc < Nft + Nfs) { msrc = s->st.fs_inward[isrc - Nft] ? i : i + 1; e = FG(lediff, Nf, idof, k, msrc, j); }
          else { msrc = s->st.fs_inward[isrc - Nft - Nfs] ? j : j + 1; e = FG(lediff, Nf, idof, k, i, msrc); }
          A3(sol->abso, k, i, j) += e * cdiv;
        }
      } else {
        cdiv = dmax(0.0, 1.0 - A3(s->a11, k, i, j) - A3(s->a12, k, i, j));
        for (isrc = 0; isrc < Nft; ++isrc) {
          msrc = s->st.ft_inward[isrc] ? k : k + 1;
          A3(sol->abso, k, i, j) += FG(lediff, Nf, isrc, msrc, i, j) * cdiv;
        }
      }
    }
    if (sol->lthermal_rad) {
      double *lsrc = (double *)calloc(ngh(s, Nf), sizeof(double));
      interior_to_ghosted(s, Nf, s->b, lsrc);
      halo_fill_5pt(s, Nf, lsrc);
      for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
        idof = 0;
        for (isrc = 0; isrc < Nft; ++isrc, ++idof) { msrc = s->st.ft_inward[isrc] ? k + 1 : k; A3(sol->abso, k, i, j) -= FG(lsrc, Nf, idof, msrc, i, j); }
        for (isrc = 0; isrc < Nfs; ++isrc, ++idof) { msrc = s->st.fs_inward[isrc] ? i + 1 : i; A3(sol->abso, k, i, j) -= FG(lsrc, Nf, idof, k, msrc, j); }
        for (isrc = 0; isrc < Nfs; ++isrc, ++idof) { msrc = s->st.fs_inward[isrc] ? j + 1 : j; A3(sol->abso, k, i, j) -= FG(lsrc, Nf, idof, k, i, msrc); }
      }
      free(lsrc);
    }
  } else { /* by flux divergence :5400-5472 */
    if (sol->lsolar_rad) {
      for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
        if (s->l1d[k]) {
          for (isrc = 0; isrc < Ndt; ++isrc) A3(sol->abso, k, i, j) += FG(ledir, Nd, isrc, k, i, j) - FG(ledir, Nd, isrc, k + 1, i, j);
        } else {
          src = 0;
          for (isrc = 0; isrc < Ndt; ++isrc, ++src) A3(sol->abso, k, i, j) += FG(ledir, Nd, src, k, i, j) - FG(ledir, Nd, src, k + 1, i, j);
          for (isrc = 0; isrc < Nds; ++isrc, ++src) A3(sol->abso, k, i, j) += FG(ledir, Nd, src, k, i + 1 - xinc, j) - FG(ledir, Nd, src, k, i + xinc, j);
          for (isrc = 0; isrc < Nds; ++isrc, ++src) A3(sol->abso, k, i, j) += FG(ledir, Nd, src, k, i, j + 1 - yinc) - FG(ledir, Nd, src, k, i, j + yinc);
        }
      }
    }
    for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k) {
      src = 0;
      for (isrc = 0; isrc < Nft; ++isrc, ++src) {
        if (s->st.ft_inward[isrc]) A3(sol->abso, k, i, j) += FG(lediff, Nf, src, k, i, j) - FG(lediff, Nf, src, k + 1, i, j);
        else A3(sol->abso, k, i, j) += FG(lediff, Nf, src, k + 1, i, j) - FG(lediff, Nf, src, k, i, j);
      }
      if (s->l1d[k]) continue;
      for (isrc = 0; isrc < Nfs; ++isrc, ++src) {
        if (s->st.fs_inward[isrc]) A3(sol->abso, k, i, j) += FG(lediff, Nf, src, k, i, j) - FG(lediff, Nf, src, k, i + 1, j);
        else A3(sol->abso, k, i, j) += FG(lediff, Nf, src, k, i + 1, j) - FG(lediff, Nf, src, k, i, j);
      }
      for (isrc = 0; isrc < Nfs; ++isrc, ++src) {
        if (s->st.fs_inward[isrc]) A3(sol->abso, k, i, j) += FG(lediff, Nf, src, k, i, j) - FG(lediff, Nf, src, k, i, j + 1);
        else A3(sol->abso, k, i, j) += FG(lediff, Nf, src, k, i, j + 1) - FG(lediff, Nf, src, k, i, j);
      }
    }
  }
  /* abso_scalevec :5477, :5483-5503 */
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz; ++k)
    A3(sol->abso, k, i, j) = A3(sol->abso, k, i, j) * (1.0 / (s->dx * s->dy * A3(s->dz, k, i, j)));
  free(ledir); free(lediff);
}

/* ---- solve_pprts / pprts / restore_solution (src/pprts.F90:2487-3033, :3991-4076) ---------------------------------- */
static orc_solution *prepare_solution(orc_solver *s, int lsolar, int lthermal) { /* src/pprts_base.F90:298-344 */
  orc_solution *sol = (orc_solution *)calloc(1, sizeof(orc_solution));
  int h;
  sol->lset = 1; sol->lsolar_rad = lsolar; sol->lthermal_rad = lthermal; sol->lchanged = 1;
  sol->lWm2_dir = 1; sol->lWm2_diff = 1;
  sol->edir = (double *)calloc(nin(s, s->st.Nd), sizeof(double));
  sol->ediff = (double *)calloc(nin(s, s->st.Nf), sizeof(double));
  sol->abso = (double *)calloc((long)s->Nz * s->Nx * s->Ny, sizeof(double));
  for (h = 0; h < ORC_NHIST; ++h) { sol->dir_hist[h] = -1; sol->diff_hist[h] = -1; }
  for (h = 0; h < ORC_NTIME; ++h) { sol->time[h] = -1.0; sol->maxnorm[h] = 0.0; }
  sol->diff_sor_omega = -1;
  return sol;
}

/* ---- error-growth tracking (src/adaptive_spectral_integration.F90, restore_solution src/pprts.F90:3991-4049) ------ */
/* polyfit :174-257: least squares through the normal equations, (X^T X)^-1 by LU with partial pivoting (LAPACK
 * dgetrf + dgetri there), coefficients = ((X^T X)^-1 X^T) y.  ierr 1: two equal abscissae, 2: singular. */
int orc_polyfit(const double *vx, const double *vy, int n, int d, double *coef) {
  double X[16][4], XTX[4][4], inv[4][4], M[4][16];
  int i, j, k, m = d + 1, p;
  if (n > 16 || m > 4) return 9;
  for (i = 0; i < n; ++i) for (j = i + 1; j < n; ++j) if (approx_d(vx[i], vx[j])) {
    for (k = 0; k < m; ++k) coef[k] = 0.0;
    coef[0] = -9999999.0; /* nil */
    return 1;
  }
  for (i = 0; i < m; ++i) for (j = 0; j < n; ++j) X[j][i] = pow(vx[j], (double)i);
  for (i = 0; i < m; ++i) for (j = 0; j < m; ++j) {
    double acc = 0; for (k = 0; k < n; ++k) acc += X[k][i] * X[k][j];
    XTX[i][j] = acc; inv[i][j] = (i == j) ? 1.0 : 0.0;
  }
  for (p = 0; p < m; ++p) { /* Gauss-Jordan with partial pivoting */
    int piv = p; double best = fabs(XTX[p][p]), ip;
    for (i = p + 1; i < m; ++i) if (fabs(XTX[i][p]) > best) { best = fabs(XTX[i][p]); piv = i; }
    if (best == 0.0) return 2;
    if (piv != p) for (j = 0; j < m; ++j) {
      double t = XTX[p][j]; XTX[p][j] = XTX[piv][j]; XTX[piv][j] = t;
      t = inv[p][j]; inv[p][j] = inv[piv][j]; inv[piv][j] = t;
    }
    ip = 1.0 / XTX[p][p];
    for (j = 0; j < m; ++j) { XTX[p][j] *= ip; inv[p][j] *= ip; }
    for (i = 0; i < m; ++i) if (i != p) {
      double f = XTX[i][p];
      for (j = 0; j < m; ++j) { XTX[i][j] -= f * XTX[p][j]; inv[i][j] -= f * inv[p][j]; }
    }
  }
  for (i = 0; i < m; ++i) for (j = 0; j < n; ++j) {
    double acc = 0; for (k = 0; k < m; ++k) acc += inv[i][k] * X[j][k];
    M[i][j] = acc;
  }
  for (i = 0; i < m; ++i) { double acc = 0; for (j = 0; j < n; ++j) acc += M[i][j] * vy[j]; coef[i] = acc; }
  return 0;
}
/* need_new_solution :38-172.  have_time = 0: `time` not present.  Returns 1 if the band has to be recomputed. */
int orc_need_new_solution(orc_solver *s, int uid, double time, int have_time) {
  enum { Nfit = 3, Nporder = 2 };
  const double max_err = s->cfg.max_solution_err, max_time = s->cfg.max_solution_time;
  const int lenable = max_err > 0.0 && max_time > 0.0; /* src/pprts.F90:456-457 */
  orc_solution *sol;
  double t[Nfit], tm[Nfit], err[2][2 * (Nfit - 1)], polyc[Nporder + 1], estimate[Nporder], error_estimate = 0.0;
  int k, ipoly, need;
  if (uid < ORC_UID_MIN || uid > ORC_UID_MAX) return 1;
  if (!have_time) return 1;                      /* 'time not provided' */
  if (!lenable) return 1;                        /* 'solutions_err_estimates disabled' */
  sol = s->sol[uid - ORC_UID_MIN];
  if (!sol || !sol->lset) return 1;              /* 'no solution yet' */
  for (k = 1; k <= Nfit; ++k) t[k - 1] = sol->time[Nfit - k];        /* t(k) = time(Nfit-k+1) */
  for (k = 1; k <= Nfit - 1; ++k) tm[k - 1] = (t[k - 1] + t[k]) * .5;
  tm[Nfit - 1] = (t[Nfit - 1] + time) * .5;
  for (k = 1; k <= Nfit - 1; ++k) {
    err[0][2 * k - 2] = tm[k - 1]; err[0][2 * k - 1] = t[k];
    err[1][2 * k - 2] = 0.0;       err[1][2 * k - 1] = sol->maxnorm[Nfit - k - 1];  /* maxnorm(Nfit-k) */
  }
  for (ipoly = 1; ipoly <= Nporder; ++ipoly) {
    double tp = 1.0;
    if (orc_polyfit(err[0], err[1], 2 * (Nfit - 1), ipoly, polyc)) return 1; /* 'problem fitting error curve' */
    estimate[ipoly - 1] = 0.0;
    for (k = 1; k <= ipoly + 1; ++k) { estimate[ipoly - 1] += polyc[k - 1] * tp; tp *= time; }
  }
  for (k = 0; k < Nporder; ++k) error_estimate = dmax(error_estimate, fabs(estimate[k]));
  need = !(error_estimate <= max_err);           /* ERR_TOL_IN_BOUND / ERR_TOL_EXCEEDED */
  for (k = 0; k < Nfit; ++k) if (t[k] < 0.0) need = 1;   /* FEW_SOLUTIONS */
  if (time - sol->time[0] > max_time) need = 1;          /* MIN_TIME_EXCEEDED */
  return need;
}
void orc_get_solution_times(orc_solver *s, int uid, double *time, double *maxnorm) {
  orc_solution *sol = s->sol[uid - ORC_UID_MIN];
  int k;
  for (k = 0; k < ORC_NTIME; ++k) { time[k] = sol ? sol->time[k] : -1.0; maxnorm[k] = sol ? sol->maxnorm[k] : 0.0; }
}

static int solve_impl(orc_solver *s, int lthermal, int lsolar, double edirTOA, int uid, double time, int have_time);
int orc_solve(orc_solver *s, int lthermal, int lsolar, double edirTOA, int uid) {
  return solve_impl(s, lthermal, lsolar, edirTOA, uid, 0.0, 0);
}
/* solve_pprts(..., opt_solution_time = time) */
int orc_solve_t(orc_solver *s, int lthermal, int lsolar, double edirTOA, int uid, double time) {
  return solve_impl(s, lthermal, lsolar, edirTOA, uid, time, 1);
}
static int solve_impl(orc_solver *s, int lthermal, int lsolar, double edirTOA, int uid, double time, int have_time) {
  orc_solution *sol;
  int derived_lsolar, rc = 0;
  long n, N;
  double b_norm, rtol, atol;
  if (!s->have_optprop) return 10;
  if (uid < ORC_UID_MIN || uid > ORC_UID_MAX) return 11;
  if (lthermal && !s->have_planck) return 12;
  derived_lsolar = lsolar && edirTOA > 0.0 && s->theta >= 0.0;
  if (derived_lsolar && lthermal) return 13; /* src/pprts.F90:2523-2530 */
  sol = s->sol[uid - ORC_UID_MIN];
  if (!sol) {
    sol = s->sol[uid - ORC_UID_MIN] = prepare_solution(s, derived_lsolar, lthermal);
    /* -initial_guess_from_last_uid (default on), src/pprts.F90:2542-2558 */
    if (uid - 1 >= ORC_UID_MIN && s->sol[uid - 1 - ORC_UID_MIN] && s->sol[uid - 1 - ORC_UID_MIN]->lset) {
      orc_solution *last = s->sol[uid - 1 - ORC_UID_MIN];
      if (last->lsolar_rad) { memcpy(sol->edir, last->edir, sizeof(double) * nin(s, s->st.Nd)); sol->lWm2_dir = last->lWm2_dir; }
      memcpy(sol->ediff, last->ediff, sizeof(double) * nin(s, s->st.Nf));
      sol->lWm2_diff = last->lWm2_diff;
    }
  } else if (sol->lsolar_rad != derived_lsolar) {
    free_solution(sol);
    sol = s->sol[uid - ORC_UID_MIN] = prepare_solution(s, derived_lsolar, lthermal);
  }
  if (!sol->lthermal_rad && !sol->lsolar_rad) { /* :2590-2596 */
    memset(sol->edir, 0, sizeof(double) * nin(s, s->st.Nd));
    memset(sol->ediff, 0, sizeof(double) * nin(s, s->st.Nf));
    memset(sol->abso, 0, sizeof(double) * (long)s->Nz * s->Nx * s->Ny);
    sol->lchanged = 0;
    return 0;
  }
  /* pprts(), :2668-3033 */
  { int ec = alloc_coeffs(s, sol->lsolar_rad); if (ec) return 20 + ec; } /* 21: geometric coeffs need 3 direct streams, 22: extinction too small */
  scale_flx(s, sol, 0);
  if (sol->lsolar_rad) rc |= explicit_edir(s, edirTOA, sol);
  setup_b(s, sol);
  determine_ksp_tolerances(s, &rtol, &atol);
  N = nin(s, s->st.Nf);
  b_norm = 0; for (n = 0; n < N; ++n) b_norm += fabs(s->b[n]);
  if (b_norm < atol) { /* :2776-2784 */
    memcpy(sol->ediff, s->b, sizeof(double) * N);
    sol->Niter_diff = 0;
    for (n = 0; n < ORC_NHIST; ++n) sol->diff_hist[n] = atol;
  } else {
    if (s->cfg.diff_solver == 0) rc |= explicit_ediff(s, sol) << 1;
    else rc |= bicgstab_ediff(s, sol) << 1;
  }
  sol->lchanged = 1;
  sol->lWm2_diff = 0;
  /* restore_solution :3991-4049 */
  {
    const int lenable = s->cfg.max_solution_err > 0.0 && s->cfg.max_solution_time > 0.0;
    const long nc = (long)s->Nz * s->Nx * s->Ny;
    double *abso_old = 0;
    int k;
    if (have_time) { /* eoshift(time, -1); time(1) = time */
      for (k = ORC_NTIME - 1; k > 0; --k) sol->time[k] = sol->time[k - 1];
      sol->time[0] = time;
    }
    if (have_time && lenable) { abso_old = (double *)malloc(sizeof(double) * nc); memcpy(abso_old, sol->abso, sizeof(double) * nc); }
    calc_flx_div(s, sol);
    scale_flx(s, sol, 1);
    sol->lchanged = 0;
    if (abso_old) { /* :4036-4049 inf-norm of the change of the absorption */
      double inf_norm = 0.0;
      for (n = 0; n < nc; ++n) inf_norm = dmax(inf_norm, fabs(abso_old[n] - sol->abso[n]));
      for (k = ORC_NTIME - 1; k > 0; --k) sol->maxnorm[k] = sol->maxnorm[k - 1];
      sol->maxnorm[0] = inf_norm;
      free(abso_old);
    }
  }
  return rc;
}

/* src/pprts.F90:5799-5917 (pprts_get_result) */
int orc_get_result(orc_solver *s, int uid, double *edn, double *eup, double *abso, double *edir) {
  LOCALS
  orc_solution *sol;
  int k, i, j, d;
  if (uid < ORC_UID_MIN || uid > ORC_UID_MAX || !(sol = s->sol[uid - ORC_UID_MIN])) return 1;
  if (sol->lchanged) return 2;
  if (!sol->lWm2_diff) return 3;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) {
    for (k = 0; k < Nz1; ++k) {
      double dn = 0, up = 0, dr = 0;
      for (d = 0; d < Nft; ++d) {
        if (s->st.ft_inward[d]) dn = dn + FI(sol->ediff, Nf, d, k, i, j) / 1.0;
        else up = up + FI(sol->ediff, Nf, d, k, i, j) / 1.0;
      }
      if (sol->lsolar_rad) {
        for (d = 0; d < Ndt; ++d) dr += FI(sol->edir, Nd, d, k, i, j);
        dr = dr / (double)s->st.dirtop_adiv;
        dr *= s->mu; dn *= s->mu; up *= s->mu;
      }
      P3(edn, k, i, j) = dn; P3(eup, k, i, j) = up;
      if (edir) P3(edir, k, i, j) = dr;
    }
    for (k = 0; k < Nz; ++k) A3(abso, k, i, j) = A3(sol->abso, k, i, j) * (sol->lsolar_rad ? s->mu : 1.0);
  }
  return 0;
}

/* pprts_get_result::fill_buildings_arr (src/pprts.F90:6011-6160): per face the direct flux hitting it (edir, NULL if
 * not wanted / not solar), the diffuse flux entering the building through it (incoming) and leaving it (outgoing),
 * W/m2.  The -pprts_fill_1D_side_walls branch only concerns solvers without side streams and is not restated. */
int orc_get_buildings_result(orc_solver *s, int uid, double *edir, double *incoming, double *outgoing) {
  LOCALS
  orc_solution *sol;
  double *ledir = 0, *lediff;
  int m, face, k, i, j, d;
  if (uid < ORC_UID_MIN || uid > ORC_UID_MAX || !(sol = s->sol[uid - ORC_UID_MIN])) return 1;
  if (sol->lchanged || !sol->lWm2_diff) return 2;
  if (sol->lsolar_rad && edir) {
    ledir = (double *)calloc(ngh(s, Nd), sizeof(double));
    interior_to_ghosted(s, Nd, sol->edir, ledir);
    halo_fill_5pt(s, Nd, ledir);
  }
  lediff = (double *)calloc(ngh(s, Nf), sizeof(double));
  interior_to_ghosted(s, Nf, sol->ediff, lediff);
  halo_fill_5pt(s, Nf, lediff);
  for (m = 0; m < s->nbld; ++m) {
    const double mu = sol->lsolar_rad ? s->mu : 1.0;
    int kk, ii, jj;
    double in = 0, out = 0;
    bld_idx(s, m, &face, &k, &i, &j);
    kk = k; ii = i; jj = j;
    if (face == ORC_BOT_FACE) kk = k + 1;
    if (face == ORC_RIGHT_FACE) ii = i + 1;
    if (face == ORC_FRONT_FACE) jj = j + 1;
    if (ledir) {
      double sum = 0;
      if (face == ORC_TOP_FACE || face == ORC_BOT_FACE) {
        for (d = 0; d < Ndt; ++d) sum += FG(ledir, Nd, d, kk, ii, jj) * s->mu;
        edir[m] = sum / (double)s->st.dirtop_adiv;
      } else {
        const int s0 = (face == ORC_LEFT_FACE || face == ORC_RIGHT_FACE) ? Ndt : Ndt + Nds;
        for (d = s0; d < s0 + Nds; ++d) sum += FG(ledir, Nd, d, kk, ii, jj) * s->mu;
        edir[m] = sum / (double)s->st.dirside_adiv;
      }
    } else if (edir) edir[m] = 0.0;
    if (face == ORC_TOP_FACE || face == ORC_BOT_FACE) {
      /* top face: inward (down) streams enter the building; bottom face: the other way round */
      for (d = 0; d < Nft; ++d) {
        const double v = FG(lediff, Nf, d, kk, ii, jj) * mu;
        if ((s->st.ft_inward[d] != 0) == (face == ORC_TOP_FACE)) in += v; else out += v;
      }
    } else {
      const int off = (face == ORC_LEFT_FACE || face == ORC_RIGHT_FACE) ? Nft : Nft + Nfs;
      const int low = face == ORC_LEFT_FACE || face == ORC_REAR_FACE;
      for (d = 0; d < Nfs; ++d) {
        const double v = FG(lediff, Nf, off + d, kk, ii, jj) * mu;
        if ((s->st.fs_inward[d] != 0) == low) in += v; else out += v;
      }
    }
    incoming[m] = in / 1.0; outgoing[m] = out / 1.0; /* difftop/diffside%area_divider = 1 */
  }
  free(ledir); free(lediff);
  return 0;
}

/* slope_correction_fluxes (src/pprts_postprocess.F90:130-196), the callers' post-processing of the direct flux on a
 * terrain-following grid: hgrad (src/pprts.F90:621-638) = centred differences of the level heights hhl = -cumsum(dz)
 * with the periodic halo of one rank; -pprts_atm_correction scales every level by (sundir.ez)/(sundir.n),
 * -pprts_slope_correction the surface level by max(0, (sundir.n)/(sundir.ez)), n the normal of the level surface.
 * edir: (Nz+1, Nx, Ny) Fortran order on the solver grid, in place.  Returns 1 without a 3-D dz (hgrad missing). */
int orc_slope_correction_fluxes(orc_solver *s, int lslope, int latm, int have_dz3d, double *edir) {
  LOCALS
  int k, i, j;
  if (!lslope && !latm) return 0;
  if (!have_dz3d) return 1;
  for (j = 0; j < Ny; ++j) for (i = 0; i < Nx; ++i) for (k = 0; k < Nz1; ++k) {
    const int ka = k + s->icollapse - 1;
    const int im = (i + Nx - 1) % Nx, ip = (i + 1) % Nx, jm = (j + Ny - 1) % Ny, jp = (j + 1) % Ny;
    const double gx = (vert_height(s, ka, ip, j) - vert_height(s, ka, im, j)) / (2.0 * s->dx);
    const double gy = (vert_height(s, ka, i, jp) - vert_height(s, ka, i, jm)) / (2.0 * s->dy);
    /* n = cross([1,0,gx],[0,1,gy]) normalised */
    double n[3] = {-gx, -gy, 1.0}, nn = sqrt(gx * gx + gy * gy + 1.0), sn, sz = s->sundir[2];
    n[0] /= nn; n[1] /= nn; n[2] /= nn;
    sn = s->sundir[0] * n[0] + s->sundir[1] * n[1] + s->sundir[2] * n[2];
    if (latm) P3(edir, k, i, j) = P3(edir, k, i, j) * (sz / sn);
    if (lslope && k == Nz) P3(edir, k, i, j) = P3(edir, k, i, j) * dmax(0.0, sn / sz);
  }
  return 0;
}

/* box_blur_1d inside convolve_ediff_srfc (src/pprts_base.F90:1816-1866), iter = 1: running window sum of 2k+1 entries
 * along the second (slow) index of im(n1, n2) (Fortran order), the window walked with the reference's bounce at both
 * array ends (an index that steps past an end repeats the end value and turns round), then scaled by 1/(2k+1). */
static void box_blur_1d(const double *im, int n1, int n2, int k, double *img) {
  double *tmp = (double *)malloc(sizeof(double) * (size_t)n1 * n2);
  const int lb = 0, ub = n2 - 1;
  int a, i, ki, l, r, inc;
#define IM(a_, b_) im[(a_) + (size_t)n1 * (b_)]
#define TMP(a_, b_) tmp[(a_) + (size_t)n1 * (b_)]
  for (a = 0; a < n1; ++a) TMP(a, lb) = IM(a, lb);
  inc = -1; l = lb;
  for (ki = 0; ki < k; ++ki) {
    l += inc;
    if (l == lb - 1) { l = lb; inc = 1; }
    if (l == ub + 1) { l = ub; inc = -1; }
    for (a = 0; a < n1; ++a) TMP(a, lb) += IM(a, l);
  }
  r = lb; inc = 1;
  for (ki = 0; ki < k; ++ki) {
    r += inc;
    if (r == ub + 1) { r = ub; inc = -1; }
    if (r == lb - 1) { r = lb; inc = 1; }
    for (a = 0; a < n1; ++a) TMP(a, lb) += IM(a, r);
  }
  for (i = lb + 1; i <= ub; ++i) {
    l = i - 1; inc = -1;
    for (ki = 0; ki < k; ++ki) {
      l += inc;
      if (l == lb - 1) { l = lb; inc = 1; }
      if (l == ub + 1) { l = ub; inc = -1; }
    }
    r = i; inc = 1;
    for (ki = 0; ki < k; ++ki) {
      r += inc;
      if (r == ub + 1) { r = ub; inc = -1; }
      if (r == lb - 1) { r = lb; inc = 1; }
    }
    for (a = 0; a < n1; ++a) TMP(a, i) = TMP(a, i - 1) + IM(a, r) - IM(a, l);
  }
  for (a = 0; a < n1 * n2; ++a) img[a] = tmp[a] * (1.0 / (double)(2 * k + 1));
#undef IM
#undef TMP
  free(tmp);
}

/* convolve_ediff_srfc (src/pprts_base.F90:1733-1868) on one rank (every neighbour is the rank itself, so the ghost
 * exchange is the periodic wrap): arr(dof, xm, ym), Fortran order, in place.  Extended array with kw ghosts per side,
 * x ghosts first, then y ghosts over the full extended x range (corners), box_filter = blur along x then along y
 * (`box_filter` transposes so that the blurred index is the slow one), interior copied back.  Returns 1 if kw does
 * not fit the domain (the reference's sections xm+1:xm+kw would leave the interior). */
int orc_convolve_ediff_srfc(int dof, int xm, int ym, int kw, double *arr) {
  const int gx = xm + 2 * kw, gy = ym + 2 * kw;
  double *wg, *imT, *tmpT, *tmp;
  int d, i, j;
  if (kw < 0 || kw > xm || kw > ym) return 1;
  wg = (double *)calloc((size_t)gx * gy, sizeof(double));
  imT = (double *)malloc(sizeof(double) * (size_t)gx * gy);
  tmpT = (double *)malloc(sizeof(double) * (size_t)gx * gy);
  tmp = (double *)malloc(sizeof(double) * (size_t)gx * gy);
#define ARR(d_, i_, j_) arr[(d_) + (size_t)dof * ((i_) + (size_t)xm * (j_))]
#define WG(i_, j_) wg[(i_) + (size_t)gx * (j_)]
  for (d = 0; d < dof; ++d) {
    for (j = 0; j < ym; ++j) for (i = 0; i < xm; ++i) WG(kw + i, kw + j) = ARR(d, i, j);
    /* pass 1: x ghosts of the interior rows: west ghost <- the east-most kw interior columns, and vice versa */
    for (j = 0; j < ym; ++j) for (i = 0; i < kw; ++i) {
      WG(i, kw + j) = WG(xm + i, kw + j);
      WG(xm + kw + i, kw + j) = WG(kw + i, kw + j);
    }
    /* pass 2: y ghosts over the full extended x range */
    for (j = 0; j < kw; ++j) for (i = 0; i < gx; ++i) {
      WG(i, j) = WG(i, ym + j);
      WG(i, ym + kw + j) = WG(i, kw + j);
    }
    /* box_filter: im_T(y, x) blurred along x, transposed back, blurred along y */
    for (j = 0; j < gy; ++j) for (i = 0; i < gx; ++i) imT[j + (size_t)gy * i] = WG(i, j);
    box_blur_1d(imT, gy, gx, kw, tmpT);
    for (j = 0; j < gy; ++j) for (i = 0; i < gx; ++i) tmp[i + (size_t)gx * j] = tmpT[j + (size_t)gy * i];
    box_blur_1d(tmp, gx, gy, kw, wg);
    for (j = 0; j < ym; ++j) for (i = 0; i < xm; ++i) ARR(d, i, j) = WG(kw + i, kw + j);
  }
#undef ARR
#undef WG
  free(wg); free(imT); free(tmpT); free(tmp);
  return 0;
}

/* find_iter_and_kernelwidth (src/pprts_postprocess.F90:86-128): Niter box blurs of half width kernel_width approximate
 * a Gaussian of the given radius [m]; the width closest to an integer among 500 candidate iteration counts, the
 * candidates whose width does not fit the local domain skipped.  Returns the reference's error (min_iter) if more
 * than 250 iterations would be needed. */
int orc_find_iter_and_kernelwidth(double radius, double dx, double dy, int xm, int ym, int *Niter, int *kernel_width) {
  enum { test_Ni = 500, Ni_limit = 250 };
  double test_k[test_Ni], best = 0.0;
  const double rpix = radius / ((dx + dy) / 2);
  const int mloc = xm < ym ? xm : ym;
  int i, min_iter = 1, arg = -1;
  for (i = 1; i <= test_Ni; ++i) {
    test_k[i - 1] = (sqrt((12 * rpix * rpix + (double)i) / (double)i + 1) - 1) / 2;
    if ((int)lround(test_k[i - 1]) >= mloc) min_iter = i + 1;
  }
  if (min_iter > Ni_limit) return min_iter;
  for (i = min_iter; i <= test_Ni; ++i) {
    const double res = fabs(test_k[i - 1] - (double)lround(test_k[i - 1]));
    if (arg < 0 || res < best) { best = res; arg = i; } /* minloc: first minimum */
  }
  *Niter = arg;
  *kernel_width = (int)lround(test_k[arg - 1]);
  if (*kernel_width == 0) *Niter = 0;
  return 0;
}

/* smooth_surface_fluxes (src/pprts_postprocess.F90:44-84), the spectral callers' post-processing of the diffuse
 * surface fluxes (rrtmg/rrtmg/pprts_rrtmg.F90:319), option -pprts_smooth_srfc_flx <radius> passed as an argument:
 * radius < 0 replaces the surface level of edn and eup by its domain mean (meanval = sum / size), radius >= 0 applies
 * Niter box blurs of half width kernel_width to it.  edn, eup: (Nz1, Nx, Ny) Fortran order, in place. */
int orc_smooth_surface_fluxes(int Nz1, int Nx, int Ny, double dx, double dy, double radius, double *edn, double *eup) {
  double *f[2], *srfc;
  const size_t n = (size_t)Nx * Ny;
  size_t c;
  int q, it, Niter = 0, kw = 0, rc = 0;
  f[0] = edn; f[1] = eup;
  srfc = (double *)malloc(sizeof(double) * n);
  if (radius >= 0 && (rc = orc_find_iter_and_kernelwidth(radius, dx, dy, Nx, Ny, &Niter, &kw))) { free(srfc); return rc; }
  for (q = 0; q < 2; ++q) {
    for (c = 0; c < n; ++c) srfc[c] = f[q][(Nz1 - 1) + (size_t)Nz1 * c];
    if (radius < 0) {
      double sum = 0.0;
      for (c = 0; c < n; ++c) sum += srfc[c];
      sum = sum / (double)n * (double)n / (double)n; /* imp_allreduce_mean: meanval(v) * size(v) / global_size */
      for (c = 0; c < n; ++c) srfc[c] = sum;
    } else {
      for (it = 0; it < Niter && !rc; ++it) rc = orc_convolve_ediff_srfc(1, Nx, Ny, kw, srfc);
    }
    for (c = 0; c < n; ++c) f[q][(Nz1 - 1) + (size_t)Nz1 * c] = srfc[c];
  }
  free(srfc);
  return rc;
}

const double *orc_coeff_ptr(orc_solver *s, int which) { return which == 1 ? s->dir2dir : which == 2 ? s->dir2diff : s->diff2diff; }
const double *orc_state_ptr(orc_solver *s, int uid, int which) {
  orc_solution *sol = s->sol[uid - ORC_UID_MIN];
  if (which == 4) return s->b;
  if (!sol) return 0;
  return which == 1 ? sol->edir : which == 2 ? sol->ediff : sol->abso;
}
const int *orc_l1d_ptr(orc_solver *s) { return s->l1d; }
void orc_get_info(orc_solver *s, int uid, int *niter_dir, int *niter_diff, double *res_dir, double *res_diff) {
  orc_solution *sol = s->sol[uid - ORC_UID_MIN];
  if (!sol) return;
  *niter_dir = sol->Niter_dir; *niter_diff = sol->Niter_diff;
  *res_dir = sol->final_res_dir; *res_diff = sol->final_res_diff;
}
