// f2c.cu -- the pprts_f2c_* entry points of the reference's C wrapper (c_wrapper/f2c_pprts.h:61-99, implemented in
// c_wrapper/f2c_pprts.F90) on top of the handle API.  Like the reference this layer keeps ONE module-level solver
// (f2c_pprts.F90:106) and turns every failure into "print + abort" (CHKERR -> mpi_abort,
// src/helper_functions.fypp:888-906).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "internal.h"
#include "lut_axes.h"

static pprts_b200_solver *g_solver = nullptr;
static pprts_b200_lut *g_lut = nullptr;
static bool g_lut_owned = false;
static int g_lut_users = 0;  // the module-level solver and every live OPP container hold the default table set
static int g_solver_id = 0;
static void (*g_abort_handler)(const char *) = nullptr;

static void chkerr(int rc, const char *where) {
  if (rc == 0) return;
  char msg[1400];
  snprintf(msg, sizeof(msg), "pprts_b200 :: %s failed (%d): %s", where, rc, pprts_b200_last_error());
  if (g_abort_handler) { g_abort_handler(msg); return; }
  fprintf(stderr, "%s\n", msg);
  abort();
}

extern "C" void pprts_b200_f2c_set_abort_handler(void (*h)(const char *)) { g_abort_handler = h; }
// Register an already opened table set for the module-level solver (tests use reduced parameter spaces).
extern "C" void pprts_b200_f2c_set_lut(pprts_b200_lut *lut) {
  if (g_lut && g_lut_owned && g_lut_users == 0) pprts_b200_lut_close(g_lut);
  g_lut = lut;
  g_lut_owned = false;
}
// the default table set is closed when its last user (module-level solver or OPP container) goes away
static void release_default_lut() {
  if (g_lut_users > 0) --g_lut_users;
  if (g_lut_users > 0) return;
  if (g_lut && g_lut_owned) pprts_b200_lut_close(g_lut);
  g_lut = nullptr;
  g_lut_owned = false;
}

static std::vector<float> linspace_f32(float lo, float hi, int n) {  // helper_functions.fypp:2735-2745
  std::vector<float> v(n);
  for (int k = 0; k < n; ++k) v[k] = n > 1 ? lo + (float)k * (hi - lo) / (float)(n - 1) : lo;
  return v;
}

// default table set of a solver: <basename>_direct_<Nd>_<Nf>.tau31.w020.aspect_zx23.g6.phi19.theta19.ds1000.nc.*.mmap4
// (src/optprop_LUT.F90:364-374, :503-518)
static int open_default_lut(int solver_id) {
  if (g_lut) return 0;
  const char *base = getenv("PPRTS_B200_LUT_BASENAME");
  if (!base) base = getenv("LUT_BASENAME");
  if (!base) base = "./LUT";
  int Nd = solver_id / 100, Nf = solver_id % 100;
  std::vector<float> phi = linspace_f32(0.f, 90.f, 19), theta = linspace_f32(0.f, 90.f, 19);
  const int n[6] = {31, 20, 23, 6, 19, 19};
  const float *ax[6] = {PB_TAU31, PB_W020, PB_ASPECT23, PB_G6, phi.data(), theta.data()};
  char dirn[900], difn[900];
  snprintf(dirn, sizeof(dirn), "%s_direct_%d_%d.tau31.w020.aspect_zx23.g6.phi19.theta19.ds1000.nc", base, Nd, Nf);
  snprintf(difn, sizeof(difn), "%s_diffuse_%d.tau31.w020.aspect_zx23.g6.ds1000.nc", base, Nf);
  std::string T = std::string(dirn) + ".Tdir.mmap4", S = std::string(dirn) + ".Sdir.mmap4", D = std::string(difn) + ".Sdiff.mmap4";
  int rc = pprts_b200_lut_open(solver_id, T.c_str(), S.c_str(), D.c_str(), n, ax, &g_lut);
  g_lut_owned = rc == 0;
  return rc;
}

extern "C" void pprts_f2c_init(int fcomm, int *solver_id, int *Nz, int *Nx, int *Ny, double *dx, double *dy, float *hhl,
                               float *phi0, float *theta0, int *collapseindex) {
  (void)fcomm;
  if (g_solver) {  // f2c_pprts.F90:149-187: re-init with the same solver type is a no-op
    if (*solver_id != g_solver_id) chkerr(1, "pprts_f2c_init: seems you changed the solver type id in between calls... "
                                             "you must destroy the solver first");
    return;
  }
  (void)collapseindex;  // broadcast back but never handed to init_pprts by the reference wrapper (f2c_pprts.F90:198-260)
  chkerr(open_default_lut(*solver_id), "pprts_f2c_init: loading LUT");
  pprts_b200_config cfg;
  pprts_b200_default_config(&cfg);
  cfg.solver_id = *solver_id;
  cfg.Nz = *Nz; cfg.Nx = *Nx; cfg.Ny = *Ny;
  cfg.dx = *dx; cfg.dy = *dy;
  std::vector<double> dz(*Nz);
  for (int k = 0; k < *Nz; ++k) dz[k] = (double)hhl[k] - (double)hhl[k + 1];  // f2c_pprts.F90:228-231
  cfg.dz1d = dz.data();
  const double d2r = 3.14159265358979323846 / 180.0, p = (double)*phi0, t = (double)*theta0;
  cfg.sundir[0] = -std::sin(t * d2r) * std::sin(p * d2r);  // spherical_2_cartesian, helper_functions.fypp:2022-2033
  cfg.sundir[1] = -std::sin(t * d2r) * std::cos(p * d2r);
  cfg.sundir[2] = -std::cos(t * d2r);
  chkerr(pprts_b200_create(&cfg, g_lut, &g_solver), "pprts_f2c_init: init_pprts");
  if (g_solver) ++g_lut_users;
  g_solver_id = *solver_id;
}

extern "C" void pprts_f2c_set_global_optical_properties(int Nz, int Nx, int Ny, float *albedo, float *kabs, float *ksca,
                                                        float *g, float *planck) {
  if (!g_solver) return;
  const size_t nc = (size_t)Nz * Nx * Ny, nl = (size_t)(Nz + 1) * Nx * Ny;
  std::vector<double> k(nc), s(nc), gg(nc), pl;
  for (size_t q = 0; q < nc; ++q) { k[q] = kabs[q]; s[q] = ksca[q]; gg[q] = g[q]; }
  bool any = false;
  if (planck) for (size_t q = 0; q < nl; ++q) if (planck[q] > 0.f) { any = true; break; }
  if (any) { pl.resize(nl); for (size_t q = 0; q < nl; ++q) pl[q] = planck[q]; }  // f2c_pprts.F90:303-307
  chkerr(pprts_b200_set_optical_properties(g_solver, (double)*albedo, k.data(), s.data(), gg.data(),
                                           any ? pl.data() : nullptr, nullptr, nullptr, 1),
         "pprts_f2c_set_global_optical_properties");
}

extern "C" void pprts_f2c_solve(int fcomm, float edirTOA) {
  (void)fcomm;
  if (!g_solver) return;
  const bool lsolar = edirTOA > 0.f;  // f2c_pprts.F90:340-342
  chkerr(pprts_b200_solve(g_solver, !lsolar, lsolar, (double)edirTOA, 0), "pprts_f2c_solve");
}

extern "C" void pprts_f2c_get_result(int Nz, int Nx, int Ny, float *edn, float *eup, float *abso, float *edir) {
  if (!g_solver) return;
  const size_t nc = (size_t)Nz * Nx * Ny, nl = (size_t)(Nz + 1) * Nx * Ny;
  std::vector<double> dn(nl), up(nl), dr(nl), ab(nc);
  chkerr(pprts_b200_get_result(g_solver, 0, dn.data(), up.data(), ab.data(), dr.data()), "pprts_f2c_get_result");
  for (size_t q = 0; q < nl; ++q) { edn[q] = (float)dn[q]; eup[q] = (float)up[q]; edir[q] = (float)dr[q]; }
  for (size_t q = 0; q < nc; ++q) abso[q] = (float)ab[q];
}

extern "C" void pprts_f2c_destroy(int lfinalizepetsc) {
  (void)lfinalizepetsc;
  if (g_solver) { pprts_b200_destroy(g_solver); g_solver = nullptr; release_default_lut(); }
}

// ---- coefficient probe (f2c_pprts.F90:591-758) ---------------------------------------------------------------------
struct OppContainer {
  pprts_b200_solver *solver;
  pprts_b200_lut *lut;
  bool holds_default_lut;
  double phi, theta;
};
extern "C" void pprts_f2c_opp_init(const int comm, const int solver_id, void **opp, int *ierr) {
  (void)comm;
  *ierr = 0;
  if (solver_id != SOLVER_ID_PPRTS_3_10) {  // f2c_pprts.F90:605-613
    fprintf(stderr, "pprts_f2c_init_OPP not implemented for solver_id %d\n", solver_id);
    *ierr = 1;
    chkerr(1, "pprts_f2c_opp_init");
    return;
  }
  auto *c = new OppContainer();
  chkerr(open_default_lut(solver_id), "pprts_f2c_opp_init: loading LUT");
  c->lut = g_lut;
  c->holds_default_lut = false;
  pprts_b200_config cfg;
  pprts_b200_default_config(&cfg);
  cfg.solver_id = solver_id;
  cfg.Nz = cfg.Nx = cfg.Ny = 1;
  cfg.dx = cfg.dy = 1.0;
  double dz = 1.0;
  cfg.dz1d = &dz;
  cfg.sundir[0] = 0; cfg.sundir[1] = 0; cfg.sundir[2] = -1;
  chkerr(pprts_b200_create(&cfg, c->lut, &c->solver), "pprts_f2c_opp_init");
  if (c->solver) { ++g_lut_users; c->holds_default_lut = true; }
  c->phi = c->theta = -999;
  *opp = c;
}
extern "C" void pprts_f2c_opp_get_coeff(void *opp, const float tauz, const float w0, const float g, const float aspect_zx,
                                        const float phi, const float theta, const int imode, const int lswitch_east,
                                        const int lswitch_north, const int Ncoeff, float *coeff, int *ierr) {
  auto *c = (OppContainer *)opp;
  *ierr = 0;
  if (imode < 1 || imode > 3) { *ierr = 1; chkerr(1, "pprts_f2c_opp_get_coeff: imode option not recognized"); return; }
  if (imode != 3 && ((double)phi != c->phi || (double)theta != c->theta)) {
    pprts_b200_internal_force_angles(c->solver, (double)phi, (double)theta);
    c->phi = phi; c->theta = theta;
  }
  const float sample[4] = {tauz, w0, aspect_zx, g};
  std::vector<double> out(Ncoeff > 100 ? Ncoeff : 100);
  int rc = pprts_b200_internal_opp_probe(c->solver, imode, sample, imode == 3 ? 0 : lswitch_east,
                                         imode == 3 ? 0 : lswitch_north, out.data());
  if (rc) { *ierr = rc; chkerr(rc, "pprts_f2c_opp_get_coeff"); return; }
  for (int q = 0; q < Ncoeff; ++q) coeff[q] = (float)out[q];
}
extern "C" void pprts_f2c_opp_destroy(void *opp, int *ierr) {
  auto *c = (OppContainer *)opp;
  *ierr = 0;
  if (!c) return;
  pprts_b200_destroy(c->solver);
  if (c->holds_default_lut && c->lut == g_lut) release_default_lut();
  delete c;
}
extern "C" void pprts_f2c_opp_get_info(void *opp, int *Ndir, int *Ndiff, float *diff_tauz_range, float *diff_w0_range,
                                       float *diff_g_range, float *diff_aspect_zx_range, float *dir_tauz_range,
                                       float *dir_w0_range, float *dir_g_range, float *dir_aspect_zx_range,
                                       float *dir_phi_range, float *dir_theta_range, int *ierr) {
  auto *c = (OppContainer *)opp;
  float r[6][2];
  *ierr = pprts_b200_internal_lut_ranges(c->lut, r, Ndir, Ndiff);
  // axis order [tau, w0, aspect_zx, g, phi, theta]; the diffuse table shares the first four axes
  for (int q = 0; q < 2; ++q) {
    diff_tauz_range[q] = dir_tauz_range[q] = r[0][q];
    diff_w0_range[q] = dir_w0_range[q] = r[1][q];
    diff_aspect_zx_range[q] = dir_aspect_zx_range[q] = r[2][q];
    diff_g_range[q] = dir_g_range[q] = r[3][q];
    dir_phi_range[q] = r[4][q];
    dir_theta_range[q] = r[5][q];
  }
}
