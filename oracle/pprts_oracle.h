/*
 * pprts_oracle.h -- CPU oracle for the TenStream pprts solve path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, fp64 restatement of the reference's own PETSc-free formulation of
 * the path (src/pprts_explicit.F90, src/pprts_shell.F90, the optprop_LUT lookup, ...).  It exists so that the
 * CUDA library can be parity-checked without a Fortran/MPI/PETSc toolchain.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product (tenstream_b200/) never does.
 *
 * PARITY PINNING: the reference cannot be compiled in this environment (no gfortran/MPI/PETSc/NetCDF/fypp) and
 * its real LUT files are not available offline.  What pins this oracle is every LUT-free known-answer test of
 * the reference (tests/test_oracle_known_answers.py: interpolation, search, eddington, symmetry involutions,
 * mmap round trip, scale round trip) plus the reference's property tests run on synthetic tables
 * (absorption-by-divergence == absorption-by-coefficients, phi / phi+180 mirror symmetry, energy closure), the
 * reference's buildings / collapse / error-growth / geometric-coefficient / slope / convolution tests restated in
 * tests/test_oracle_*.py, and a second, independent restatement of the path as sparse matrices in numpy / scipy
 * (tests/test_oracle_independent_restatement.py: direct beam, sources, diffuse operator and solve, absorption for
 * 3_10 / 8_10 / 3_24; tests/test_oracle_open_bc.py for the open-boundary beam) that agrees with this file to rounding.
 * End-to-end parity against reference-produced flux files is UNPINNED.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference tree).
 */
#ifndef PPRTS_ORACLE_H
#define PPRTS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_DIM 6
#define ORC_MAX_DOF 32

typedef struct {
  int solver_id;                 /* c_wrapper/f2c_solver_ids.h: 310, 810, 324 */
  int Ndt, Nds, Nft, Nfs;        /* dirtop%dof, dirside%dof, difftop%dof, diffside%dof (src/pprts.F90:332-396) */
  int Nd, Nf;                    /* Ndt + 2 Nds, Nft + 2 Nfs */
  int dirtop_adiv, dirside_adiv; /* area_divider (src/pprts.F90:250-254, :362-366) */
  int ft_inward[ORC_MAX_DOF];    /* difftop%is_inward */
  int fs_inward[ORC_MAX_DOF];    /* diffside%is_inward */
} orc_streams;

typedef struct {
  int ndim;                      /* 6 (direct) or 4 (diffuse) */
  int n[ORC_MAX_DIM];
  const float *ax[ORC_MAX_DIM];  /* axis values (fp32, as irealLUT) */
  long off[ORC_MAX_DIM];         /* ndarray_offsets: first axis fastest */
} orc_lut_axes;

typedef struct {
  orc_lut_axes dir, diff;
  const float *Tdir, *Sdir, *Sdiff; /* (Ncoef, Nlattice) column major == [lattice][coef] in C */
  int Nd, Nf;
  void *map[3];
  unsigned long maplen[3];
} orc_lut;

/* ---- scalar building blocks ------------------------------------------------------------------------------- */
double orc_search_sorted_bisection(const double *arr, int n, double val);      /* src/search.fypp:177-232 */
double orc_find_real_location_f32axis(const float *arr, int n, float val);     /* same, fp32 axis, fp64 arithmetic */
int orc_dim_needs_interpolation(double pti);                                   /* src/interpolation.F90:546-556 */
void orc_interp_vec_nd(const double *pti, int ndim, const float *db, int ncoef, const long *offsets,
                       double *out);                                           /* src/interpolation.F90:317-360 */
void orc_eddington_coeff_ec(double dtau, double w0, double g, double mu0, double *t, double *r, double *rdir,
                            double *sdir, double *tdir);                       /* src/eddington.F90:167-242 */
double orc_B_eff(double B_far, double B_near, double tau);                     /* src/schwarzschild.F90:36-67 */
void orc_delta_scale(double *kabs, double *ksca, double *g);                   /* helper_functions.fypp:1622-1666 */
/* adding (src/pprts.F90:2125-2198): collapse n 1-D layers (TOA first) into the last one, in place; with dtau/planck
 * (n / n+1 values, may be NULL) also the layer's thermal emission Btop, Bbot via schwarzschild (src/schwarzschild.F90:82-150) */
void orc_adding(int n, double *a11, double *a12, double *a21, double *a22, double *a13, double *a23, double *a33,
                const double *dtau, const double *planck, double *Btop, double *Bbot);
void orc_schwarzschild2(int ke, const double *dtau, double albedo, double *Edn, double *Eup, const double *planck,
                        const double *opt_srfc_emission);
int orc_streams_init(orc_streams *st, int solver_id);                          /* src/pprts.F90:256-453 */
void orc_dir2diff_symmetry(const orc_streams *st, double *c, int lswitch_east, int lswitch_north);
void orc_dir2dir_symmetry(const orc_streams *st, double *c, int lswitch_east, int lswitch_north);

/* ---- LUT --------------------------------------------------------------------------------------------------- */
int orc_mmap4_open(const char *path, const float **data, long *ncoef, long *nlat, void **map,
                   unsigned long *maplen);                                     /* src/mmap.F90:129-203 */
int orc_mmap4_write(const char *path, const float *data, long ncoef, long nlat); /* src/mmap.F90:63-127 */
int orc_lut_load(orc_lut *lut, int solver_id, const char *path_Tdir, const char *path_Sdir,
                 const char *path_Sdiff, const int *ndir, const float *const *axdir);
void orc_lut_free(orc_lut *lut);
/* src/pprts_base.F90:1491-1543 + src/optprop.F90:550-593 + src/optprop_LUT.F90:1452-1613.
 * mode 1 = dir2dir, 2 = dir2diff, 3 = diff2diff (c_wrapper/f2c_pprts.F90:659-681). phi = symmetry_phi. */
int orc_get_coeff(const orc_lut *lut, const orc_streams *st, double kabs, double ksca, double g, double dz,
                  double dx, int mode, double phi, double theta, int lswitch_east, int lswitch_north,
                  double *coeff);
/* direct LUT-space variant used by the coefficient probe (c_wrapper/f2c_pprts.F90:628-685) */
int orc_opp_get_coeff(const orc_lut *lut, const orc_streams *st, float tauz, float w0, float g, float aspect_zx,
                      float phi, float theta, int mode, int lswitch_east, int lswitch_north, double *coeff);

/* ---- solver ------------------------------------------------------------------------------------------------- */
typedef struct orc_solver orc_solver;

typedef struct {
  int solver_id;
  int Nz, Nx, Ny;
  double dx, dy;
  const double *dz;      /* (Nz,Nx,Ny) Fortran order: z fastest */
  double sundir[3];
  double twostr_ratio;   /* src/tenstream_options.F90: 2.0 */
  int coeff_fp32;        /* 1: round stored 3-D coefficients through fp32 (what the reference's fp32 LUT path yields) */
  int by_coeff_divergence; /* src/pprts.F90:5286 default .true. */
  int diff_solver;       /* 0: explicit SOR (src/pprts_explicit.F90:461), 1: BiCGStab on the shell operator */
  double rtol, atol;     /* <0: reference defaults (src/pprts_base.F90:1128-1131) */
  int maxit_dir, maxit_diff; /* <=0: reference defaults 1000 / 10000 */
  int nthreads;          /* OpenMP threads for the embarrassingly parallel loops (coefficients, MatMult) */
  int collapseindex;     /* init_pprts argument (src/pprts.F90:692-706): > 1 merges the top `collapseindex` atmosphere
                            layers into one 1-D layer; Nz above is then the ATMOSPHERE layer count, the solver grid and
                            all results have Nz - collapseindex + 1 layers. <= 1: off */
  int geometric_coeffs;     /* -pprts_geometric_coeffs (src/pprts.F90:3046-3053): dir2dir from the cell geometry
                               (src/geometric_coeffs.F90), planar and parallel top/bottom faces, dir2diff rescaled to
                               conserve the LUT's absorption (-pprts_conserve_lut_atm_abso, :3324-3351); 3 direct streams;
                               + 2 = -pprts_top_bottom_faces_planar no, + 4 = -pprts_top_bottom_planes_parallel no */
  double max_solution_err;  /* -max_solution_err [W], src/tenstream_options.F90:90 default 5e3/86400; <= 0 disables */
  double max_solution_time; /* -max_solution_time [s], :94 default 0 = error estimates disabled */
  int open_bc;              /* -pprts_open_bc (src/pprts.F90:462-463): the direct beam enters through the sunward domain
                               edges with the side fluxes of a single-column solve and leaves through the others
                               (src/pprts_base.F90:1189-1316, src/pprts_explicit.F90:262-306, :358-378) */
} orc_config;

orc_solver *orc_solver_create(const orc_config *cfg, const orc_lut *lut);      /* init_pprts, src/pprts.F90:213 */
void orc_solver_destroy(orc_solver *s);
void orc_set_angles(orc_solver *s, const double sundir[3]);                    /* src/pprts.F90:1100-1183 */
/* set_optical_properties (src/pprts.F90:1764-2076). Arrays (Nz,Nx,Ny) / planck (Nz+1,Nx,Ny) / 2-D (Nx,Ny),
 * Fortran order; planck, planck_srfc, albedo_2d may be NULL. */
int orc_set_optical_properties(orc_solver *s, double albedo, const double *kabs, const double *ksca,
                               const double *g, const double *planck, const double *planck_srfc,
                               const double *albedo_2d, int ldelta_scaling);
int orc_solve(orc_solver *s, int lthermal, int lsolar, double edirTOA, int uid); /* src/pprts.F90:2487-3033 */
/* pprts_get_result (src/pprts.F90:5799-5917): edn,eup,edir (Nz+1,Nx,Ny), abso (Nz,Nx,Ny); edir may be NULL */
int orc_get_result(orc_solver *s, int uid, double *edn, double *eup, double *abso, double *edir);

/* error-growth tracking (src/adaptive_spectral_integration.F90:38-257; restore_solution src/pprts.F90:3991-4049):
 * orc_solve_t = solve_pprts(..., opt_solution_time = time); need_new_solution tells the band loop whether the stored
 * solution of `uid` is still good at `time` (have_time = 0: time not present -> always 1) */
int orc_solve_t(orc_solver *s, int lthermal, int lsolar, double edirTOA, int uid, double time);
int orc_need_new_solution(orc_solver *s, int uid, double time, int have_time);
int orc_polyfit(const double *vx, const double *vy, int n, int d, double *coef);
void orc_get_solution_times(orc_solver *s, int uid, double *time30, double *maxnorm30);

/* dir2dir3_geometric_coeffs (src/geometric_coeffs.F90:34-133): analytic dir2dir (9 values, flat index dst*3 + src as
 * everywhere) of a 3-stream box with vertices a..h (24 values) for the sun vector `sundir` and extinction c_ext [1/m];
 * num_intervals <= 0: the reference default of 10 */
int orc_dir2dir3_geometric_coeffs(const double *vertices, const double *sundir, double c_ext, int num_intervals, double *coeffs);

/* the helpers of the geometric coefficients: hit_plane, line_intersection_3d (src/intersection.fypp:79-91, :395-434),
 * compute_normal_3d, triangle_area_by_vertices (src/helper_functions.fypp:1960-1974, :2116-2136) */
double orc_hit_plane(const double *p_loc, const double *p_dir, const double *po, const double *pn);
void orc_compute_normal_3d(const double *p1, const double *p2, const double *p3, double *n);
int orc_line_intersection_3d(const double *p1, const double *d1, const double *p3, const double *d2, double *c1, double *c2);
double orc_triangle_area_by_vertices(const double *v1, const double *v2, const double *v3);

/* opt_buildings of solve_pprts / pprts_get_result (src/buildings.F90:40-69, src/pprts.F90:3194-3212, :3579-3677,
 * :4989-5145, :6011-6160).  iface: 1-based flattened index into (6,Nz,Nx,Ny), first axis fastest; face ids 1..6 =
 * top, bottom, left (x+0), right (x+1), rear (y+0), front (y+1).  planck may be NULL.  nfaces = 0: no buildings. */
int orc_set_buildings(orc_solver *s, int nfaces, const int *iface, const double *albedo, const double *planck);
int orc_get_buildings_result(orc_solver *s, int uid, double *edir, double *incoming, double *outgoing);

/* slope_correction_fluxes (src/pprts_postprocess.F90:130-196): -pprts_slope_correction / -pprts_atm_correction applied
 * to a direct-flux result (Nz+1,Nx,Ny) in place; have_dz3d = 0 (no hgrad) with a correction requested returns 1 */
int orc_slope_correction_fluxes(orc_solver *s, int lslope, int latm, int have_dz3d, double *edir);
/* smooth_surface_fluxes (src/pprts_postprocess.F90:44-128) and the separable box blur it applies, convolve_ediff_srfc
 * (src/pprts_base.F90:1733-1868), restated for one rank; pinned on tests/test_convolution/test_convolution.F90 */
int orc_convolve_ediff_srfc(int dof, int xm, int ym, int kw, double *arr);
int orc_find_iter_and_kernelwidth(double radius, double dx, double dy, int xm, int ym, int *Niter, int *kernel_width);
int orc_smooth_surface_fluxes(int Nz1, int Nx, int Ny, double dx, double dy, double radius, double *edn, double *eup);

/* introspection used by the parity tests */
const double *orc_coeff_ptr(orc_solver *s, int which); /* 1 dir2dir, 2 dir2diff, 3 diff2diff: (ncoef,Nz,Nx,Ny) */
const double *orc_state_ptr(orc_solver *s, int uid, int which); /* 1 edir (Nd,Nz+1,Nx,Ny), 2 ediff, 3 abso, 4 b */
const int *orc_l1d_ptr(orc_solver *s);
void orc_get_info(orc_solver *s, int uid, int *niter_dir, int *niter_diff, double *res_dir, double *res_diff);
void orc_get_suninfo(orc_solver *s, double *phi, double *theta, double *symmetry_phi, int *xinc, int *yinc);
/* y = (I - T) x on interior arrays (Nf,Nz+1,Nx,Ny): src/pprts_shell.F90:366-541 */
void orc_mat_mult_ediff(orc_solver *s, const double *x, double *y);
/* one scale_flx direction on a raw (dof,Nz+1,Nx,Ny) array; ldir selects direct/diffuse scale vector */
void orc_scale_flx_arr(orc_solver *s, double *arr, int ldir, int lWm2);

#ifdef __cplusplus
}
#endif
#endif
