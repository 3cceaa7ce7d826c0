/*
 * pprts_b200.h -- C ABI of libpprts_b200.so: the TenStream pprts solve path on B200 (sm_100a).
 *
 * Two layers, both plain C (pointers + sizes, host memory unless stated otherwise):
 *
 *  1. pprts_f2c_*  : the entry points the reference's own C wrapper exports (c_wrapper/f2c_pprts.h:61-99,
 *                    implemented in c_wrapper/f2c_pprts.F90).  Same names, same argument meaning, same "abort on
 *                    error" behaviour, so c_wrapper/pprts.c links against this library unchanged.
 *  2. pprts_b200_* : a handle-based API one level lower, which is what a Fortran ISO_C_BINDING shim binds to
 *                    (fortran/m_pprts_b200.F90).  It mirrors the Fortran module API of src/pprts.F90:192-200:
 *                    init_pprts / set_angles / set_optical_properties / solve_pprts / pprts_get_result /
 *                    destroy_pprts.  Every function returns 0 on success; pprts_b200_last_error() has the text.
 *
 * Array conventions of layer 2 are the reference's Fortran ones: 3-D fields are (Nz, Nx, Ny) column-major, i.e.
 * z fastest in memory, index 0 = top of atmosphere; level fields are (Nz+1, Nx, Ny); 2-D fields (Nx, Ny).
 * There is no CPU fallback: every call that computes needs a CUDA device.
 */
#ifndef PPRTS_B200_H
#define PPRTS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------
 * Layer 1: drop-in for c_wrapper/f2c_pprts.h
 * ------------------------------------------------------------------------------------------------------------ */
#define SOLVER_ID_PPRTS_3_10 310 /* c_wrapper/f2c_solver_ids.h */
#define SOLVER_ID_PPRTS_3_24 324
#define SOLVER_ID_PPRTS_8_10 810

/* c_wrapper/f2c_pprts.F90:130-284.  fcomm is the caller's Fortran MPI communicator handle; inside one box the
 * library runs one process per GPU and does not use it.  Scalars are in/out like the reference (rank 0's values
 * are "broadcast back").  hhl: Nz+1 level heights [m], TOA first.  LUT files are located through the
 * environment variable PPRTS_B200_LUT_BASENAME (fallback LUT_BASENAME, default "./LUT"), the reference's
 * -lut_basename option (src/optprop_LUT.F90:503-518). */
void pprts_f2c_init(int fcomm, int *solver_id, int *Nz, int *Nx, int *Ny, double *dx, double *dy, float *hhl,
                    float *phi0, float *theta0, int *collapseindex);
/* c_wrapper/f2c_pprts.F90:286-323. arrays [Ny][Nx][Nz] (planck [Ny][Nx][Nz+1]) in C order; planck used iff any > 0 */
void pprts_f2c_set_global_optical_properties(int Nz, int Nx, int Ny, float *albedo, float *kabs, float *ksca,
                                             float *g, float *planck);
/* c_wrapper/f2c_pprts.F90:325-356: solar iff edirTOA > 0, thermal otherwise */
void pprts_f2c_solve(int fcomm, float edirTOA);
/* c_wrapper/f2c_pprts.F90:358-406: W/m2 on levels [Ny][Nx][Nz+1], abso W/m3 [Ny][Nx][Nz] */
void pprts_f2c_get_result(int Nz, int Nx, int Ny, float *edn, float *eup, float *abso, float *edir);
/* c_wrapper/f2c_pprts.F90:458-478: by VALUE in the implementation (the header's int* prototype is stale) */
void pprts_f2c_destroy(int lfinalizepetsc);

/* coefficient probe, c_wrapper/f2c_pprts.F90:591-758.  imode 1 = dir2dir, 2 = dir2diff, 3 = diff2diff;
 * coeff is [dst][src] float. */
void pprts_f2c_opp_init(const int comm, const int solver_id, void **opp, int *ierr);
void pprts_f2c_opp_get_coeff(void *opp, const float tauz, const float w0, const float g, const float aspect_zx,
                             const float phi, const float theta, const int imode, const int lswitch_east,
                             const int lswitch_north, const int Ncoeff, float *coeff, int *ierr);
void pprts_f2c_opp_destroy(void *opp, int *ierr);
void pprts_f2c_opp_get_info(void *opp, int *Ndir, int *Ndiff, float *diff_tauz_range, float *diff_w0_range,
                            float *diff_g_range, float *diff_aspect_zx_range, float *dir_tauz_range,
                            float *dir_w0_range, float *dir_g_range, float *dir_aspect_zx_range,
                            float *dir_phi_range, float *dir_theta_range, int *ierr);

/* Not in the reference: register an already opened table set for the module-level solver of the pprts_f2c_* layer
 * (tests use reduced parameter spaces; the reference always loads `-lut_basename`, src/tenstream_options.F90:108-110),
 * and replace the "print + abort" of CHKERR (src/helper_functions.fypp:888-906) by a callback. */
struct pprts_b200_lut;
void pprts_b200_f2c_set_lut(struct pprts_b200_lut *lut);
void pprts_b200_f2c_set_abort_handler(void (*handler)(const char *message));

/* ------------------------------------------------------------------------------------------------------------
 * Layer 2: handle-based API (what fortran/m_pprts_b200.F90 binds)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct pprts_b200_solver pprts_b200_solver;
typedef struct pprts_b200_lut pprts_b200_lut;

/* optprop tables (src/optprop_LUT.F90:186-330, :1336-1435): the three .mmap4 files are memory-mapped on the host;
 * ndir/axes are the six direct axes [tau, w0, aspect_zx, g, phi, theta] (the diffuse table uses the first four). */
int pprts_b200_lut_open(int solver_id, const char *path_Tdir, const char *path_Sdir, const char *path_Sdiff,
                        const int ndir[6], const float *const axes[6], pprts_b200_lut **lut);
int pprts_b200_lut_close(pprts_b200_lut *lut);

typedef struct {
  int solver_id;       /* 310, 810, 324 */
  int Nz, Nx, Ny;      /* local domain: layers, columns */
  double dx, dy;       /* must be equal (src/pprts.F90:459) */
  const double *dz1d;  /* (Nz) or NULL */
  const double *dz3d;  /* (Nz,Nx,Ny) or NULL; exactly one of dz1d / dz3d */
  double sundir[3];    /* cartesian, pointing away from the sun (init_pprts argument) */
  double twostr_ratio; /* -twostr_ratio, <= 0 -> 2.0 (src/tenstream_options.F90) */
  int coeff_fp32;      /* 1: store 3-D transport coefficients as fp32 (the LUT's own precision), 0: fp64 */
  int by_coeff_divergence; /* -absorption_by_coeff_divergence, default 1 (src/pprts.F90:5286) */
  double rtol, atol;   /* < 0: reference defaults (src/pprts_base.F90:1128-1131) */
  int maxit_dir;       /* <= 0: 1000 (src/pprts_explicit.F90:78) */
  int maxit_diff;      /* <= 0: 1000 (src/pprts_base.F90:1126) */
  int precond;         /* diffuse Krylov preconditioner: 0 none, 1 vertical-line (column) solve, 2 two-stream column only,
                        * 3 (default) column solve + coarse-grid correction on 4x4-column aggregates (3_10 / 8_10; other
                        * layouts fall back to 1) */
  int device;          /* CUDA ordinal, < 0: current device */
  int keep_solutions;  /* 1: cache edir/ediff/abso per uid for warm starts like solver%solutions(uid) */
  /* x/y decomposition over the GPUs of one box (src/pprts_base.F90:721-828): this rank's cartesian position.
   * nxproc = nyproc = 1 means a single periodic domain. */
  int nxproc, nyproc, xrank, yrank;
  /* -solar_diff_explicit / -thermal_diff_explicit (src/pprts.F90:2793-2800 -> explicit_ediff,
   * src/pprts_explicit.F90:461-713): stationary relaxation with the reference's ||x_n - x_{n-1}||_2 atol/rtol test
   * instead of the Krylov solve.  The reference's cell-by-cell SOR ordering is sequential; here every sweep relaxes
   * all columns at once (x += M^-1 (b - A x), M = the column operator of `precond`, plain Jacobi when precond = 0). */
  int diff_explicit;
  /* 1: mixed-precision Krylov solve -- fp32 work vectors inside BiCGStab, fp64 solution and fp64 true residual with
   * iterative refinement (the convergence test is always evaluated on an fp64 residual).  Halves the vector traffic
   * of the iteration; on a decomposed solver the halos travel as fp32 too.  Default 1 (pprts_b200_default_config). */
  int krylov_fp32;
  /* init_pprts collapseindex (src/pprts.F90:692-706, handle_atm_collapse :2080-2198): > 1 merges the top
   * `collapseindex` atmosphere layers (forced 1-D) into ONE 1-D layer by the adding method; Nz, dz and all optical
   * property arrays are then on the ATMOSPHERE grid (Nz layers) while the solver grid and every result array have
   * Nz - collapseindex + 1 layers.  <= 1: off. */
  int collapseindex;
  /* error-growth tracking of stored solutions (src/adaptive_spectral_integration.F90): -max_solution_err [W]
   * (default 5e3/86400, src/tenstream_options.F90:90) and -max_solution_time [s] (default 0).  Both > 0 enables
   * lenable_solutions_err_estimates (src/pprts.F90:456-457). */
  double max_solution_err, max_solution_time;
  /* -pprts_geometric_coeffs (src/pprts.F90:3046-3053): the direct-beam coefficients of every 3-D cell come from its
   * geometry (src/geometric_coeffs.F90; sloped top/bottom faces of a dz3d terrain-following grid, planar faces and
   * parallel planes as in the reference's defaults) instead of the LUT, and dir2diff is rescaled to conserve the LUT's
   * absorption (-pprts_conserve_lut_atm_abso, :3324-3351).  3 direct streams only (3_10, 3_24).  Default 0.
   * Bit flags on top of 1 select the non-default vertex variants of init_vertices (:3057-3085): + 2 =
   * -pprts_top_bottom_faces_planar no (the fourth vertex of a face keeps its own height), + 4 =
   * -pprts_top_bottom_planes_parallel no (the top face takes the level heights instead of bottom + dz). */
  int geometric_coeffs;
  /* -accept_incomplete_solve (src/pprts.F90:4270-4275, src/pprts_explicit.F90:134-136, :213-217): 1 = a solve that
   * hits its iteration limit returns its last iterate instead of an error.  Default 0: a failed diffuse Krylov solve is
   * first retried from a zero initial guess with restarted GMRES, as the reference does (src/pprts.F90:4277-4296), and
   * only then reported (error 22). */
  int accept_incomplete_solve;
  int gmres_restart;   /* restart length of that GMRES, <= 0: 30 (PETSc's default) */
  /* keep_solutions = 1 only: at most this many solutions stay cached (least recently used dropped first, never the
   * uid being solved or uid - 1); <= 0: unlimited, like solver%solutions(-1000:1000) */
  int max_cached_solutions;
  /* -pprts_open_bc (src/pprts.F90:462-463): open instead of cyclic boundaries for the DIRECT beam -- it enters through
   * the sunward domain edges with the side fluxes of a single-column solve of the edge columns (set_open_bc /
   * single_column_solve, src/pprts_base.F90:1189-1458) and leaves through the others without re-entering
   * (exchange_direct_boundary sends zeros across the edge, src/pprts_explicit.F90:262-306; the sweep takes the inflow,
   * :358-378).  As there, the stored solution keeps zeros on the inflow faces it owns, and the source term, absorption
   * and the diffuse solve go on treating the stored beam periodically.  The boundary columns always take the LUT's
   * dir2dir, also with geometric_coeffs or buildings (get_coeff in single_column_solve, :1368-1382). */
  int open_bc;
} pprts_b200_config;

void pprts_b200_default_config(pprts_b200_config *cfg);

/* init_pprts (src/pprts.F90:213-553) */
int pprts_b200_create(const pprts_b200_config *cfg, const pprts_b200_lut *lut, pprts_b200_solver **solver);
/* destroy_pprts (src/pprts_base.F90:877) */
int pprts_b200_destroy(pprts_b200_solver *solver);
/* set_angles (src/pprts.F90:1100-1183) */
int pprts_b200_set_angles(pprts_b200_solver *solver, const double sundir[3]);
/* set_optical_properties (src/pprts.F90:1764-2076). planck, planck_srfc, albedo_2d may be NULL. */
int pprts_b200_set_optical_properties(pprts_b200_solver *solver, double albedo, const double *kabs,
                                      const double *ksca, const double *g, const double *planck,
                                      const double *planck_srfc, const double *albedo_2d, int ldelta_scaling);
/* solve_pprts (src/pprts.F90:2487-2665). uid in [-1000, 1000]. */
int pprts_b200_solve(pprts_b200_solver *solver, int lthermal, int lsolar, double edirTOA, int uid);
/* solve_pprts(..., opt_solution_time = time) (src/pprts.F90:2487-2493, restore_solution :3991-4049): as
 * pprts_b200_solve, and records `time` and -- with the error estimates enabled -- the inf-norm of the change of the
 * absorption in the solution's history */
int pprts_b200_solve_at_time(pprts_b200_solver *solver, int lthermal, int lsolar, double edirTOA, int uid, double time);
/* need_new_solution(comm, solver%solutions(uid), time, solver%lenable_solutions_err_estimates)
 * (src/adaptive_spectral_integration.F90:38-172), the test the band loops make before every solve
 * (rrtmg/rrtmg/pprts_rrtmg.F90:623, :1001).  have_time = 0: `time` not present.  *need: 1 = solve again. */
int pprts_b200_need_new_solution(pprts_b200_solver *solver, int uid, double time, int have_time, int *need);
/* polyfit (src/adaptive_spectral_integration.F90:174-257): least-squares polynomial of degree d through n points;
 * 0 on success, 1 for equal abscissae, 2 for a singular system.  Pure host code (no device needed). */
int pprts_b200_polyfit(const double *vx, const double *vy, int n, int d, double *coef);
/* solution%time(30), solution%maxnorm(30) (src/pprts_base.F90:161-162), newest first */
int pprts_b200_get_solution_times(pprts_b200_solver *solver, int uid, double time[30], double maxnorm[30]);
/* pprts_get_result (src/pprts.F90:5799-5917): caller-allocated, edir may be NULL */
int pprts_b200_get_result(pprts_b200_solver *solver, int uid, double *edn, double *eup, double *abso,
                          double *edir);

/* set_optical_properties (src/pprts.F90:1764-2076) split into its host->device copy and its device part, and
 * pprts_get_result (src/pprts.F90:5799-5917) without the device->host copy -- used for throughput measurement with inputs
 * already in HBM: stage() uploads a set of optical properties into slot (untimed), set_optical_properties_staged()
 * consumes it on the device, and compute_result_device() runs the result reduction only. */
int pprts_b200_stage_optical_properties(pprts_b200_solver *solver, int slot, double albedo, const double *kabs,
                                        const double *ksca, const double *g, const double *planck,
                                        const double *planck_srfc, const double *albedo_2d);
int pprts_b200_set_optical_properties_staged(pprts_b200_solver *solver, int slot, int ldelta_scaling);
/* Overlapped host transfers for a caller's own band loop (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053; what
 * pprts_b200_solve_bands does internally), the same two reference routines with their copies made asynchronous:
 * stage_optical_properties_async only ENQUEUES the upload of the next band on a copy stream (the host arrays must stay
 * valid, and should be page-locked, until set_optical_properties_staged(slot) has consumed them); get_result_async
 * returns once the result is on its way to the host arrays, which are complete after wait_results. */
int pprts_b200_stage_optical_properties_async(pprts_b200_solver *solver, int slot, double albedo, const double *kabs,
                                              const double *ksca, const double *g, const double *planck,
                                              const double *planck_srfc, const double *albedo_2d);
int pprts_b200_get_result_async(pprts_b200_solver *solver, int uid, double *edn, double *eup, double *abso,
                                double *edir);
int pprts_b200_wait_results(pprts_b200_solver *solver);
int pprts_b200_compute_result_device(pprts_b200_solver *solver, int uid);
int pprts_b200_synchronize(pprts_b200_solver *solver);

/* spectral accumulation for band sharding: acc += weight * result(uid) on the device, read back once -- the sum the
 * callers keep on the host after every pprts_get_result (rrtmg/rrtmg/pprts_rrtmg.F90:1048-1051, :702-706) */
int pprts_b200_accumulate_result(pprts_b200_solver *solver, int uid, double weight);
int pprts_b200_reset_accumulator(pprts_b200_solver *solver);
int pprts_b200_get_accumulated(pprts_b200_solver *solver, double *edn, double *eup, double *abso, double *edir);

/* Band loop of the spectral callers in one call (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053 solar, :621-708 thermal;
 * specint/specint_pprts.F90:89-213): for ib in 0..nbands-1
 *     set_optical_properties(albedo, kabs(:,:,:,ib), ksca(:,:,:,ib), g(:,:,:,ib), planck(:,:,:,ib), planck_srfc(:,:,ib))
 *     solve_pprts(lthermal = !lsolar, lsolar, edirTOA(ib), uid = uid0 + ib)
 *     edn, eup, abso, edir += weights(ib) * pprts_get_result(uid0 + ib)
 * The per-band host arrays never come back: the optical properties of band ib+1 are uploaded on a copy stream while
 * band ib is being solved (pass page-locked memory for the copy to be asynchronous), the spectral sum is kept on
 * the device and read back once.  Fields are (Nz,Nx,Ny[,nbands]) column-major like set_optical_properties; a band
 * stride of 0 elements re-uses the same field for every band (e.g. ksca = g = 0 in the thermal loop). */
typedef struct {
  int nbands;
  int lsolar;                 /* 1: solar bands, 0: thermal bands */
  int uid0;                   /* solution uid of band 0 (reference: ib for solar, 500 + ib for thermal) */
  int ldelta_scaling;
  double albedo;
  const double *albedo_2d;    /* (Nx,Ny) or NULL */
  const double *kabs, *ksca, *g;                       /* band ib at ptr + ib * stride */
  long stride_kabs, stride_ksca, stride_g;             /* in elements; 0: shared by all bands */
  const double *planck, *planck_srfc;                  /* thermal: (Nz+1,Nx,Ny[,nbands]), (Nx,Ny[,nbands]) or NULL */
  long stride_planck, stride_planck_srfc;
  const double *edirTOA;      /* (nbands), solar */
  const double *weights;      /* (nbands) or NULL: 1 */
  int accumulate_only;        /* 1: add to the device accumulator without resetting it first */
  /* opt_time of the spectral callers (rrtmg/rrtmg/pprts_rrtmg.F90:1001, :623): with have_time = 1 every band is first
   * asked need_new_solution(uid, time); bands that do not need one are not uploaded or solved, their stored solution
   * enters the sum (niter_diff = -1), the others are solved with opt_solution_time = time.  Needs keep_solutions. */
  int have_time;
  double time;
} pprts_b200_bands;
/* the loop rrtmg/rrtmg/pprts_rrtmg.F90:999-1053 / :621-708 itself.  edn/eup/edir (Nz+1,Nx,Ny), abso (Nz,Nx,Ny): spectral
 * sums, any may be NULL (NULL for all four leaves the sums in the device accumulator for
 * pprts_b200_allreduce_accumulated / pprts_b200_get_accumulated); niter_diff (nbands) or NULL */
int pprts_b200_solve_bands(pprts_b200_solver *solver, const pprts_b200_bands *bands, double *edn, double *eup,
                           double *abso, double *edir, int *niter_diff);

/* NCCL plumbing (one process per GPU), in place of the MPI communicator and the DMDA of init_pprts
 * (setup_grid / setup_coord_native, src/pprts_base.F90:721-828): the unique id is produced on rank 0 and distributed
 * by the caller (MPI_Bcast in a Fortran host, torch.distributed in the tests). */
int pprts_b200_nccl_unique_id(char id[128]);
int pprts_b200_comm_init(pprts_b200_solver *solver, int nranks, int rank, const char id[128]);
/* sum of the accumulated spectral integrals over all ranks (final flux / heating-rate reduction): with the bands sharded
 * over the processes this completes the callers' host-side sum (rrtmg/rrtmg/pprts_rrtmg.F90:1048-1051) */
int pprts_b200_allreduce_accumulated(pprts_b200_solver *solver);

/* slope_correction_fluxes(solver, edir) (src/pprts_postprocess.F90:130-196), the spectral callers' post-processing of
 * the direct flux on a terrain-following grid: lslope = -pprts_slope_correction (surface level x max(0, sundir.n /
 * sundir.ez), n the surface normal from hgrad, src/pprts.F90:621-638), latm = -pprts_atm_correction (every level x
 * sundir.ez / sundir.n).  edir: host (Nz+1,Nx,Ny) column-major, in place.  Needs dz3d; single periodic domain. */
int pprts_b200_slope_correction_fluxes(pprts_b200_solver *solver, int lslope, int latm, double *edir);
/* the same applied to the direct flux of the device-side spectral sum (before pprts_b200_get_accumulated), where the
 * callers apply it: after the band sum (rrtmg/rrtmg/pprts_rrtmg.F90:320) */
int pprts_b200_slope_correct_accumulated(pprts_b200_solver *solver, int lslope, int latm);

/* smooth_surface_fluxes(solver, edn, eup) (src/pprts_postprocess.F90:44-84), the callers' post-processing of the diffuse
 * surface fluxes (rrtmg/rrtmg/pprts_rrtmg.F90:319); radius = the value of -pprts_smooth_srfc_flx [m] (the caller skips
 * the call when the option is absent).  radius < 0: the surface level of edn and eup becomes its domain mean;
 * radius >= 0: Niter passes of the separable periodic box blur of half width kernel_width (convolve_ediff_srfc,
 * src/pprts_base.F90:1733-1868), both from find_iter_and_kernelwidth (src/pprts_postprocess.F90:86-128, pure host
 * arithmetic: no solver, no GPU).  edn, eup: host (Nz+1,Nx,Ny) column-major, in place.  Single periodic domain. */
int pprts_b200_find_iter_and_kernelwidth(double radius, double dx, double dy, int xm, int ym, int *Niter,
                                         int *kernel_width);
int pprts_b200_smooth_surface_fluxes(pprts_b200_solver *solver, double radius, double *edn, double *eup);
/* the same applied to edn / eup of the device-side spectral sum (before pprts_b200_get_accumulated), where the callers
 * apply it: after the band sum (rrtmg/rrtmg/pprts_rrtmg.F90:319) */
int pprts_b200_smooth_accumulated(pprts_b200_solver *solver, double radius);

/* ---- buildings: the opt_buildings argument of solve_pprts / pprts_get_result ------------------------------------
 * Replaces type(t_pprts_buildings) (src/buildings.F90:40-69) as consumed by alloc_coeff_dir2dir / alloc_coeff_diff2diff
 * (src/pprts.F90:3194-3212, :3579-3677), setup_b (:4989-5145) and pprts_get_result (:6011-6160).
 * iface[m]: 1-based flattened index of face m in a (6, Nz, Nx, Ny) array, first axis fastest
 * (faceidx_by_cell_plus_offset, src/buildings.F90:215-222); face ids 1..6 = PPRTS_TOP_FACE, BOT, LEFT (x+0),
 * RIGHT (x+1), REAR (y+0), FRONT (y+1) (src/boxmc_geometry.F90:45-51).  albedo[m]: Lambertian albedo of the face;
 * planck[m]: emitted Planck radiance (NULL: no emission).  The list stays attached until replaced; nfaces = 0
 * removes it.  Periodic single-GPU domains only (error 60 on a decomposed solver). */
int pprts_b200_set_buildings(pprts_b200_solver *solver, int nfaces, const int *iface, const double *albedo,
                             const double *planck);
/* fill_buildings_arr of pprts_get_result(opt_buildings) (src/pprts.F90:6011-6160).  Per face, W/m2: direct flux on the
 * face (0 for thermal solutions; may be NULL), diffuse flux entering the building through it (incoming) and leaving it
 * (outgoing) */
int pprts_b200_get_buildings_result(pprts_b200_solver *solver, int uid, double *edir, double *incoming,
                                    double *outgoing);

/* -pprts_set_abso_in_buildings <val> (src/pprts.F90:5986-6009, the last step of pprts_get_result(opt_buildings)): the
 * stored absorption of the cells that carry a building face becomes val; later pprts_b200_get_result calls for this uid
 * return it (x mu for a solar solution), results already fetched are unaffected.  Call it after get_result, as the
 * reference does. */
int pprts_b200_set_abso_in_buildings(pprts_b200_solver *solver, int uid, double val);

/* introspection (parity tests, profiling): the members of t_solver the reference's own tests read -- solver%dir2dir /
 * dir2diff / diff2diff (src/pprts_base.F90:92-268), solutions(uid)%edir / ediff / abso, solver%b, the iteration counts and
 * residuals of the KSP history (src/pprts.F90:4206-4303) -- and the matrix-free operator op_mat_mult_ediff
 * (src/pprts_shell.F90:366-541) applied to a host vector */
/* which: 1 dir2dir, 2 dir2diff, 3 diff2diff -> out (ncoef, Nz, Nx, Ny) column-major, zeros in 1-D layers */
int pprts_b200_get_coeffs(pprts_b200_solver *solver, int which, double *out);
/* which: 1 edir (Nd,Nz+1,Nx,Ny), 2 ediff (Nf,Nz+1,Nx,Ny), 3 abso (Nz,Nx,Ny), 4 b (Nf,Nz+1,Nx,Ny); W/m2 resp. W */
int pprts_b200_get_state(pprts_b200_solver *solver, int uid, int which, double *out);
int pprts_b200_get_info(pprts_b200_solver *solver, int uid, int *niter_dir, int *niter_diff, double *res_dir,
                        double *res_diff);
/* y = (I - T) x with the current coefficients, x and y (Nf,Nz+1,Nx,Ny) column-major host arrays */
int pprts_b200_mat_mult_ediff(pprts_b200_solver *solver, const double *x, double *y);
/* fp64 coefficient lookup for n sample cells on the device (pprts_base.F90:1491 get_coeff): mode 1/2/3,
 * out (ncoef, n) column-major */
int pprts_b200_lookup_coeffs(pprts_b200_solver *solver, int mode, long n, const double *kabs, const double *ksca,
                             const double *g, const double *dz, double *out);
/* counters: kernels launched by this library since creation, and the ids of the last timed kernels (the role of the
 * reference's PETSc log stages, t_solver_log_events, src/pprts_base.F90:176-200) */
long pprts_b200_kernel_launches(pprts_b200_solver *solver);
/* the cudaStream_t every kernel of this solver is launched on (to record timing events on the launching stream) */
void *pprts_b200_stream(pprts_b200_solver *solver);
/* device time [ms] and launch count accumulated per kernel family since the last reset:
 * 0 lookup, 1 edir sweep, 2 source, 3 matmult, 4 precond, 5 vector ops, 6 absorption+result, 7 ingest/egress,
 * 8 halo pack/unpack; entries 9..15 are reserved (zero) */
int pprts_b200_kernel_stats(pprts_b200_solver *solver, int reset, double ms[16], long launches[16]);
int pprts_b200_enable_kernel_timing(pprts_b200_solver *solver, int on);
/* how often the GMRES retry of a failed diffuse solve ran since creation (src/pprts.F90:4277-4296), and a fault
 * injection hook for testing it: the next n primary Krylov solves report "did not converge" without iterating */
long pprts_b200_krylov_retries(pprts_b200_solver *solver);
int pprts_b200_debug_fail_next_krylov(pprts_b200_solver *solver, int n);

const char *pprts_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
