// solver.cu -- host side of libpprts_b200.so: the handle-based C ABI of include/pprts_b200.h.
// Mirrors the Fortran module API of the reference (src/pprts.F90:192-200); the orchestration below follows
// solve_pprts -> pprts() -> restore_solution (src/pprts.F90:2487-3033, :3991-4076) step by step, every heavy step
// being one of the kernels in kernels.cuh.  There is no CPU fallback anywhere in this file.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <type_traits>
#include <string>
#include <array>
#include <vector>

#include "../../include/pprts_b200.h"
#include "kernels.cuh"
#include "internal.h"

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char *pprts_b200_last_error(void) { return g_err; }

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return set_err(100 + (int)e_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// LUT (host side): memory-mapped .mmap4 files (src/mmap.F90:129-203)
// ---------------------------------------------------------------------------------------------------------------
struct pprts_b200_lut {
  int solver_id, Nd, Nf;
  int n[6];
  std::vector<float> ax[6];
  long off[6];
  const float *Tdir = nullptr, *Sdir = nullptr, *Sdiff = nullptr;
  void *map[3] = {nullptr, nullptr, nullptr};
  size_t maplen[3] = {0, 0, 0};
};

static int stream_counts(int solver_id, int &Nd, int &Nf) {
  switch (solver_id) {
    case 310: Nd = L310::Nd; Nf = L310::Nf; return 0;
    case 810: Nd = L810::Nd; Nf = L810::Nf; return 0;
    case 324: Nd = L324::Nd; Nf = L324::Nf; return 0;
  }
  return 1;
}

static int mmap4_open(const char *path, const float **data, long *ncoef, long *nlat, void **map, size_t *maplen) {
  long pagesize = sysconf(_SC_PAGESIZE);
  size_t nhdr = (size_t)pagesize / sizeof(size_t);
  std::vector<size_t> header(nhdr);
  int fd = open(path, O_RDONLY);
  if (fd < 0) return set_err(2, "cannot open LUT file %s", path);
  if (read(fd, header.data(), nhdr * sizeof(size_t)) != (ssize_t)(nhdr * sizeof(size_t))) {
    close(fd);
    return set_err(2, "short read on %s", path);
  }
  if (header[0] != sizeof(float)) { close(fd); return set_err(2, "%s: dtype size %zu != 4", path, header[0]); }
  int ndim = 0;
  for (size_t i = 3; i < nhdr; ++i) ndim += header[i] != 0;
  if (ndim != 2) { close(fd); return set_err(2, "%s: expected a 2-d table", path); }
  *ncoef = (long)header[3];
  *nlat = (long)header[4];
  // header consistency (src/mmap.F90:85-93: [dtype_size, n_elements, n_bytes, dims...]) and the real file size: a
  // truncated table would otherwise only show up as a SIGBUS in the middle of a lookup
  struct stat stt;
  if (header[1] != header[3] * header[4] || header[2] != header[1] * sizeof(float) || fstat(fd, &stt) != 0 ||
      (size_t)stt.st_size < header[2] + (size_t)pagesize) {
    close(fd);
    return set_err(2, "%s: header (%zu elements, %zu bytes, %zu x %zu) does not match the file size", path, header[1],
                   header[2], header[3], header[4]);
  }
  *maplen = header[2] + (size_t)pagesize;
  *map = mmap(0, *maplen, PROT_READ, MAP_PRIVATE | MAP_NORESERVE, fd, 0);
  close(fd);
  if (*map == MAP_FAILED) return set_err(2, "mmap failed on %s", path);
  *data = (const float *)((const char *)(*map) + pagesize);
  return 0;
}

extern "C" int pprts_b200_lut_open(int solver_id, const char *pT, const char *pS, const char *pD, const int ndir[6],
                                   const float *const axes[6], pprts_b200_lut **out) {
  auto *lut = new pprts_b200_lut();
  lut->solver_id = solver_id;
  if (stream_counts(solver_id, lut->Nd, lut->Nf)) { delete lut; return set_err(1, "unknown solver id %d", solver_id); }
  long nlat_dir = 1, nlat_diff = 1;
  for (int d = 0; d < 6; ++d) {
    lut->n[d] = ndir[d];
    lut->ax[d].assign(axes[d], axes[d] + ndir[d]);
    for (int q = 1; q < ndir[d]; ++q)
      if (!(axes[d][q] > axes[d][q - 1])) { delete lut; return set_err(1, "LUT axis %d is not strictly ascending", d); }
    lut->off[d] = d == 0 ? 1 : lut->off[d - 1] * ndir[d - 1];
    nlat_dir *= ndir[d];
    if (d < 4) nlat_diff *= ndir[d];
  }
  long nc, nl;
  int rc;
  if (pT && pS) {
    if ((rc = mmap4_open(pT, &lut->Tdir, &nc, &nl, &lut->map[0], &lut->maplen[0]))) { delete lut; return rc; }
    if (nc != lut->Nd * lut->Nd || nl != nlat_dir) { delete lut; return set_err(3, "%s: shape (%ld,%ld) does not match the axes", pT, nc, nl); }
    if ((rc = mmap4_open(pS, &lut->Sdir, &nc, &nl, &lut->map[1], &lut->maplen[1]))) { delete lut; return rc; }
    if (nc != lut->Nd * lut->Nf || nl != nlat_dir) { delete lut; return set_err(3, "%s: shape (%ld,%ld) does not match the axes", pS, nc, nl); }
  }
  if ((rc = mmap4_open(pD, &lut->Sdiff, &nc, &nl, &lut->map[2], &lut->maplen[2]))) { delete lut; return rc; }
  if (nc != lut->Nf * lut->Nf || nl != nlat_diff) { delete lut; return set_err(3, "%s: shape (%ld,%ld) does not match the axes", pD, nc, nl); }
  *out = lut;
  return 0;
}
extern "C" int pprts_b200_lut_close(pprts_b200_lut *lut) {
  if (!lut) return 0;
  for (int d = 0; d < 3; ++d) if (lut->map[d]) munmap(lut->map[d], lut->maplen[d]);
  delete lut;
  return 0;
}

// search_sorted_bisection on the host (src/search.fypp:177-232), fp32 axis, fp64 arithmetic
static double h_find_real_location(const std::vector<float> &arr, float val) {
  int i = 1, j = (int)arr.size();
  const double v = (double)val;
  for (;;) {
    int k = (i + j) / 2;
    if (v < (double)arr[k - 1]) j = k; else i = k;
    if (i + 1 >= j) {
      double inc = (i == j) ? 0.0 : (v - (double)arr[i - 1]) / ((double)arr[j - 1] - (double)arr[i - 1]);
      return std::min(std::max(1.0, (double)i + inc), (double)arr.size());
    }
  }
}

// stream permutation gather maps (src/optprop.F90:986-1301): new[n] = old[map[n]]
static void apply_blocks(std::vector<short> &m, int Nsrc, const std::vector<std::pair<int, int>> &blocks,
                         const int *srcperm) {
  std::vector<short> bak = m;
  for (auto &b : blocks)
    for (int s = 0; s < Nsrc; ++s) m[b.first * Nsrc + s] = bak[b.second * Nsrc + (srcperm ? srcperm[s] : s)];
}
static void build_sym_maps(int solver_id, bool le, bool ln, SymMaps &sym) {
  int Nd, Nf;
  stream_counts(solver_id, Nd, Nf);
  std::vector<short> d2d(Nd * Nd), d2f(Nd * Nf);
  for (int n = 0; n < Nd * Nd; ++n) d2d[n] = (short)n;
  for (int n = 0; n < Nd * Nf; ++n) d2f[n] = (short)n;
  static const int pe[8] = {1, 0, 3, 2, 4, 5, 6, 7}, pn[8] = {2, 3, 0, 1, 4, 5, 6, 7};
  if (solver_id == 310) {  // :1009-1045
    if (le) apply_blocks(d2f, Nd, {{2, 3}, {3, 2}, {4, 5}, {5, 4}}, nullptr);
    if (ln) apply_blocks(d2f, Nd, {{6, 7}, {7, 6}, {8, 9}, {9, 8}}, nullptr);
  } else if (solver_id == 810) {  // :1148-1184 (dir2diff), :1268-1301 (dir2dir)
    if (le) apply_blocks(d2f, Nd, {{2, 3}, {3, 2}, {4, 5}, {5, 4}}, pe);
    if (ln) apply_blocks(d2f, Nd, {{6, 7}, {7, 6}, {8, 9}, {9, 8}}, pn);
    if (le) apply_blocks(d2d, Nd, {{0, 1}, {1, 0}, {2, 3}, {3, 2}, {4, 4}, {5, 5}, {6, 6}, {7, 7}}, pe);
    if (ln) apply_blocks(d2d, Nd, {{0, 2}, {1, 3}, {2, 0}, {3, 1}, {4, 4}, {5, 5}, {6, 6}, {7, 7}}, pn);
  } else if (solver_id == 324) {  // :1098-1124
    if (le) apply_blocks(d2f, Nd, {{0, 2}, {2, 0}, {4, 6}, {6, 4}, {1, 3}, {3, 1}, {5, 7}, {7, 5},
                                   {9, 8}, {11, 10}, {13, 12}, {15, 14}, {8, 9}, {10, 11}, {12, 13}, {14, 15}}, nullptr);
    if (ln) apply_blocks(d2f, Nd, {{0, 4}, {2, 6}, {4, 0}, {6, 2}, {1, 5}, {3, 7}, {5, 1}, {7, 3},
                                   {17, 16}, {19, 18}, {21, 20}, {23, 22}, {16, 17}, {18, 19}, {20, 21}, {22, 23}}, nullptr);
  }
  memset(&sym, 0, sizeof(sym));
  for (int n = 0; n < Nd * Nd; ++n) sym.d2d[n] = d2d[n];
  for (int n = 0; n < Nd * Nf; ++n) sym.d2f[n] = d2f[n];
}

// ---------------------------------------------------------------------------------------------------------------
// solver object
// ---------------------------------------------------------------------------------------------------------------
struct Solution {  // t_state_container (src/pprts_base.F90:137-169)
  bool lset = false, lsolar_rad = false, lthermal_rad = false, lchanged = true, lWm2_dir = true, lWm2_diff = true;
  double *edir = nullptr, *ediff = nullptr, *abso = nullptr;
  bool first_dir = true, first_diff = true;  // residual_history(1) < 0
  bool ediff_zero = false;  // ediff is exactly zero (fresh container, no warm start): A x needs no MatMult
  long last_use = 0;  // max_cached_solutions: the least recently used stored solution is dropped first
  int niter_dir = 0, niter_diff = 0;
  double res_dir = 0, res_diff = 0;
  // error-growth tracking (src/pprts_base.F90:161-162): times of the last solves, inf-norm of the change of abso
  double time[30], maxnorm[30];
  Solution() { for (int k = 0; k < 30; ++k) { time[k] = -1.0; maxnorm[k] = 0.0; } }
};

enum { FAM_LOOKUP = 0, FAM_EDIR, FAM_SOURCE, FAM_MATMULT, FAM_PRECOND, FAM_VEC, FAM_ABSO, FAM_IO, FAM_HALO, FAM_COUNT };

struct Staged {
  double albedo = 0;
  double *kabs = nullptr, *ksca = nullptr, *g = nullptr, *planck = nullptr, *Bsrfc = nullptr, *albedo2d = nullptr;
  bool have_planck = false, have_Bsrfc = false, have_albedo2d = false;
  cudaEvent_t ready = nullptr;  // stage_optical_properties_async: recorded on the copy stream behind the slot's uploads
  cudaEvent_t consumed = nullptr;  // set_optical_properties_staged: recorded on the solver's stream behind the ingest kernels
};

struct pprts_b200_solver {
  pprts_b200_config cfg;
  const pprts_b200_lut *lut = nullptr;
  int device = 0;
  Grid g;
  int Nd = 0, Nf = 0, n1d = 0;
  // atmosphere collapse (src/pprts_base.F90:118-122, atmk :1088-1093): the solver grid g starts at atmosphere layer
  // icollapse-1.  The atmosphere fields are allocated with all Nz_atm layers (atm_*); dz/kabs/ksca/gg/planck below point
  // at layer icollapse-1 of them (k is the slowest index of the device layout, so the offset is a pointer shift).
  int Nz_atm = 0, icollapse = 1;
  long ncell_atm = 0, nlev_atm = 0;
  double *atm_dz = nullptr, *atm_kabs = nullptr, *atm_ksca = nullptr, *atm_gg = nullptr, *atm_planck = nullptr;
  double *Btop = nullptr, *Bbot = nullptr;
  double unconstrained_fraction = 1.0;
  std::vector<unsigned char> l1d, l1d_atm;
  // sun (t_suninfo, src/pprts_base.F90:128-135)
  double phi = 0, theta = 0, mu = 0, costheta = 0, symmetry_phi = 0;
  int xinc = 0, yinc = 0;
  bool angles_dirty = true, coeffs_dirty = true, coeffs_have_dir = false;
  // device state
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t band_uploaded[2] = {nullptr, nullptr};
  double *dz = nullptr, *kabs = nullptr, *ksca = nullptr, *gg = nullptr, *planck = nullptr, *Bsrfc = nullptr,
         *albedo = nullptr, *raw = nullptr;
  bool have_planck = false, have_Bsrfc = false, have_optprop = false;
  double *a11 = nullptr, *a12 = nullptr, *a13 = nullptr, *a23 = nullptr, *a33 = nullptr;
  void *d2d = nullptr, *d2f = nullptr, *f2f = nullptr;
  double *cdiv_dir = nullptr, *cdiv_diff = nullptr;
  unsigned int *tile_class = nullptr;  // per coefficient tile: widening class of its diff2diff entries (common.cuh)
  int mm_variant = 2;  // 0: 2 stages + register prefetch, 1: 2 stages, 2: 1 stage x 4 CTAs/SM (fastest measured)
  double *b = nullptr, *kr[8] = {nullptr}, *edir_tmp = nullptr;
  // CUDA graphs of one mixed-precision BiCGStab iteration (first / odd / even iteration of a cycle) for small domains,
  // where the ~20 launches of an iteration cost more host time than GPU time
  struct IterGraph { cudaGraphExec_t exec = nullptr; int key = -1; long launches = 0; long fam[FAM_COUNT] = {0}; };
  IterGraph iter_graph[3];
  bool iter_graph_off = false;
  double *blur_buf[4] = {nullptr, nullptr, nullptr, nullptr};  // smooth_surface_fluxes on a decomposed solver: send lo/hi, halo lo/hi
  size_t blur_cap = 0;
  double *bcx = nullptr, *bcy = nullptr;  // -pprts_open_bc: boundary inflow of the sunward edge columns [d][k][j] / [d][k][i]
  double *precR = nullptr, *precInv = nullptr;
  double *vh_atm = nullptr, *vh = nullptr;  // -pprts_geometric_coeffs: vertex heights on the atmosphere / solver levels
  double *vh_ghost[4] = {nullptr, nullptr, nullptr, nullptr};  // decomposed: level heights of the neighbours' edge columns (slope correction)
  bool have_dz3d = false;                   // init_pprts(dz3d=...): atm%hgrad exists (src/pprts.F90:621)
  float *precRf = nullptr, *precInvf = nullptr;  // fp32 copies for the sweeps on fp32 vectors (krylov_fp32, 3_24 block form)
  void *precPU = nullptr, *precPD = nullptr;     // half2 planes: packed coefficients of the fp32 column sweeps (kernels.cuh PrecPack)
  bool prec_ready = false;
  // coarse-grid correction (cfg.precond = 3): aggregates of cg.AG x cg.AG columns, see kernels.cuh
  struct CgLevel {
    CoarseGrid c{};
    float *coef = nullptr, *R = nullptr, *sA = nullptr, *sB = nullptr, *rbar = nullptr, *e = nullptr, *x = nullptr;
  };
  int ncgl = 0;       // number of aggregate levels in use (0: plain column solve)
  bool cg_off = false;  // apply the column solve alone (stationary relaxation needs a convergent splitting)
  CgLevel cgl[3];     // [0] is folded into the column sweeps, further (coarser) ones are added by k_cg_prolong
  int prec_block[6] = {0, 0, 0, 0, 0, 0};  // chosen block sizes of k_prec_up / k_prec_down (column_block)
  double *res[4] = {nullptr}, *acc[4] = {nullptr}, *out_stage = nullptr;
  bool res_valid = false;
  int res_uid = 0;
  double *partials = nullptr, *sc = nullptr;
  unsigned int *counter = nullptr;
  int *bad = nullptr;
  double *h_sc = nullptr, *h_sc_dev = nullptr;  // pinned + mapped: the device writes the scalars straight into it
  int *h_bad = nullptr, *h_bad_dev = nullptr;  // mapped, like h_sc
  // LUT on device
  float *d_ax[4] = {nullptr};
  double *d_Tdir4 = nullptr, *d_Sdir4 = nullptr;
  float *d_Sdiff = nullptr;
  LutDev lutdev;
  SymMaps sym;
  std::map<int, Solution> solutions;
  std::map<int, Staged> staged;
  // stats
  long launches = 0;
  bool timing = false;
  long fam_launch[FAM_COUNT] = {0};
  double fam_ms[FAM_COUNT] = {0};
  struct Ev { int fam; cudaEvent_t a, b; };
  std::vector<Ev> evs;
  std::vector<cudaEvent_t> evpool;
  // NCCL
  void *nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // x/y decomposition: neighbour ranks (rank = xi + yi*nxproc, periodic; src/pprts_base.F90:721-828), halo buffers
  bool decomposed = false;
  int nb_west = 0, nb_east = 0, nb_south = 0, nb_north = 0;
  double glob_ncol = 0;  // global number of columns (KSP tolerances use the global domain, pprts_base.F90:1128-1131)
  double *hs[2] = {nullptr, nullptr}, *hr[2] = {nullptr, nullptr};  // send / receive staging, [0] x faces, [1] y faces
  long halo_exchanges = 0;
  double *egress[4] = {nullptr, nullptr, nullptr, nullptr};  // get_result_async: results in the caller's layout
  cudaEvent_t egress_ready = nullptr, egress_done = nullptr;
  bool egress_pending = false;
  long use_clock = 0;
  std::vector<std::array<double *, 3>> spare;  // (edir, ediff, abso) of dropped solutions, reused by the next new uid
  int fail_next_krylov = 0;  // test hook: the next n primary Krylov solves report non-convergence
  long krylov_retries = 0;
  int num_sms = 148;
  bool use_tma = true;
  // opt_buildings (t_pprts_buildings, src/buildings.F90:40-69): faces of building cells, device copies
  int nbld = 0;
  bool bld_have_planck = false;
  int *bld_iface = nullptr;
  double *bld_albedo = nullptr, *bld_planck = nullptr, *bld_out = nullptr;  // bld_out: edir | incoming | outgoing
};

// ---------------------------------------------------------------------------------------------------------------
// NCCL entry points (dlopen'ed: single-GPU users do not need the library)
// ---------------------------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_ncclGetUniqueId)(nccl_uid_t *);
typedef int (*fn_ncclCommInitRank)(void **, int, nccl_uid_t, int);
typedef int (*fn_ncclAllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_ncclSend)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_ncclRecv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_ncclGroup)(void);
typedef int (*fn_ncclCommDestroy)(void *);
typedef const char *(*fn_ncclGetErrorString)(int);
static struct {
  void *h = nullptr;
  fn_ncclGetUniqueId GetUniqueId; fn_ncclCommInitRank CommInitRank; fn_ncclAllReduce AllReduce;
  fn_ncclSend Send; fn_ncclRecv Recv; fn_ncclGroup GroupStart, GroupEnd;
  fn_ncclCommDestroy CommDestroy; fn_ncclGetErrorString GetErrorString;
} g_nccl;
static const int NCCL_F64 = 8, NCCL_F32 = 7, NCCL_SUM = 0, NCCL_MAX = 2;  // ncclFloat64, ncclFloat32, ncclSum, ncclMax
template <class VT> static constexpr int nccl_type() { return std::is_same<VT, float>::value ? NCCL_F32 : NCCL_F64; }
static int load_nccl() {
  if (g_nccl.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) { g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
  if (!g_nccl.h) return set_err(40, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.Send = (fn_ncclSend)dlsym(g_nccl.h, "ncclSend");
  g_nccl.Recv = (fn_ncclRecv)dlsym(g_nccl.h, "ncclRecv");
  g_nccl.GroupStart = (fn_ncclGroup)dlsym(g_nccl.h, "ncclGroupStart");
  g_nccl.GroupEnd = (fn_ncclGroup)dlsym(g_nccl.h, "ncclGroupEnd");
  g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Send || !g_nccl.Recv ||
      !g_nccl.GroupStart || !g_nccl.GroupEnd)
    return set_err(40, "libnccl is missing symbols");
  return 0;
}
#define NC(call)                                                                                              \
  do {                                                                                                        \
    int e_ = (call);                                                                                          \
    if (e_) return set_err(44, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "?"); \
  } while (0)

static cudaEvent_t get_event(pprts_b200_solver *s) {
  if (!s->evpool.empty()) { cudaEvent_t e = s->evpool.back(); s->evpool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
static void collect_events(pprts_b200_solver *s) {
  if (s->evs.empty()) return;
  cudaStreamSynchronize(s->stream);
  for (auto &ev : s->evs) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ev.a, ev.b);
    s->fam_ms[ev.fam] += ms;
    s->evpool.push_back(ev.a);
    s->evpool.push_back(ev.b);
  }
  s->evs.clear();
}

#define LAUNCH(S, FAM, KERNEL, GRID, BLOCK, ...)                           \
  do {                                                                     \
    cudaEvent_t ea_ = nullptr, eb_ = nullptr;                              \
    if ((S)->timing) { ea_ = get_event(S); eb_ = get_event(S); cudaEventRecord(ea_, (S)->stream); } \
    KERNEL<<<(GRID), (BLOCK), 0, (S)->stream>>>(__VA_ARGS__);              \
    if ((S)->timing) { cudaEventRecord(eb_, (S)->stream); (S)->evs.push_back({FAM, ea_, eb_}); } \
    (S)->launches++;                                                       \
    (S)->fam_launch[FAM]++;                                                \
  } while (0)

#define DISPATCH_L(S, ...)                                          \
  do {                                                              \
    switch ((S)->cfg.solver_id) {                                   \
      case 310: { using L = L310; __VA_ARGS__; } break;             \
      case 810: { using L = L810; __VA_ARGS__; } break;             \
      case 324: { using L = L324; __VA_ARGS__; } break;             \
    }                                                               \
  } while (0)
#define DISPATCH_LC(S, ...)                                                             \
  do {                                                                                  \
    if ((S)->cfg.coeff_fp32) { using CT = float; DISPATCH_L(S, __VA_ARGS__); }          \
    else { using CT = double; DISPATCH_L(S, __VA_ARGS__); }                             \
  } while (0)

static inline unsigned int nblk(long n, int bs) { return (unsigned int)((n + bs - 1) / bs); }
static const int VEC_BLOCK = 256;
static inline unsigned int vec_grid(long n) { return (unsigned int)std::min<long>((n + VEC_BLOCK - 1) / VEC_BLOCK, 148L * 16); }

// ---------------------------------------------------------------------------------------------------------------
// x/y decomposition: halo exchange over NCCL send/recv (replaces the PETSc VecScatter / MPI_Isend+Irecv of
// src/pprts_explicit.F90:232-328, :715-848 and the reverse-add of src/pprts_base.F90:1676-1731) and the scalar
// all-reduce behind every dot product.  All traffic is enqueued on the solver's stream: no host synchronisation.
// ---------------------------------------------------------------------------------------------------------------
struct HaloDofs { int d0x, ndx, d0y, ndy; };  // x-side dofs [d0x, d0x+ndx), y-side dofs [d0y, d0y+ndy) of the vector

static int need_comm(const pprts_b200_solver *s) {
  if (s->decomposed && !s->nccl_comm) return set_err(43, "decomposed solver: call pprts_b200_comm_init first");
  return 0;
}
// sum the n doubles at sc over all ranks (dot products, norms)
static int allreduce_scalars(pprts_b200_solver *s, double *sc, int n) {
  if (!s->decomposed) return 0;
  NC(g_nccl.AllReduce(sc, sc, (size_t)n, NCCL_F64, NCCL_SUM, s->nccl_comm, s->stream));
  return 0;
}
// Ghost FILL: my low faces (i = 0 / j = 0) -> the ghost block of the west / south neighbour; my ghost block <- the
// low faces of the east / north neighbour.  fill_x / fill_y select the directions (the direct beam needs only one).
template <class VT>
static int halo_fill(pprts_b200_solver *s, VT *v, int nd_total, HaloDofs h, bool fill_x, bool fill_y) {
  constexpr int NT = nccl_type<VT>();
  VT *hs0 = (VT *)s->hs[0], *hs1 = (VT *)s->hs[1];
  const Grid &g = s->g;
  fill_x = fill_x && g.gx; fill_y = fill_y && g.gy;
  if (!fill_x && !fill_y) return 0;
  const size_t nx = (size_t)h.ndx * (g.Nz + 1) * g.Ny, ny = (size_t)h.ndy * (g.Nz + 1) * g.Nx;
  if (fill_x) LAUNCH(s, FAM_HALO, k_halo_pack<VT>, nblk((long)nx, 256), 256, g, 0, h.d0x, h.ndx, (const VT *)v, hs0);
  if (fill_y) LAUNCH(s, FAM_HALO, k_halo_pack<VT>, nblk((long)ny, 256), 256, g, 1, h.d0y, h.ndy, (const VT *)v, hs1);
  NC(g_nccl.GroupStart());
  if (fill_x) {
    NC(g_nccl.Send(hs0, nx, NT, s->nb_west, s->nccl_comm, s->stream));
    NC(g_nccl.Recv(v + ghost_x_base(g, nd_total) + (size_t)h.d0x * (g.Nz + 1) * g.Ny, nx, NT, s->nb_east, s->nccl_comm, s->stream));
  }
  if (fill_y) {
    NC(g_nccl.Send(hs1, ny, NT, s->nb_south, s->nccl_comm, s->stream));
    NC(g_nccl.Recv(v + ghost_y_base(g, nd_total) + (size_t)h.d0y * (g.Nz + 1) * g.Nx, ny, NT, s->nb_north, s->nccl_comm, s->stream));
  }
  NC(g_nccl.GroupEnd());
  s->halo_exchanges++;
  return 0;
}
// Ghost PUSH: rows this rank produced on its high faces (its ghost block) -> the east / north neighbour, which stores
// them on its low faces (dofs in maskx / masky, levels [n1d, Nz)); ndot > 0 adds the rows' dot products to `dots`.
template <class VT>
static int halo_push(pprts_b200_solver *s, VT *v, int nd_total, HaloDofs h, bool push_x, bool push_y,
                     unsigned int maskx, unsigned int masky, int ndot, const VT *w0, double *dots) {
  constexpr int NT = nccl_type<VT>();
  VT *hr[2] = {(VT *)s->hr[0], (VT *)s->hr[1]};
  const Grid &g = s->g;
  push_x = push_x && g.gx; push_y = push_y && g.gy;
  if (!push_x && !push_y) return 0;
  const size_t nx = (size_t)h.ndx * (g.Nz + 1) * g.Ny, ny = (size_t)h.ndy * (g.Nz + 1) * g.Nx;
  NC(g_nccl.GroupStart());
  if (push_x) {
    NC(g_nccl.Send(v + ghost_x_base(g, nd_total) + (size_t)h.d0x * (g.Nz + 1) * g.Ny, nx, NT, s->nb_east, s->nccl_comm, s->stream));
    NC(g_nccl.Recv(hr[0], nx, NT, s->nb_west, s->nccl_comm, s->stream));
  }
  if (push_y) {
    NC(g_nccl.Send(v + ghost_y_base(g, nd_total) + (size_t)h.d0y * (g.Nz + 1) * g.Nx, ny, NT, s->nb_north, s->nccl_comm, s->stream));
    NC(g_nccl.Recv(hr[1], ny, NT, s->nb_south, s->nccl_comm, s->stream));
  }
  NC(g_nccl.GroupEnd());
  s->halo_exchanges++;
  for (int dir = 0; dir < 2; ++dir) {
    if (!(dir == 0 ? push_x : push_y)) continue;
    const size_t n = dir == 0 ? nx : ny;
    const int d0 = dir == 0 ? h.d0x : h.d0y, nd = dir == 0 ? h.ndx : h.ndy;
    const unsigned int mask = dir == 0 ? maskx : masky;
    double *tmp = s->sc + SC_TMP;
    if (ndot == 0)
      LAUNCH(s, FAM_HALO, (k_halo_unpack<0, VT>), nblk((long)n, 256), 256, g, dir, d0, nd, mask, s->n1d, g.kb1d, (const VT *)hr[dir], v, w0, s->partials, s->counter, tmp, dots);
    else if (ndot == 1)
      LAUNCH(s, FAM_HALO, (k_halo_unpack<1, VT>), nblk((long)n, 256), 256, g, dir, d0, nd, mask, s->n1d, g.kb1d, (const VT *)hr[dir], v, w0, s->partials, s->counter, tmp, dots);
    else
      LAUNCH(s, FAM_HALO, (k_halo_unpack<2, VT>), nblk((long)n, 256), 256, g, dir, d0, nd, mask, s->n1d, g.kb1d, (const VT *)hr[dir], v, w0, s->partials, s->counter, tmp, dots);
  }
  return 0;
}
static HaloDofs diff_halo(const pprts_b200_solver *s) {
  int nft = 2, nfs = 4;
  DISPATCH_L(s, { nft = L::Nft; nfs = L::Nfs; });
  return HaloDofs{nft, nfs, nft + nfs, nfs};
}
static HaloDofs dir_halo(const pprts_b200_solver *s) {
  int ndt = 1, nds = 1;
  DISPATCH_L(s, { ndt = L::Ndt; nds = L::Nds; });
  return HaloDofs{ndt, nds, ndt + nds, nds};
}
// bit d set: side dof d (within its direction) is an inward stream, i.e. produced on the neighbour's high face
static unsigned int diff_inward_mask(const pprts_b200_solver *s) {
  unsigned int m = 0;
  DISPATCH_L(s, { for (int d = 0; d < L::Nfs; ++d) if (L::side_in(d)) m |= 1u << d; });
  return m;
}

extern "C" void pprts_b200_default_config(pprts_b200_config *cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->solver_id = 310;
  cfg->twostr_ratio = 2.0;
  cfg->coeff_fp32 = 1;
  cfg->by_coeff_divergence = 1;
  cfg->rtol = -1; cfg->atol = -1;
  cfg->precond = 3;
  cfg->device = -1;
  cfg->keep_solutions = 1;
  cfg->nxproc = cfg->nyproc = 1;
  cfg->diff_explicit = 0;
  cfg->krylov_fp32 = 1;  // fp32 Krylov work vectors; the convergence test still sees an fp64 true residual
  cfg->collapseindex = 1;
  cfg->max_solution_err = 5e3 / (3600.0 * 24.0);  // src/tenstream_options.F90:90
  cfg->max_solution_time = 0.0;                    // :94 -> error estimates disabled unless the caller sets it
}

template <class T>
static int dalloc(T **p, size_t n) {
  CU(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
  CU(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}
#define DA(p, n) do { int rc_ = dalloc(&(p), (n)); if (rc_) return rc_; } while (0)

// src/pprts.F90:1118-1183 (setup_suninfo) incl. cartesian_2_spherical (helper_functions.fypp:1993-2016)
static void setup_suninfo(pprts_b200_solver *s, const double sundir[3]) {
  const double d2r = PB_PI / 180.0;
  // norm2() of the reference does not underflow for tiny components (test_helper_functions.F90:774-792): hypot
  double n = std::hypot(std::hypot(sundir[0], sundir[1]), sundir[2]);
  double sn2 = n > 0 ? sundir[2] / n : sundir[2];
  double theta = std::acos(std::max(-1.0, std::min(1.0, -sn2))) / d2r;
  double px = sundir[0], py = sundir[1], pn = std::hypot(px, py);
  if (pn > 0) { px /= pn; py /= pn; }
  double az = std::acos(std::max(-1.0, std::min(1.0, -py)));
  az = az * ((-px) >= 0 ? 1.0 : -1.0);
  s->phi = az / d2r;
  s->theta = theta;
  s->mu = std::max(std::cos(s->theta * d2r), 0.0);
  if (s->theta >= 90.0) s->theta = -1.0;
  s->costheta = std::max(std::cos(s->theta * d2r), 0.0);
  double sp = std::acos(std::cos(s->phi * d2r));
  s->symmetry_phi = std::min(90.0, std::max(0.0, std::asin(std::sin(sp)) / d2r));
  s->xinc = (std::sin(s->phi * d2r) > 0.0) ? 0 : 1;
  s->yinc = (std::cos(s->phi * d2r) < 0.0) ? 1 : 0;
  s->angles_dirty = true;
  s->coeffs_dirty = true;
}

// Contract the 6-D direct tables over (phi, theta) for the current sun: interp_vec_bilinear_iterative restricted to
// the two slowest axes (snapping included), fp64.  Touches <= 4 contiguous (phi,theta) blocks of the mapped files.
static int contract_direct_tables(pprts_b200_solver *s) {
  const pprts_b200_lut *lut = s->lut;
  if (!lut->Tdir || !lut->Sdir) return set_err(4, "direct LUT tables not loaded");
  const double snap = (double)1e-3f;
  long lat4 = 1;
  for (int d = 0; d < 4; ++d) lat4 *= lut->n[d];
  double pti[2] = {h_find_real_location(lut->ax[4], (float)s->symmetry_phi),
                   h_find_real_location(lut->ax[5], (float)s->theta)};
  long off_lo[2], off_hi[2], base = 0;
  double wlo[2], whi[2];
  int Ninterp = 0;
  for (int d = 0; d < 2; ++d) {
    double frac = pti[d] - (double)(long)pti[d];
    if (frac < snap || frac > 1.0 - snap) base += lut->off[4 + d] * (std::lround(pti[d]) - 1);
    else {
      long b = (long)pti[d];
      whi[Ninterp] = pti[d] - (double)b; wlo[Ninterp] = 1.0 - whi[Ninterp];
      off_lo[Ninterp] = lut->off[4 + d] * (b - 1); off_hi[Ninterp] = lut->off[4 + d] * b;
      ++Ninterp;
    }
  }
  const int ncorner = 1 << Ninterp;
  long cofs[4];
  double cw[4];
  for (int b = 0; b < ncorner; ++b) {
    long o = base; double w = 1.0;
    for (int d = 0; d < Ninterp; ++d) { if (b & (1 << d)) { o += off_hi[d]; w *= whi[d]; } else { o += off_lo[d]; w *= wlo[d]; } }
    cofs[b] = o; cw[b] = w;
  }
  for (int pass = 0; pass < 2; ++pass) {
    const int nc = pass == 0 ? s->Nd * s->Nd : s->Nd * s->Nf;
    const float *tab = pass == 0 ? lut->Tdir : lut->Sdir;
    std::vector<double> h((size_t)lat4 * nc, 0.0);
    for (int b = 0; b < ncorner; ++b) {
      const float *src = tab + cofs[b] * nc;
      const double w = cw[b];
      for (size_t q = 0; q < h.size(); ++q) h[q] = std::fma(w, (double)src[q], h[q]);
    }
    CU(cudaMemcpyAsync(pass == 0 ? s->d_Tdir4 : s->d_Sdir4, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  build_sym_maps(s->cfg.solver_id, s->xinc == 0, s->yinc == 0, s->sym);
  s->angles_dirty = false;
  return 0;
}


// determine_1d_layers, second half (src/pprts.F90:708-724): the number of 1-D layers is made the same on all ranks
// (N1dlayers_max) and the first N1dlayers_max layers become 1-D -- note that a 1-D BOTTOM layer found on its own stays
// 1-D as well: l1d = [T..T, F..F, T].  count3d_frac < 0: single rank (fraction from the local mask).
static int finalize_1d_layers(pprts_b200_solver *s, int n1d_max, double unconstrained_fraction) {
  Grid &g = s->g;
  const int Nza = s->Nz_atm, ci = s->icollapse;
  for (int k = 0; k < n1d_max && k < Nza; ++k) s->l1d_atm[k] = 1;
  int count3d = 0;
  for (int k = 0; k < Nza; ++k) count3d += !s->l1d_atm[k];
  s->unconstrained_fraction = unconstrained_fraction >= 0.0 ? unconstrained_fraction : (double)count3d / (double)Nza;
  s->l1d.assign(g.Nz, 0);
  for (int k = 0; k < g.Nz; ++k) s->l1d[k] = s->l1d_atm[k + ci - 1];  // atmk, src/pprts_base.F90:1088-1093
  const int n1d_old = s->n1d, kb1d_old = g.kb1d;
  s->n1d = 0;
  while (s->n1d < g.Nz && s->l1d[s->n1d]) ++s->n1d;
  g.kb1d = (s->n1d < g.Nz && s->l1d[g.Nz - 1]) ? g.Nz - 1 : g.Nz;
  if (s->a11 && s->n1d == n1d_old && g.kb1d == kb1d_old) return 0;
  for (int k = s->n1d; k < g.kb1d; ++k)
    if (s->l1d[k]) return set_err(1, "1-D layer mask is neither contiguous from the top nor the bottom layer alone");
  // Eddington coefficients: fields over the top n1d layers, or over all layers when the bottom layer is 1-D too
  const size_t n1c = (size_t)(g.kb1d < g.Nz ? g.Nz : s->n1d) * g.ncol;
  double **a[5] = {&s->a11, &s->a12, &s->a13, &s->a23, &s->a33};
  for (auto pp : a) { if (*pp) { cudaFree(*pp); *pp = nullptr; } DA(*pp, n1c); }
  s->coeffs_dirty = true; s->prec_ready = false; s->have_optprop = false;
  return 0;
}

static int create_impl(const pprts_b200_config *cfg, const pprts_b200_lut *lut, pprts_b200_solver *s, int Nd, int Nf);
extern "C" int pprts_b200_create(const pprts_b200_config *cfg, const pprts_b200_lut *lut, pprts_b200_solver **out) {
  if (!cfg || !lut || !out) return set_err(1, "null argument");
  int Nd, Nf;
  if (stream_counts(cfg->solver_id, Nd, Nf)) return set_err(1, "unknown solver id %d", cfg->solver_id);
  if (lut->solver_id != cfg->solver_id) return set_err(1, "LUT was opened for solver %d, not %d", lut->solver_id, cfg->solver_id);
  if (cfg->Nz < 1 || cfg->Nx < 1 || cfg->Ny < 1) return set_err(1, "bad grid %d %d %d", cfg->Nz, cfg->Nx, cfg->Ny);
  if (std::fabs(cfg->dx - cfg->dy) > 1e-9 * std::fabs(cfg->dx))  // src/pprts.F90:459
    return set_err(1, "dx and dy currently have to be the same %g vs %g", cfg->dx, cfg->dy);
  if ((cfg->dz1d == nullptr) == (cfg->dz3d == nullptr)) return set_err(1, "have to give either dz1d or dz3d");
  if (cfg->nxproc < 1 || cfg->nyproc < 1 || cfg->xrank < 0 || cfg->xrank >= cfg->nxproc || cfg->yrank < 0 ||
      cfg->yrank >= cfg->nyproc)
    return set_err(1, "bad process grid %d x %d, position (%d, %d)", cfg->nxproc, cfg->nyproc, cfg->xrank, cfg->yrank);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err(90, "no CUDA device: libpprts_b200 has no CPU fallback");
  if (cfg->geometric_coeffs && Nd != 3) return set_err(23, "geometric coeffs currently only implemented for 3 direct streams");
  if ((cfg->collapseindex > 1 ? cfg->collapseindex : 1) > cfg->Nz)
    return set_err(1, "collapseindex %d exceeds the %d atmosphere layers", cfg->collapseindex, cfg->Nz);
  auto *s = new pprts_b200_solver();
  const int rc = create_impl(cfg, lut, s, Nd, Nf);
  if (rc) {  // every allocation made so far hangs off the solver object: one cleanup path for all failures
    char msg[sizeof(g_err)];
    snprintf(msg, sizeof(msg), "%s", g_err);
    pprts_b200_destroy(s);
    return set_err(rc, "%s", msg);
  }
  *out = s;
  return 0;
}
static int create_impl(const pprts_b200_config *cfg, const pprts_b200_lut *lut, pprts_b200_solver *s, int Nd, int Nf) {
  s->cfg = *cfg;
  if (s->cfg.twostr_ratio <= 0) s->cfg.twostr_ratio = 2.0;
  s->cfg.dz1d = s->cfg.dz3d = nullptr;
  s->have_dz3d = cfg->dz3d != nullptr;
  s->lut = lut;
  s->Nd = Nd; s->Nf = Nf;
  if (cfg->device >= 0) { CU(cudaSetDevice(cfg->device)); s->device = cfg->device; }
  else CU(cudaGetDevice(&s->device));
  Grid &g = s->g;
  const int ci = cfg->collapseindex > 1 ? cfg->collapseindex : 1;
  s->icollapse = ci; s->Nz_atm = cfg->Nz;
  g.Nz = cfg->Nz - ci + 1; g.Nx = cfg->Nx; g.Ny = cfg->Ny;
  g.ncol = (long)g.Nx * g.Ny; g.ncell = g.ncol * g.Nz; g.nlev = g.ncol * (g.Nz + 1);
  g.ncp = (g.ncell + PB_TILE - 1) / PB_TILE * PB_TILE;  // whole coefficient tiles (tile-blocked layout, common.cuh)
  g.dx = cfg->dx; g.dy = cfg->dy;
  g.gx = cfg->nxproc > 1; g.gy = cfg->nyproc > 1;
  s->decomposed = g.gx || g.gy;
  s->glob_ncol = (double)g.ncol;  // replaced by the sum over ranks in pprts_b200_comm_init
  CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CU(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, s->device));
  s->use_tma = getenv("PPRTS_B200_NO_TMA") == nullptr;
  if (const char *v = getenv("PPRTS_B200_MM_VARIANT")) s->mm_variant = atoi(v);

  // dz in device layout + determine_1d_layers (src/pprts.F90:649-725), both on the atmosphere grid
  const int Nza = s->Nz_atm;
  s->ncell_atm = g.ncol * Nza; s->nlev_atm = g.ncol * (Nza + 1);
  std::vector<double> hdz((size_t)s->ncell_atm);
  std::vector<unsigned char> thick(Nza, 0), l1d_atm(Nza, 0);
  for (int j = 0; j < g.Ny; ++j) for (int i = 0; i < g.Nx; ++i) for (int k = 0; k < Nza; ++k) {
    double v = cfg->dz1d ? cfg->dz1d[k] : cfg->dz3d[k + (size_t)Nza * (i + (size_t)g.Nx * j)];
    hdz[((size_t)k * g.Ny + j) * g.Nx + i] = v;
    if (v / cfg->dx > s->cfg.twostr_ratio) thick[k] = 1;
  }
  l1d_atm[Nza - 1] = thick[Nza - 1];
  for (int k = Nza - 2; k >= 0; --k) if (thick[k]) { for (int kk = 0; kk <= k; ++kk) l1d_atm[kk] = 1; break; }
  for (int k = 0; k < ci && ci > 1; ++k) l1d_atm[k] = 1;  // :692-706 layers that are collapsed have to be 1-D
  s->l1d_atm = l1d_atm;
  int n1d_atm = 0;
  for (int k = 0; k < Nza; ++k) n1d_atm += l1d_atm[k];
  { int rc_ = finalize_1d_layers(s, n1d_atm, -1.0); if (rc_) return rc_; }

  const size_t aoff = (size_t)(ci - 1) * g.ncol;
  DA(s->atm_dz, s->ncell_atm); DA(s->atm_kabs, s->ncell_atm); DA(s->atm_ksca, s->ncell_atm); DA(s->atm_gg, s->ncell_atm);
  DA(s->atm_planck, s->nlev_atm); DA(s->Bsrfc, g.ncol); DA(s->albedo, g.ncol); DA(s->raw, s->nlev_atm);
  s->dz = s->atm_dz + aoff; s->kabs = s->atm_kabs + aoff; s->ksca = s->atm_ksca + aoff; s->gg = s->atm_gg + aoff;
  s->planck = s->atm_planck + aoff;
  if (ci > 1) { DA(s->Btop, g.ncol); DA(s->Bbot, g.ncol); }
  const size_t csz = cfg->coeff_fp32 ? 4 : 8;
  CU(cudaMalloc(&s->d2d, (size_t)Nd * Nd * g.ncp * csz)); CU(cudaMemset(s->d2d, 0, (size_t)Nd * Nd * g.ncp * csz));
  CU(cudaMalloc(&s->d2f, (size_t)Nd * Nf * g.ncp * csz)); CU(cudaMemset(s->d2f, 0, (size_t)Nd * Nf * g.ncp * csz));
  CU(cudaMalloc(&s->f2f, (size_t)Nf * Nf * g.ncp * csz)); CU(cudaMemset(s->f2f, 0, (size_t)Nf * Nf * g.ncp * csz));
  DA(s->cdiv_dir, (size_t)Nd * g.ncp); DA(s->cdiv_diff, (size_t)Nf * g.ncp);
  DA(s->tile_class, (size_t)(g.ncp / PB_TILE));
  DA(s->b, (size_t)state_len(g, Nf));
  for (int q = 0; q < 8; ++q) DA(s->kr[q], (size_t)state_len(g, Nf) + 2);  // +2: two 16-byte aligned fp32 vectors fit in one
  DA(s->edir_tmp, (size_t)state_len(g, Nd));
  if (cfg->open_bc) { DA(s->bcx, (size_t)Nd * g.Nz * g.Ny); DA(s->bcy, (size_t)Nd * g.Nz * g.Nx); }
  if (s->decomposed) {
    const size_t hx = (size_t)std::max(Nd, Nf) * (g.Nz + 1) * g.Ny, hy = (size_t)std::max(Nd, Nf) * (g.Nz + 1) * g.Nx;
    DA(s->hs[0], hx); DA(s->hr[0], hx); DA(s->hs[1], hy); DA(s->hr[1], hy);
  }
  {
    int np = 1;  // up/down pairs of top streams: the column preconditioner keeps np x np matrices per level
    DISPATCH_L(s, np = L::Nft / 2);
    DA(s->precR, (size_t)np * np * g.nlev); DA(s->precInv, (size_t)np * np * g.ncell);
    if (np > 1 || cfg->krylov_fp32) { DA(s->precRf, (size_t)np * np * g.nlev); DA(s->precInvf, (size_t)np * np * g.ncell); }
    if (np == 1 && cfg->krylov_fp32 && !getenv("PPRTS_B200_PREC_FP32")) {  // PPRTS_B200_PREC_FP32: sweeps read T itself, for comparison
      int pu = 0, pd = 0;
      DISPATCH_L(s, (pu = PrecPack<L>::PU, pd = PrecPack<L>::PD));
      CU(cudaMalloc(&s->precPU, (size_t)pu * g.ncp * sizeof(__half2)));
      CU(cudaMalloc(&s->precPD, (size_t)pd * g.ncp * sizeof(__half2)));
    }
  }
  if (cfg->precond == 3) {  // coarse-grid correction: layouts with one up/down pair of top streams
    int nft = 0, nfs = 0;
    DISPATCH_L(s, (nft = L::Nft, nfs = L::Nfs));
    if (nft == 2) {
      // aggregate sizes, finest first (default: one level of 4 x 4 columns); PPRTS_B200_CG_AGG = "4,16" adds levels
      std::vector<int> ags;
      if (const char *v = getenv("PPRTS_B200_CG_AGG")) {
        for (const char *q = v; *q;) { ags.push_back(std::max(2, atoi(q))); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
      }
      if (ags.empty()) { ags.push_back(4); ags.push_back(8); }
      const int ns = 2 * nfs, ncoef = 4 + 4 * ns + ns * ns;
      for (int ag : ags) {
        if (s->ncgl >= 3) break;
        if (s->ncgl > 0 && (ag % s->cgl[s->ncgl - 1].c.AG != 0 || ag <= s->cgl[s->ncgl - 1].c.AG))
          return set_err(2, "PPRTS_B200_CG_AGG: every aggregate size has to be a multiple of the previous one");
        auto &lv = s->cgl[s->ncgl];
        CoarseGrid &c = lv.c;
        c.AG = ag; c.NxC = (g.Nx + ag - 1) / ag; c.NyC = (g.Ny + ag - 1) / ag; c.Nz = g.Nz;
        c.ncolC = (long)c.NxC * c.NyC; c.ncc = c.ncolC * g.Nz; c.nclev = c.ncolC * (g.Nz + 1);
        DA(lv.coef, (size_t)ncoef * c.ncc); DA(lv.R, (size_t)c.ncolC); DA(lv.sA, (size_t)4 * c.ncc); DA(lv.sB, (size_t)2 * c.ncc);
        DA(lv.rbar, (size_t)Nf * c.nclev); DA(lv.e, (size_t)2 * c.ncc); DA(lv.x, (size_t)Nf * c.nclev);
        s->ncgl++;
      }
    }
  }
  for (int q = 0; q < 4; ++q) { DA(s->res[q], g.nlev); DA(s->acc[q], g.nlev); }
  DA(s->out_stage, (size_t)std::max(Nd, Nf) * g.nlev);
  DA(s->partials, (size_t)4 * (nblk(g.ncell, 128) + 148 * 16 + 64));
  DA(s->sc, SC_COUNT); DA(s->counter, 4); DA(s->bad, 1);
  CU(cudaHostAlloc((void **)&s->h_sc, SC_COUNT * sizeof(double), cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void **)&s->h_sc_dev, s->h_sc, 0));
  CU(cudaHostAlloc((void **)&s->h_bad, sizeof(int), cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void **)&s->h_bad_dev, s->h_bad, 0));
  CU(cudaMemcpy(s->atm_dz, hdz.data(), hdz.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (cfg->geometric_coeffs || cfg->dz3d) {  // determine_vertex_heights / atm%hhl (src/pprts.F90:585-600, :726-755)
    DA(s->vh_atm, s->nlev_atm);
    k_vert_heights<<<nblk(g.ncol, 128), 128, 0, s->stream>>>(g.ncol, s->Nz_atm, s->atm_dz, s->vh_atm);
    s->vh = s->vh_atm + (size_t)(s->icollapse - 1) * g.ncol;
  }

  // LUT to the device: the four shared axes, the diffuse table; direct tables are contracted per sun angle
  long lat4 = 1;
  for (int d = 0; d < 4; ++d) {
    lat4 *= lut->n[d];
    DA(s->d_ax[d], lut->n[d]);
    CU(cudaMemcpy(s->d_ax[d], lut->ax[d].data(), lut->n[d] * sizeof(float), cudaMemcpyHostToDevice));
    s->lutdev.n[d] = lut->n[d]; s->lutdev.off[d] = lut->off[d]; s->lutdev.ax[d] = s->d_ax[d];
  }
  DA(s->d_Sdiff, (size_t)lat4 * Nf * Nf);
  CU(cudaMemcpy(s->d_Sdiff, lut->Sdiff, (size_t)lat4 * Nf * Nf * sizeof(float), cudaMemcpyHostToDevice));
  DA(s->d_Tdir4, (size_t)lat4 * Nd * Nd); DA(s->d_Sdir4, (size_t)lat4 * Nd * Nf);
  s->lutdev.Tdir4 = s->d_Tdir4; s->lutdev.Sdir4 = s->d_Sdir4; s->lutdev.Sdiff = s->d_Sdiff;
  build_sym_maps(cfg->solver_id, false, false, s->sym);
  setup_suninfo(s, cfg->sundir);
  return 0;
}

static void free_solution(Solution &sol) {
  cudaFree(sol.edir); cudaFree(sol.ediff); cudaFree(sol.abso);
  sol = Solution();
}

extern "C" int pprts_b200_destroy(pprts_b200_solver *s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (auto &kv : s->solutions) free_solution(kv.second);
  for (auto &kv : s->staged) { cudaFree(kv.second.kabs); cudaFree(kv.second.ksca); cudaFree(kv.second.g);
                               cudaFree(kv.second.planck); cudaFree(kv.second.Bsrfc); cudaFree(kv.second.albedo2d); }
  double *ptrs[] = {s->atm_dz, s->atm_kabs, s->atm_ksca, s->atm_gg, s->atm_planck, s->Btop, s->Bbot, s->Bsrfc, s->albedo, s->raw, s->a11, s->a12, s->a13,
                    s->a23, s->a33, s->cdiv_dir, s->cdiv_diff, s->b, s->edir_tmp, s->precR, s->precInv, s->out_stage,
                    s->partials, s->sc, s->d_Tdir4, s->d_Sdir4};
  for (double *p : ptrs) cudaFree(p);
  for (int q = 0; q < 8; ++q) cudaFree(s->kr[q]);
  cudaFree(s->bcx); cudaFree(s->bcy);
  for (double *p : s->blur_buf) cudaFree(p);
  for (auto &ig : s->iter_graph) if (ig.exec) cudaGraphExecDestroy(ig.exec);
  cudaFree(s->precRf); cudaFree(s->precInvf); cudaFree(s->precPU); cudaFree(s->precPD); cudaFree(s->vh_atm);
  for (double *p : s->vh_ghost) cudaFree(p);
  for (auto &sp : s->spare) { cudaFree(sp[0]); cudaFree(sp[1]); cudaFree(sp[2]); }
  for (auto &lv : s->cgl) { cudaFree(lv.coef); cudaFree(lv.R); cudaFree(lv.sA); cudaFree(lv.sB); cudaFree(lv.rbar); cudaFree(lv.e); cudaFree(lv.x); }
  for (int q = 0; q < 4; ++q) cudaFree(s->egress[q]);
  if (s->egress_ready) cudaEventDestroy(s->egress_ready);
  if (s->egress_done) cudaEventDestroy(s->egress_done);
  for (auto &kv : s->staged) {
    if (kv.second.ready) cudaEventDestroy(kv.second.ready);
    if (kv.second.consumed) cudaEventDestroy(kv.second.consumed);
  }
  cudaFree(s->bld_iface); cudaFree(s->bld_albedo); cudaFree(s->bld_planck); cudaFree(s->bld_out);
  for (int q = 0; q < 4; ++q) { cudaFree(s->res[q]); cudaFree(s->acc[q]); cudaFree(s->d_ax[q]); }
  cudaFree(s->tile_class); cudaFree(s->d2d); cudaFree(s->d2f); cudaFree(s->f2f); cudaFree(s->d_Sdiff); cudaFree(s->counter); cudaFree(s->bad);
  for (int q = 0; q < 2; ++q) { cudaFree(s->hs[q]); cudaFree(s->hr[q]); }
  cudaFreeHost(s->h_sc); cudaFreeHost(s->h_bad);
  for (auto &ev : s->evs) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (auto e : s->evpool) cudaEventDestroy(e);
  if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
  for (int q = 0; q < 2; ++q) if (s->band_uploaded[q]) cudaEventDestroy(s->band_uploaded[q]);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

extern "C" int pprts_b200_set_angles(pprts_b200_solver *s, const double sundir[3]) {
  if (!s) return set_err(1, "null solver");
  setup_suninfo(s, sundir);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// set_optical_properties (src/pprts.F90:1764-2076)
// ---------------------------------------------------------------------------------------------------------------
static int ingest_field(pprts_b200_solver *s, const double *src_dev_or_host, bool on_device, int nz, double *dst) {
  const Grid &g = s->g;
  const double *in = src_dev_or_host;
  if (!on_device) {
    CU(cudaMemcpyAsync(s->raw, src_dev_or_host, (size_t)nz * g.ncol * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    in = s->raw;
  }
  dim3 grid((g.Nx + 31) / 32, (nz + 31) / 32, g.Ny), block(32, 8);
  LAUNCH(s, FAM_IO, k_zxy_to_kji<double>, grid, block, nz, g.Nx, g.Ny, in, dst);
  return 0;
}
// 2-D (Nx,Ny) column-major is already x fastest == device layout
static int ingest_2d(pprts_b200_solver *s, const double *src, bool on_device, double *dst) {
  CU(cudaMemcpyAsync(dst, src, (size_t)s->g.ncol * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s->stream));
  return 0;
}

static int set_optprop_impl(pprts_b200_solver *s, double albedo, const double *kabs, const double *ksca,
                            const double *g, const double *planck, const double *planck_srfc, const double *albedo_2d,
                            int ldelta, bool on_device) {
  const Grid &gr = s->g;
  CU(cudaSetDevice(s->device));
  int rc;
  if (albedo_2d) { if ((rc = ingest_2d(s, albedo_2d, on_device, s->albedo))) return rc; }
  else {
    std::vector<double> h((size_t)gr.ncol, albedo);
    CU(cudaMemcpyAsync(s->albedo, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  // the staging buffer `raw` is reused: serialize the three uploads on the stream (copies are stream ordered)
  // the caller's fields live on the atmosphere grid (Nz_atm layers); the solver grid is their lower part
  if (kabs) { if ((rc = ingest_field(s, kabs, on_device, s->Nz_atm, s->atm_kabs))) return rc; }
  if (ksca) { if ((rc = ingest_field(s, ksca, on_device, s->Nz_atm, s->atm_ksca))) return rc; }
  if (g) { if ((rc = ingest_field(s, g, on_device, s->Nz_atm, s->atm_gg))) return rc; }
  s->have_planck = planck != nullptr;
  if (planck) { if ((rc = ingest_field(s, planck, on_device, s->Nz_atm + 1, s->atm_planck))) return rc; }
  s->have_Bsrfc = planck_srfc != nullptr;
  if (planck_srfc) { if ((rc = ingest_2d(s, planck_srfc, on_device, s->Bsrfc))) return rc; }
  CU(cudaMemsetAsync(s->bad, 0, sizeof(int), s->stream));
  LAUNCH(s, FAM_IO, k_delta_scale, nblk(s->ncell_atm, 256), 256, s->ncell_atm, s->atm_kabs, s->atm_ksca, s->atm_gg, ldelta, s->bad);
  if (s->n1d > 0)
    LAUNCH(s, FAM_IO, k_eddington, nblk((long)s->n1d * gr.ncol, 256), 256, 0L, (long)s->n1d * gr.ncol, s->costheta, s->kabs,
           s->ksca, s->gg, s->dz, s->a11, s->a12, s->a13, s->a23, s->a33);
  if (gr.kb1d < gr.Nz)  // the bottom layer on its own (is1d, common.cuh)
    LAUNCH(s, FAM_IO, k_eddington, nblk((long)(gr.Nz - gr.kb1d) * gr.ncol, 256), 256, (long)gr.kb1d * gr.ncol, gr.ncell,
           s->costheta, s->kabs, s->ksca, s->gg, s->dz, s->a11, s->a12, s->a13, s->a23, s->a33);
  if (s->icollapse > 1)  // handle_atm_collapse (src/pprts.F90:2080-2123)
    LAUNCH(s, FAM_IO, k_collapse, nblk(gr.ncol, 128), 128, gr.ncol, s->icollapse, s->costheta, (int)s->have_planck,
           s->atm_kabs, s->atm_ksca, s->atm_gg, s->atm_dz, s->atm_planck, s->a11, s->a12, s->a13, s->a23, s->a33,
           s->Btop, s->Bbot);
  k_export_flag<<<1, 1, 0, s->stream>>>(s->bad, s->h_bad_dev);
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  if (*s->h_bad) return set_err(5, "set_optical_properties :: found illegal value in delta_scaled optical properties!");
  s->have_optprop = true;
  s->coeffs_dirty = true;
  s->prec_ready = false;
  return 0;
}

extern "C" int pprts_b200_set_optical_properties(pprts_b200_solver *s, double albedo, const double *kabs,
                                                 const double *ksca, const double *g, const double *planck,
                                                 const double *planck_srfc, const double *albedo_2d, int ldelta) {
  if (!s) return set_err(1, "null solver");
  return set_optprop_impl(s, albedo, kabs, ksca, g, planck, planck_srfc, albedo_2d, ldelta, false);
}

extern "C" int pprts_b200_stage_optical_properties(pprts_b200_solver *s, int slot, double albedo, const double *kabs,
                                                   const double *ksca, const double *g, const double *planck,
                                                   const double *planck_srfc, const double *albedo_2d) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const Grid &gr = s->g;
  Staged &st = s->staged[slot];
  st.albedo = albedo;
  if (st.consumed) CU(cudaStreamWaitEvent(s->copy_stream, st.consumed, 0));  // the previous content is still being ingested
  auto up = [&](double **dst, const double *src, size_t n) -> int {
    if (!*dst) CU(cudaMalloc((void **)dst, n * sizeof(double)));
    CU(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
  };
  int rc;
  if ((rc = up(&st.kabs, kabs, s->ncell_atm)) || (rc = up(&st.ksca, ksca, s->ncell_atm)) || (rc = up(&st.g, g, s->ncell_atm))) return rc;
  st.have_planck = planck != nullptr;
  if (planck && (rc = up(&st.planck, planck, s->nlev_atm))) return rc;
  st.have_Bsrfc = planck_srfc != nullptr;
  if (planck_srfc && (rc = up(&st.Bsrfc, planck_srfc, gr.ncol))) return rc;
  st.have_albedo2d = albedo_2d != nullptr;
  if (albedo_2d && (rc = up(&st.albedo2d, albedo_2d, gr.ncol))) return rc;
  return 0;
}
extern "C" int pprts_b200_set_optical_properties_staged(pprts_b200_solver *s, int slot, int ldelta) {
  if (!s) return set_err(1, "null solver");
  auto it = s->staged.find(slot);
  if (it == s->staged.end()) return set_err(1, "no optical properties staged in slot %d", slot);
  Staged &st = it->second;
  CU(cudaSetDevice(s->device));
  if (st.ready) CU(cudaStreamWaitEvent(s->stream, st.ready, 0));  // uploaded by stage_optical_properties_async
  const int rc = set_optprop_impl(s, st.albedo, st.kabs, st.ksca, st.g, st.have_planck ? st.planck : nullptr,
                                  st.have_Bsrfc ? st.Bsrfc : nullptr, st.have_albedo2d ? st.albedo2d : nullptr, ldelta, true);
  if (rc) return rc;
  // the slot may be refilled (stage_optical_properties_async) as soon as the ingest kernels have read it
  if (!st.consumed) CU(cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming));
  CU(cudaEventRecord(st.consumed, s->stream));
  return 0;
}
static int ensure_copy_stream(pprts_b200_solver *s) {
  if (s->copy_stream) return 0;
  CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  for (int q = 0; q < 2; ++q) CU(cudaEventCreateWithFlags(&s->band_uploaded[q], cudaEventDisableTiming));
  return 0;
}
// As stage_optical_properties, but the upload is only ENQUEUED (copy stream): it overlaps whatever the solver's stream
// is doing -- the solve of the previous band in a caller's band loop.  The host arrays must stay valid (and should be
// page-locked) until the slot has been consumed by set_optical_properties_staged, which waits for the upload.
extern "C" int pprts_b200_stage_optical_properties_async(pprts_b200_solver *s, int slot, double albedo, const double *kabs,
                                                         const double *ksca, const double *g, const double *planck,
                                                         const double *planck_srfc, const double *albedo_2d) {
  if (!s) return set_err(1, "null solver");
  if (!kabs || !ksca || !g) return set_err(2, "stage_optical_properties_async: kabs, ksca and g are required");
  CU(cudaSetDevice(s->device));
  int rc = ensure_copy_stream(s);
  if (rc) return rc;
  const Grid &gr = s->g;
  Staged &st = s->staged[slot];
  st.albedo = albedo;
  if (st.consumed) CU(cudaStreamWaitEvent(s->copy_stream, st.consumed, 0));  // the previous content is still being ingested
  auto up = [&](double **dst, const double *src, size_t n) -> int {
    if (!*dst) CU(cudaMalloc((void **)dst, n * sizeof(double)));
    CU(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, s->copy_stream));
    return 0;
  };
  if ((rc = up(&st.kabs, kabs, s->ncell_atm)) || (rc = up(&st.ksca, ksca, s->ncell_atm)) || (rc = up(&st.g, g, s->ncell_atm))) return rc;
  st.have_planck = planck != nullptr;
  if (planck && (rc = up(&st.planck, planck, s->nlev_atm))) return rc;
  st.have_Bsrfc = planck_srfc != nullptr;
  if (planck_srfc && (rc = up(&st.Bsrfc, planck_srfc, gr.ncol))) return rc;
  st.have_albedo2d = albedo_2d != nullptr;
  if (albedo_2d && (rc = up(&st.albedo2d, albedo_2d, gr.ncol))) return rc;
  if (!st.ready) CU(cudaEventCreateWithFlags(&st.ready, cudaEventDisableTiming));
  CU(cudaEventRecord(st.ready, s->copy_stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// solve_pprts
// ---------------------------------------------------------------------------------------------------------------
static void determine_ksp_tolerances(const pprts_b200_solver *s, double &rtol, double &atol) {  // pprts_base.F90:1128-1131
  rtol = 1e-5;
  atol = std::max(1e-8, 1e-4 * s->glob_ncol * (double)(s->g.Nz + 1) * s->unconstrained_fraction);
}

template <class L, class CT>
static bool launch_geometric(pprts_b200_solver *s, long n3d, double sx, double sy, double sz) {
  if constexpr (L::Nd == 3) {
    const Grid &g = s->g;
    LAUNCH(s, FAM_LOOKUP, (k_geometric_coeffs<L, CT>), nblk(n3d, 128), 128, g, s->n1d, s->cfg.geometric_coeffs, sx, sy, sz,
           (const double *)s->kabs, (const double *)s->ksca, (const double *)s->dz, (const double *)s->vh,
           (CT *)s->d2d, (CT *)s->d2f, s->cdiv_dir, s->bad);
    return true;
  } else {
    return false;
  }
}

template <class L, class CT>
static bool launch_lookup_tile(pprts_b200_solver *s, bool lsolar) {
  using K = LookupCfg<L>;
  const Grid &g = s->g;
  constexpr size_t smem = (size_t)4 * 32 * K::STRIDE * sizeof(CT);
  static bool attr_set[64] = {false};
  auto kern = k_lookup_tile<L, CT>;
  const int dev = s->device & 63;
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
    attr_set[dev] = true;
  }
  const long c_begin = (long)s->n1d * g.ncol, c_end = (long)g.kb1d * g.ncol;
  const long tile0 = c_begin / PB_TILE, ntile = (c_end - 1) / PB_TILE - tile0 + 1;
  cudaEvent_t ea = nullptr, eb = nullptr;
  if (s->timing) { ea = get_event(s); eb = get_event(s); cudaEventRecord(ea, s->stream); }
  kern<<<(unsigned int)ntile, PB_TILE, smem, s->stream>>>(g, s->lutdev, s->sym, (int)lsolar, tile0, c_begin, c_end, s->kabs, s->ksca,
                                                        s->gg, s->dz, (CT *)s->d2d, (CT *)s->d2f, (CT *)s->f2f, s->cdiv_dir,
                                                        s->cdiv_diff, s->tile_class);
  if (s->timing) { cudaEventRecord(eb, s->stream); s->evs.push_back({FAM_LOOKUP, ea, eb}); }
  s->launches++;
  s->fam_launch[FAM_LOOKUP]++;
  return true;
}
static int ensure_coeffs(pprts_b200_solver *s, bool lsolar) {
  if (!s->coeffs_dirty && (!lsolar || s->coeffs_have_dir)) return 0;
  const Grid &g = s->g;
  if (lsolar && s->angles_dirty) { int rc = contract_direct_tables(s); if (rc) return rc; }
  const long n3d = (long)(g.kb1d - s->n1d) * g.ncol;
  if (n3d > 0) {
    CU(cudaMemsetAsync(s->tile_class, 0, (size_t)(g.ncp / PB_TILE) * sizeof(unsigned int), s->stream));
    static const bool old_lookup = getenv("PPRTS_B200_LOOKUP_PER_THREAD") != nullptr;  // the round-1 kernel, for comparison
    if (old_lookup)
      DISPATCH_LC(s, LAUNCH(s, FAM_LOOKUP, (k_lookup<L, CT>), nblk(n3d, 128), 128, g, s->lutdev, s->sym, (int)lsolar,
                             s->n1d, s->kabs, s->ksca, s->gg, s->dz, (CT *)s->d2d, (CT *)s->d2f, (CT *)s->f2f,
                             s->cdiv_dir, s->cdiv_diff, (unsigned char *)s->tile_class));
    else {
      bool ok = true;
      DISPATCH_LC(s, ok = (launch_lookup_tile<L, CT>(s, lsolar)));
      if (!ok) return set_err(100, "k_lookup_tile: cannot reserve its shared-memory stage");
    }
  }
  if (s->cfg.geometric_coeffs && lsolar && n3d > 0) {  // src/pprts.F90:3142-3149, :3324-3351
    const double d2r = PB_PI / 180.0;  // sun%sundir = spherical_2_cartesian(phi, theta) (:1148)
    const double th = s->theta < 0 ? 90.0 : s->theta;
    const double sx = -std::sin(th * d2r) * std::sin(s->phi * d2r), sy = -std::sin(th * d2r) * std::cos(s->phi * d2r),
                 sz = -std::cos(th * d2r);
    CU(cudaMemsetAsync(s->bad, 0, sizeof(int), s->stream));
    bool launched = false;
    DISPATCH_LC(s, (launched = launch_geometric<L, CT>(s, n3d, sx, sy, sz)));
    if (!launched) return set_err(23, "geometric coeffs currently only implemented for 3 direct streams");
    k_export_flag<<<1, 1, 0, s->stream>>>(s->bad, s->h_bad_dev);
    CU(cudaStreamSynchronize(s->stream));
    if (*s->h_bad) return set_err(24, "Extinction coeff too small for the geometric coefficients (kabs + ksca < tiny in a 3-D cell)");
  }
  if (s->nbld > 0) {  // set_buildings_coeff (src/pprts.F90:3188-3212, :3492-3494, :3579-3677)
    DISPATCH_LC(s, LAUNCH(s, FAM_LOOKUP, (k_buildings_coeff<L, CT>), nblk(s->nbld, 128), 128, g, (int)lsolar, s->nbld,
                           (const int *)s->bld_iface, (const double *)s->bld_albedo, (CT *)s->d2d, (CT *)s->f2f,
                           (unsigned char *)s->tile_class));
    DISPATCH_LC(s, LAUNCH(s, FAM_LOOKUP, (k_buildings_cdiv<L, CT>), nblk(s->nbld, 128), 128, g, (int)lsolar, s->nbld,
                           (const int *)s->bld_iface, (const CT *)s->d2d, (const CT *)s->d2f, (const CT *)s->f2f,
                           s->cdiv_dir, s->cdiv_diff));
  }
  s->coeffs_dirty = false;
  s->coeffs_have_dir = lsolar;
  s->prec_ready = false;
  return 0;
}

static int scale_flx(pprts_b200_solver *s, Solution &sol, bool lWm2) {
  const Grid &g = s->g;
  if (sol.lsolar_rad && sol.lWm2_dir != lWm2) {
    DISPATCH_L(s, LAUNCH(s, FAM_VEC, (k_scale_flx<L::Ndt, L::Nds, L::dirtop_adiv, L::dirside_adiv>), nblk(g.nlev, 256),
                          256, g, s->dz, sol.edir, (int)lWm2));
    sol.lWm2_dir = lWm2;
  }
  if (sol.lWm2_diff != lWm2) {
    DISPATCH_L(s, LAUNCH(s, FAM_VEC, (k_scale_flx<L::Nft, L::Nfs, 1, 1>), nblk(g.nlev, 256), 256, g, s->dz, sol.ediff,
                          (int)lWm2));
    sol.lWm2_diff = lWm2;
  }
  return 0;
}

static int read_scalars(pprts_b200_solver *s) {
  // The scalars travel as stores of a one-warp kernel into mapped host memory, not as a D2H copy: a copy queues on the
  // copy engine behind whatever bulk download is in flight (get_result_async of the previous band: ~7 ms at C2), which
  // stalled every Krylov iteration that ran meanwhile (measured: end-to-end step 9 ms above the device-resident one).
  k_export_scalars<<<1, SC_COUNT, 0, s->stream>>>(s->sc, s->h_sc_dev);
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

// explicit_edir (src/pprts_explicit.F90:60-230)
// exchange_direct_boundary (src/pprts_explicit.F90:232-328): rows produced on a high face travel downstream to their
// owner (sun from the west/south: xinc/yinc = 1), or the owner's low face is mirrored into the upstream ghost block
static int edir_exchange(pprts_b200_solver *s, double *v) {
  if (!s->decomposed) return 0;
  const HaloDofs h = dir_halo(s);
  int rc = halo_fill(s, v, s->Nd, h, s->xinc == 0, s->yinc == 0);
  if (rc) return rc;
  return halo_push<double>(s, v, s->Nd, h, s->xinc != 0, s->yinc != 0, ~0u, ~0u, 0, nullptr, nullptr);
}

static int solve_edir(pprts_b200_solver *s, Solution &sol, double edirTOA) {
  static const bool jacobi = getenv("PPRTS_B200_EDIR_JACOBI") != nullptr;  // the round-1 fully lagged passes, for comparison
  const Grid &g = s->g;
  int maxiter = s->cfg.maxit_dir > 0 ? s->cfg.maxit_dir : 1000;
  double rtol_d, atol_d;
  determine_ksp_tolerances(s, rtol_d, atol_d);
  double rtol = s->cfg.rtol >= 0 ? s->cfg.rtol : rtol_d, atol = s->cfg.atol >= 0 ? s->cfg.atol : atol_d;
  if (sol.first_dir) { maxiter += 2; rtol = std::min(rtol, rtol_d); atol = std::min(atol, atol_d); }
  // A pass (k_edir_tile_pass) is the reference's sun-ordered Gauss-Seidel sweep inside a tile of 32 x EDIR_TY columns
  // and lags the inputs ACROSS tiles (and across the periodic wrap / the rank boundaries, like the reference) by one
  // pass: the beam needs one pass per tile boundary it crosses, after which a pass changes nothing.
  // The reference's test ||x_n - x_{n-1}|| < atol or < rtol * first (:186-196) is written for a sweep whose change IS
  // its error.  The change of a lagged pass only says how much arrived in this pass, not how much is still on its way:
  // stopping at the same numbers left edir 8e-5 (C1) away from the reference's result at the same tolerances
  // (tests/test_gpu_config_goldens.py).  The passes therefore run until the change is 1e-4 of what the reference asks
  // for -- in practice until the recurrence has ended (it is finite: the beam leaves the domain).
  const double lag = 1e-4;
  rtol *= lag; atol *= lag;
  int adiv = 1;
  DISPATCH_L(s, adiv = L::dirtop_adiv);
  const double fac = edirTOA * g.dx * g.dy / (double)adiv;
  double *xold = sol.edir, *xnew = s->edir_tmp;
  double res1 = 0;
  bool converged = false;
  // -pprts_open_bc (src/pprts_base.F90:1189-1316): the ranks on the sunward domain edges take their side inflow from the
  // single-column solution of their edge columns instead of the periodic image / the neighbour across the edge
  const bool edge_x = s->cfg.open_bc && (s->xinc ? s->cfg.xrank == 0 : s->cfg.xrank == s->cfg.nxproc - 1);
  const bool edge_y = s->cfg.open_bc && (s->yinc ? s->cfg.yrank == 0 : s->cfg.yrank == s->cfg.nyproc - 1);
  const double *bcx = edge_x ? s->bcx : nullptr, *bcy = edge_y ? s->bcy : nullptr;
  if (edge_x || edge_y)
    DISPATCH_LC(s, LAUNCH(s, FAM_EDIR, (k_open_bc_columns<L, CT>), nblk(g.Nx + g.Ny, 64), 64, g, s->xinc, s->yinc, s->n1d, fac,
                           s->costheta, (int)edge_x, (int)edge_y, (const CT *)s->d2d, (const double *)s->kabs,
                           (const double *)s->dz, s->bcx, s->bcy, (int)(s->cfg.geometric_coeffs != 0 || s->nbld > 0), s->lutdev,
                           s->sym, (const double *)s->ksca, (const double *)s->gg));
  // The horizontal inputs of a pass are one pass old, so information needs one pass per column it crosses; the
  // norm is only read back every `check` passes to keep the host out of the loop.
  const int check = jacobi ? 4 : 2;
  const int NI = (g.Nx + 31) / 32, NJ = (g.Ny + EDIR_TY - 1) / EDIR_TY;
  int iter = 0;
  int rc = edir_exchange(s, xold);  // warm-start ghosts from the previous solution (:167)
  if (rc) return rc;
  while (iter < maxiter && !converged) {
    int nb = std::min(check, maxiter - iter);
    for (int q = 0; q < nb; ++q) {
      if (jacobi)
        DISPATCH_LC(s, LAUNCH(s, FAM_EDIR, (k_edir_pass<L, CT>), nblk(g.ncol, 128), 128, g, s->xinc, s->yinc, s->n1d, fac,
                               (const CT *)s->d2d, s->a33, xold, xnew, s->partials, s->counter, s->sc + SC_EDIR0 + q, bcx, bcy));
      else
        DISPATCH_LC(s, LAUNCH(s, FAM_EDIR, (k_edir_tile_pass<L, CT>), NI * NJ, 32 * EDIR_TY, g, s->xinc, s->yinc, s->n1d, fac,
                               (const CT *)s->d2d, s->a33, xold, xnew, s->partials, s->counter, s->sc + SC_EDIR0 + q, bcx, bcy));
      if ((rc = edir_exchange(s, xnew))) return rc;
      if (s->decomposed) {  // residual = mean over ranks of the local 2-norms (:176-183)
        LAUNCH(s, FAM_VEC, k_sqrt_mean, 1, 32, s->sc + SC_EDIR0 + q, 1, 1.0 / (double)s->nranks);
        if ((rc = allreduce_scalars(s, s->sc + SC_EDIR0 + q, 1))) return rc;
      }
      std::swap(xold, xnew);
    }
    if ((rc = read_scalars(s))) return rc;
    for (int q = 0; q < nb; ++q) {
      ++iter;
      double res = std::max(DBL_MIN, s->decomposed ? s->h_sc[SC_EDIR0 + q] : std::sqrt(s->h_sc[SC_EDIR0 + q]));
      if (iter == 1) res1 = res;
      double rel = (res1 <= std::sqrt(DBL_MIN)) ? 0.0 : res / res1;
      sol.res_dir = res;
      sol.niter_dir = iter;
      if (res < atol || rel < rtol) { converged = true; /* later passes of this batch only refine */ sol.niter_dir = iter + (nb - 1 - q); break; }
    }
  }
  if (xold != sol.edir) {  // result of the last pass lives in edir_tmp: swap the buffers
    s->edir_tmp = sol.edir;
    sol.edir = xold;
  }
  if ((edge_x && s->xinc) || (edge_y && s->yinc)) {  // the owned inflow faces keep what the last exchange received: zero
    const int cx = edge_x && s->xinc, cy = edge_y && s->yinc;
    const long n = (long)s->Nd * (g.Nz + 1) * (g.Nx + g.Ny);
    DISPATCH_L(s, LAUNCH(s, FAM_EDIR, (k_open_bc_clear<L>), nblk(n, 128), 128, g, cx, cy, sol.edir));
  }
  sol.first_dir = false;
  sol.lchanged = true;
  sol.lWm2_dir = false;
  if (!converged && !s->cfg.accept_incomplete_solve)  // src/pprts_explicit.F90:213-217
    return set_err(20, "solar_dir_ did not converge in %d passes (res %g); use accept_incomplete_solve to continue", iter, sol.res_dir);
  return 0;
}

static int setup_b(pprts_b200_solver *s, Solution &sol) {
  const Grid &g = s->g;
  CU(cudaMemsetAsync(s->b, 0, (size_t)state_len(g, s->Nf) * sizeof(double), s->stream));
  if (sol.lsolar_rad) {
    DISPATCH_LC(s, LAUNCH(s, FAM_SOURCE, (k_source_solar<L, CT>), nblk(g.ncell, 128), 128, g, s->xinc, s->yinc, s->n1d,
                           (const CT *)s->d2f, s->a13, s->a23, s->albedo, sol.edir, s->b));
  }
  if (sol.lthermal_rad) {
    DISPATCH_L(s, LAUNCH(s, FAM_SOURCE, (k_source_thermal<L>), nblk(g.ncell, 128), 128, g, s->n1d, (int)s->have_Bsrfc,
                          (const double *)(s->icollapse > 1 ? s->Btop : nullptr), (const double *)s->Bbot, s->cdiv_diff, s->a11, s->a12, s->kabs, s->dz, s->planck, s->Bsrfc, s->albedo, s->b));
  }
  if (s->nbld > 0) {  // set_buildings_reflection / set_buildings_emission (src/pprts.F90:4669-4672)
    DISPATCH_L(s, LAUNCH(s, FAM_SOURCE, (k_buildings_source<L>), nblk(s->nbld, 128), 128, g, (int)sol.lsolar_rad,
                          (int)sol.lthermal_rad, s->nbld, (const int *)s->bld_iface, (const double *)s->bld_albedo,
                          (const double *)(s->bld_have_planck ? s->bld_planck : nullptr), (const double *)s->dz,
                          (const double *)sol.edir, s->b));
  }
  // source emitted into a neighbour's face: stored there by its one producer (replaces halo_reduce_5pt, :4678)
  if (s->decomposed) {
    const unsigned int m = diff_inward_mask(s);
    int rc = halo_push<double>(s, s->b, s->Nf, diff_halo(s), true, true, m, m, 0, nullptr, nullptr);
    if (rc) return rc;
  }
  return 0;
}

template <class L, int NDOT, int NSTAGE, bool PREFETCH, class VT>
static bool launch_matmult_tma_v(pprts_b200_solver *s, const VT *x, VT *y, const VT *w, double *dots) {
  constexpr size_t smem = (size_t)NSTAGE * L::Nf * L::Nf * PB_TILE * sizeof(float);
  if constexpr (smem > 200 * 1024) return false;
  else {
    const Grid &g = s->g;
    static bool attr_set[64] = {false};  // the opt-in above 48 KB is per device (one process may drive several)
    auto kern = k_matmult_tma<L, NDOT, NSTAGE, PREFETCH, VT>;
    const int dev = s->device & 63;
    if (!attr_set[dev]) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
      attr_set[dev] = true;
    }
    const long ntiles = (g.ncell + PB_TILE - 1) / PB_TILE;
    const int per_sm = std::max(1, std::min<int>(NSTAGE == 1 ? 4 : 2, (int)((220 * 1024) / smem)));
    const unsigned int grid = (unsigned int)std::min<long>(ntiles, (long)per_sm * s->num_sms);
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (s->timing) { ea = get_event(s); eb = get_event(s); cudaEventRecord(ea, s->stream); }
    kern<<<grid, PB_TILE, smem, s->stream>>>(g, s->n1d, ntiles, (const float *)s->f2f, s->tile_class, s->a11, s->a12,
                                             s->albedo, x, y, w, s->partials, s->counter, dots);
    if (s->timing) { cudaEventRecord(eb, s->stream); s->evs.push_back({FAM_MATMULT, ea, eb}); }
    s->launches++;
    s->fam_launch[FAM_MATMULT]++;
    return true;
  }
}
template <class L, int NDOT, class VT>
static bool launch_matmult_tma(pprts_b200_solver *s, const VT *x, VT *y, const VT *w, double *dots) {
  if constexpr (std::is_same<VT, float>::value) return launch_matmult_tma_v<L, NDOT, 1, false, VT>(s, x, y, w, dots);
  else {
    switch (s->mm_variant) {
      case 0: return launch_matmult_tma_v<L, NDOT, 2, true, VT>(s, x, y, w, dots);
      case 1: return launch_matmult_tma_v<L, NDOT, 2, false, VT>(s, x, y, w, dots);
      default: return launch_matmult_tma_v<L, NDOT, 1, false, VT>(s, x, y, w, dots);
    }
  }
}

template <int NDOT, class VT = double>
static void launch_matmult_local(pprts_b200_solver *s, const VT *x, VT *y, const VT *w, double *dots) {
  const Grid &g = s->g;
  if (s->cfg.coeff_fp32 && s->use_tma) {
    bool done = false;
    DISPATCH_L(s, done = (launch_matmult_tma<L, NDOT, VT>(s, x, y, w, dots)));
    if (done) return;
  }
  DISPATCH_LC(s, LAUNCH(s, FAM_MATMULT, (k_matmult<L, CT, NDOT, VT>), nblk(g.ncell, 128), 128, g, s->n1d, (const CT *)s->f2f,
                         s->a11, s->a12, s->albedo, x, y, w, s->partials, s->counter, dots));
}

// y = (I - T) x on the decomposed domain: ghost fill of x, local kernel, rows produced for the neighbour pushed to
// their owner (which also adds their share of the fused dot products), dot products summed over the ranks
template <int NDOT, class VT = double>
static int launch_matmult(pprts_b200_solver *s, VT *x, VT *y, const VT *w, double *dots) {
  if (!s->decomposed) { launch_matmult_local<NDOT, VT>(s, x, y, w, dots); return 0; }
  const HaloDofs h = diff_halo(s);
  int rc = halo_fill<VT>(s, x, s->Nf, h, true, true);
  if (rc) return rc;
  launch_matmult_local<NDOT, VT>(s, x, y, w, dots);
  const unsigned int m = diff_inward_mask(s);
  if ((rc = halo_push<VT>(s, y, s->Nf, h, true, true, m, m, NDOT, w, dots))) return rc;
  if (NDOT > 0) rc = allreduce_scalars(s, dots, NDOT);
  return rc;
}

static bool precond_available(const pprts_b200_solver *s) { return s->cfg.precond >= 1; }
// PPRTS_B200_KRYLOV_TRACE: residual history of the diffuse solves on stderr (tools/krylov_trace.py)
static bool krylov_trace() { static const bool on = getenv("PPRTS_B200_KRYLOV_TRACE") != nullptr; return on; }

// Column kernels are one thread per column and latency-bound per thread, so their run time is (number of waves) x
// (time of one column): pick the block size (<= PREC_BLOCK) for which all columns are resident in the fewest waves.
// 256x256 columns with 64-thread blocks at 144 registers are 1024 blocks on 888 slots = 2 waves (ncu: 0.30 ms, 51 % of
// DRAM peak, profiles/r1_hot_kernels.txt); 32-thread blocks fit in one.
template <class K>
static int column_block(pprts_b200_solver *s, K kernel, int *cache) {
  if (*cache) return *cache;
  cudaFuncAttributes fa;
  int bs0 = PREC_BLOCK;  // a kernel compiled with __launch_bounds__(32, ...) cannot be launched with 64 threads
  if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) bs0 = std::min(bs0, std::max(32, fa.maxThreadsPerBlock / 32 * 32));
  int best = bs0;
  long best_waves = -1;
  for (int bs = bs0; bs >= 32; bs >>= 1) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, bs, 0) != cudaSuccess || per_sm < 1) continue;
    const long slots = (long)per_sm * s->num_sms, blocks = (s->g.ncol + bs - 1) / bs;
    const long waves = (blocks + slots - 1) / slots;
    if (best_waves < 0 || waves < best_waves) { best_waves = waves; best = bs; }
  }
  return *cache = best;
}

// once per set of coefficients: reflectances of the column solve, packed planes, Galerkin column blocks of the aggregates
template <class L, class CT>
static void prec_setup_t(pprts_b200_solver *s) {
  if constexpr (L::Nft == 2) {
    const Grid &g = s->g;
    if (!s->prec_ready) {
      LAUNCH(s, FAM_PRECOND, (k_prec_setup<L, CT>), nblk(g.ncol, PREC_BLOCK), PREC_BLOCK, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
             s->albedo, s->precR, s->precInv, s->precRf, s->precInvf);
      if (s->precPU)
        LAUNCH(s, FAM_PRECOND, (k_prec_pack<L, CT>), nblk(g.ncell, 128), 128, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
               (const float *)s->precRf, (const float *)s->precInvf, (__half2 *)s->precPU, (__half2 *)s->precPD);
      if (s->ncgl > 0) {
        for (int l0 = 0; l0 < s->ncgl; l0 += 2) {  // one pass over the fine coefficients per pair of nested levels
          CgLevels lv{};
          lv.n = std::min(2, s->ncgl - l0);
          for (int l = 0; l < lv.n; ++l) { lv.c[l] = s->cgl[l0 + l].c; lv.coef[l] = s->cgl[l0 + l].coef; }
          const CoarseGrid &top = lv.c[lv.n - 1];
          // tile-staged variant (k_cg_accumulate_tiles) where its layout conditions hold; PPRTS_B200_CG_ACC_WALK keeps the
          // per-aggregate walk for comparison
          const char *walk = getenv("PPRTS_B200_CG_ACC_WALK");
          bool tiles = walk == nullptr && std::is_same<CT, float>::value && g.Nx % PB_TILE == 0 && top.AG <= CG_ROWS_MAX;
          for (int l = 0; l < lv.n; ++l) tiles = tiles && (lv.c[l].AG & (lv.c[l].AG - 1)) == 0;
          bool done = false;
          if constexpr (sizeof(float) * CG_ROWS_MAX * L::Nf * PB_TILE <= 47 * 1024) {
            if (tiles) {
              const long nblocks = (long)(g.Nx / PB_TILE) * top.NyC * g.Nz * L::Nf;
              LAUNCH(s, FAM_PRECOND, (k_cg_accumulate_tiles<L>), (unsigned int)nblocks, PB_TILE, g, lv, s->n1d, (const float *)s->f2f, s->a11, s->a12);
              done = true;
            }
          }
          if (!done)
            LAUNCH(s, FAM_PRECOND, (k_cg_accumulate<L, CT>), nblk(top.ncc * L::Nf, 128), 128, g, lv, s->n1d, (const CT *)s->f2f, s->a11, s->a12);
        }
        for (int l = 0; l < s->ncgl; ++l) {
          const CoarseGrid &c = s->cgl[l].c;
          LAUNCH(s, FAM_PRECOND, (k_cg_condense<L>), nblk(c.ncc, 64), 64, c, s->cgl[l].coef);
          LAUNCH(s, FAM_PRECOND, k_cg_setup, nblk(c.ncolC, 64), 64, g, c, (const double *)s->albedo, (const float *)s->cgl[l].coef,
                 (float4 *)s->cgl[l].sA, (float2 *)s->cgl[l].sB, s->cgl[l].R);
        }
      }
      s->prec_ready = true;
    }
  }
}
// can the restriction of the coarse-grid correction ride on the vector update that produces the preconditioner's input?
static bool cg_fused_restrict(const pprts_b200_solver *s) {
  static const bool off = getenv("PPRTS_B200_CG_SEPARATE_RESTRICT") != nullptr;
  const Grid &g = s->g;
  return !off && s->ncgl > 0 && s->cgl[0].c.AG == 4 && (g.Nx & 3) == 0 && (g.Ny & 3) == 0;
}

template <class L, class CT, class VT>
static void launch_prec_t(pprts_b200_solver *s, const VT *r, VT *z, bool have_rbar = false) {
  if constexpr (L::Nft == 2) {
    const Grid &g = s->g;
    prec_setup_t<L, CT>(s);
    constexpr bool can_pack = std::is_same<VT, float>::value;
    const bool packed = can_pack && s->precPU;
    const int slot = (std::is_same<VT, float>::value ? 2 : 0) + (packed ? 2 : 0);
    const int bu = packed ? column_block(s, k_prec_up<L, CT, VT, can_pack>, &s->prec_block[slot])
                          : column_block(s, k_prec_up<L, CT, VT, false>, &s->prec_block[slot]);
    const int bd = packed ? column_block(s, k_prec_down<L, CT, VT, can_pack>, &s->prec_block[slot + 1])
                          : column_block(s, k_prec_down<L, CT, VT, false>, &s->prec_block[slot + 1]);
    const VT *pR, *pI;  // the sweeps read R and Inv in the precision of their vectors
    if constexpr (std::is_same<VT, float>::value) { pR = s->precRf; pI = s->precInvf; } else { pR = s->precR; pI = s->precInv; }
    // coarse-grid correction: xc = (P^T A P)_col^-1 P^T r first, the sweeps add P xc to z where they store its entries
    const int ncgl = s->cg_off ? 0 : s->ncgl;
    CgAdd add{nullptr, 0, 0, 1, 1};
    static const bool cg_unfused = getenv("PPRTS_B200_CG_UNFUSED") != nullptr;  // separate prolongation pass, for comparison
    for (int l = 0; l < ncgl; ++l) {  // restriction: the finest level reads r, the others their parent level
      auto &lv = s->cgl[l];
      const CoarseGrid &c = lv.c;
      if (l == 0) {
        if (!have_rbar)
          LAUNCH(s, FAM_PRECOND, (k_cg_restrict<L, VT>), nblk(c.nclev, 128), 128, g, c, r, (const float *)lv.coef, lv.rbar, (float2 *)lv.e);
      } else
        LAUNCH(s, FAM_PRECOND, (k_cg_restrict_coarse<L>), nblk(c.nclev, 128), 128, g, s->cgl[l - 1].c, c, (const float *)s->cgl[l - 1].rbar,
               (const float *)lv.coef, lv.rbar, (float2 *)lv.e);
    }
    for (int l = ncgl - 1; l >= 0; --l) {  // aggregate-column solves, coarsest first: each level hands its result down
      auto &lv = s->cgl[l];
      const CoarseGrid &c = lv.c;
      static const bool cg_serial = getenv("PPRTS_B200_CG_SERIAL_COLUMN") != nullptr;
      if (g.Nz <= 32 * CG_C && !cg_serial)
        LAUNCH(s, FAM_PRECOND, k_cg_column_scan, nblk(c.ncolC * 32, 128), 128, c, (const float4 *)lv.sA, (const float2 *)lv.sB,
               (const float *)lv.R, (const float *)lv.rbar, (const float2 *)lv.e, lv.x);
      else
        LAUNCH(s, FAM_PRECOND, k_cg_column, nblk(c.ncolC, 32), 32, c, (const float4 *)lv.sA, (const float2 *)lv.sB,
               (const float *)lv.R, (const float *)lv.rbar, (const float2 *)lv.e, lv.x);
      const bool coarser = l + 1 < ncgl;
      LAUNCH(s, FAM_PRECOND, (k_cg_sides<L>), nblk(c.nclev, 128), 128, c, (const float *)lv.coef, (const float *)lv.rbar, lv.x,
             coarser ? s->cgl[l + 1].c : c, coarser ? (const float *)s->cgl[l + 1].x : (const float *)nullptr);
    }
    if (ncgl > 0 && !cg_unfused) { const CoarseGrid &c = s->cgl[0].c; add = CgAdd{s->cgl[0].x, c.nclev, c.ncolC, c.AG, c.NxC}; }
    if (packed) {
      LAUNCH(s, FAM_PRECOND, (k_prec_up<L, CT, VT, can_pack>), nblk(g.ncol, bu), bu, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
             pR, pI, r, z, (int)(s->cfg.precond != 2), add, (const __half2 *)s->precPU);
      LAUNCH(s, FAM_PRECOND, (k_prec_down<L, CT, VT, can_pack>), nblk(g.ncol, bd), bd, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
             pR, pI, r, z, (int)(s->cfg.precond != 2), add, (const __half2 *)s->precPD);
    } else {
      LAUNCH(s, FAM_PRECOND, (k_prec_up<L, CT, VT, false>), nblk(g.ncol, bu), bu, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
             pR, pI, r, z, (int)(s->cfg.precond != 2), add, (const __half2 *)nullptr);
      LAUNCH(s, FAM_PRECOND, (k_prec_down<L, CT, VT, false>), nblk(g.ncol, bd), bd, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
             pR, pI, r, z, (int)(s->cfg.precond != 2), add, (const __half2 *)nullptr);
    }
    if (ncgl > 0 && cg_unfused)
      LAUNCH(s, FAM_PRECOND, (k_cg_prolong<VT>), vec_grid((long)L::Nf * g.nlev / 4), VEC_BLOCK, g, s->cgl[0].c, (int)L::Nf, (const float *)s->cgl[0].x, z);
  } else {  // several up/down pairs of top streams (3_24): block recurrence, no software pipelining
    const Grid &g = s->g;
    if (!s->prec_ready) {
      LAUNCH(s, FAM_PRECOND, (k_precb_setup<L, CT>), nblk(g.ncol, 32), 32, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
             s->albedo, s->precR, s->precInv, s->precRf, s->precInvf);
      s->prec_ready = true;
    }
    const VT *pR, *pI;
    if constexpr (std::is_same<VT, float>::value) { pR = s->precRf; pI = s->precInvf; } else { pR = s->precR; pI = s->precInv; }
    LAUNCH(s, FAM_PRECOND, (k_precb_up<L, CT, VT>), nblk(g.ncol, PRECB_BLOCK), PRECB_BLOCK, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
           pR, pI, r, z, (int)(s->cfg.precond != 2));
    LAUNCH(s, FAM_PRECOND, (k_precb_down<L, CT, VT>), nblk(g.ncol, PRECB_BLOCK), PRECB_BLOCK, g, s->n1d, (const CT *)s->f2f, s->a11, s->a12,
           pR, pI, r, z, (int)(s->cfg.precond != 2));
  }
}
static void launch_prec(pprts_b200_solver *s, const double *r, double *z) { DISPATCH_LC(s, (launch_prec_t<L, CT, double>(s, r, z))); }
static void launch_prec32(pprts_b200_solver *s, const float *r, float *z, bool have_rbar = false) { DISPATCH_LC(s, (launch_prec_t<L, CT, float>(s, r, z, have_rbar))); }
static void prec_setup(pprts_b200_solver *s) { DISPATCH_LC(s, (prec_setup_t<L, CT>(s))); }

// explicit_ediff (src/pprts_explicit.F90:461-713) analogue: stationary relaxation x += M^-1 (b - A x), all columns
// at once; residual = ||x_n - x_{n-1}||_2 (mean over ranks of the local norms, :620-626), converged when it is < atol
// or < rtol * residual(1) (:641-642); the first solve of a uid uses at most the default tolerances (:548-554).
static int solve_ediff_explicit(pprts_b200_solver *s, Solution &sol) {
  const Grid &g = s->g;
  const long n = (long)s->Nf * g.nlev;
  const int maxit = s->cfg.maxit_diff > 0 ? s->cfg.maxit_diff : 10000;
  double rtol_d, atol_d;
  determine_ksp_tolerances(s, rtol_d, atol_d);
  double rtol = s->cfg.rtol >= 0 ? s->cfg.rtol : rtol_d, atol = s->cfg.atol >= 0 ? s->cfg.atol : atol_d;
  if (sol.first_diff) { rtol = std::min(rtol, rtol_d); atol = std::min(atol, atol_d); }
  double *x = sol.ediff, *r = s->kr[0], *z = s->kr[1];
  const bool prec = precond_available(s);
  // x += M^-1 (b - A x) converges for the column splitting; the ADDITIVE coarse-grid correction counts the smooth modes
  // twice, which a Krylov method does not mind but a stationary iteration does: the relaxation uses the column solve alone
  struct CgOff { pprts_b200_solver *s; bool was; CgOff(pprts_b200_solver *s_) : s(s_), was(s_->cg_off) { s->cg_off = true; } ~CgOff() { s->cg_off = was; } } cg_guard(s);
  int rc = need_comm(s);
  if (rc) return rc;
  double first = 0.0;
  bool converged = false;
  sol.niter_diff = 0;
  for (int it = 1; it <= maxit && !converged; ++it) {
    if ((rc = launch_matmult<0, double>(s, x, r, nullptr, nullptr))) return rc;
    LAUNCH(s, FAM_VEC, k_relax_residual, vec_grid(n), VEC_BLOCK, n, s->b, r);
    if (prec) launch_prec(s, r, z);
    LAUNCH(s, FAM_VEC, k_relax_update, vec_grid(n), VEC_BLOCK, n, 1.0, prec ? z : r, x, s->partials, s->counter, s->sc + SC_TMP);
    if (s->decomposed) {  // mean over ranks of the local 2-norms, like the direct sweep
      LAUNCH(s, FAM_VEC, k_sqrt_mean, 1, 1, s->sc + SC_TMP, 1, 1.0 / s->nranks);
      if ((rc = allreduce_scalars(s, s->sc + SC_TMP, 1))) return rc;
    }
    if ((rc = read_scalars(s))) return rc;
    const double res = s->decomposed ? s->h_sc[SC_TMP] : std::sqrt(s->h_sc[SC_TMP]);
    if (it == 1) first = res;
    sol.niter_diff = it;
    sol.res_diff = res;
    if (!(res == res)) return set_err(21, "explicit diffuse solve broke down (NaN residual) at iteration %d", it);
    const double rel = first <= std::sqrt(DBL_MIN) ? 0.0 : res / first;
    if (res < atol || rel < rtol) converged = true;
  }
  sol.first_diff = false;
  sol.lchanged = true;
  sol.lWm2_diff = false;
  if (!converged) return set_err(22, "explicit diffuse solve did not converge -- %d iterations, residual %g", sol.niter_diff, sol.res_diff);
  return 0;
}


// Mixed-precision BiCGStab (cfg.krylov_fp32): the eight Krylov work vectors, the operator and the preconditioner
// applied to them are fp32 (half the vector traffic, FFMA instead of widening + DFMA); the solution x, the right-hand
// side and the residual the convergence test sees are fp64.  A cycle is an fp32 BiCGStab run started from
// r64 = b - A x (iterative refinement).  When the cycle's recursive residual meets the tolerance, the TRUE fp64 residual
// is evaluated: MyKSPConverged's test (src/pprts.F90:4437-4486) is only ever passed by a residual that carries no fp32
// rounding.  If the true residual is not there yet although the recursion says so, the same recursion continues with a
// tightened target (the Krylov space is kept); once the recursion has dropped by 2e-6 -- the range an fp32 recursion
// can be trusted over -- a new cycle starts from a fresh fp64 residual.  Typical solve at the default rtol = 1e-5: one
// cycle plus two fp64 MatMults.
static int solve_ediff_mixed(pprts_b200_solver *s, Solution &sol, double rtol, double atol, int maxit) {
  const Grid &g = s->g;
  const long n = (long)s->Nf * g.nlev, n4 = (state_len(g, s->Nf) + 3) & ~3L;  // fp32 vectors carry their ghost blocks too
  double *x = sol.ediff, *ax = s->kr[0];
  float *r = (float *)s->kr[1], *rh = r + n4, *p = (float *)s->kr[2], *v = p + n4, *sv = (float *)s->kr[3], *t = sv + n4;
  const bool prec = precond_available(s);
  float *ph = prec ? (float *)s->kr[4] : p, *sh = prec ? ph + n4 : sv;
  float *xc = (float *)s->kr[5];  // the cycle's correction of x, flushed into the fp64 solution before it is looked at
  int rc;
  double r0 = 0.0;
  bool dirty = false;
  // experiment switch (tools/mixed_check.py): residual replacement after a failed true-residual check.  Measured on a
  // 320x320x128, dx = 25 m domain: solar 254 -> 229 iterations but thermal 198 -> 286, so it is off by default.
  static const bool replace_r = getenv("PPRTS_B200_REPLACE_R") != nullptr;
  if (prec) prec_setup(s);  // the fused update + restriction kernels read the aggregates' coefficients
  const bool fused = prec && cg_fused_restrict(s);
  static const int flush_every = getenv("PPRTS_B200_FLUSH_EVERY") ? std::max(1, atoi(getenv("PPRTS_B200_FLUSH_EVERY"))) : (1 << 30);
  CU(cudaMemsetAsync(xc, 0, (size_t)n * sizeof(float), s->stream));
  int total = 0;
  bool converged = false;
  sol.niter_diff = 0;
  bool have_ax = false;  // ax = A x of the current x is already there (a true-residual check that led to a new cycle)
  // The ~20 launches of an iteration are replayed as ONE CUDA graph launch: small domains are launch-bound (configs[0]:
  // kernels of a few microseconds, 6.2 -> 5.5 ms per solve), and at configs[1] the shorter gaps between the kernels are
  // still worth 0.9 ms of a 93 ms step.  Not while kernels are being timed, not on decomposed solvers (NCCL calls inside).
  // PPRTS_B200_GRAPH_MAX_CELLS limits it to domains up to that many cells (0 switches it off).
  const char *gmc = getenv("PPRTS_B200_GRAPH_MAX_CELLS");  // read per solve: tests switch it
  const long graph_max_cells = gmc ? strtol(gmc, nullptr, 10) : LONG_MAX;
  const bool use_graph = !s->timing && !s->decomposed && !krylov_trace() && g.ncell <= graph_max_cells;
  const int graph_key = (prec ? 1 : 0) | (fused ? 2 : 0) | (s->cg_off ? 4 : 0) | (s->cfg.precond << 3);
  bool graphs_warm = false;
  for (int cycle = 0; !converged && cycle < 64; ++cycle) {
    if (cycle == 0 && sol.ediff_zero) {  // cold start: A x = 0 without looking (one fp64 MatMult per band)
      CU(cudaMemsetAsync(ax, 0, (size_t)n * sizeof(double), s->stream));
    } else if (!have_ax && (rc = launch_matmult<0, double>(s, x, ax, nullptr, nullptr))) return rc;
    have_ax = false;
    sol.ediff_zero = false;
    LAUNCH(s, FAM_VEC, k_mixed_init, vec_grid(n), VEC_BLOCK, n, s->b, ax, r, rh, s->partials, s->counter, s->sc);
    if ((rc = allreduce_scalars(s, s->sc + SC_RR, 3))) return rc;
    if ((rc = read_scalars(s))) return rc;
    const double rc0 = std::sqrt(s->h_sc[SC_RR]);  // true residual this cycle starts from
    if (cycle == 0) r0 = std::max(DBL_MIN, rc0);
    sol.res_diff = rc0;
    if (!(rc0 == rc0)) return set_err(21, "diffuse solve broke down (NaN residual) in refinement cycle %d", cycle);
    const double target = std::max(rtol * r0, atol);  // MyKSPConverged: rn/r0 <= rtol or rn <= atol
    // A TRUE residual at the rounding floor of b - A x cannot be reduced any further (e.g. a warm start from a solution
    // that is already converged, asked for another rtol): the reference's test, which sees a recursive residual, passes
    // there because the recursion keeps dropping; here it is the end of the solve as well.
    if (rc0 <= target || rc0 <= 1e-14 * std::sqrt(s->h_sc[SC_RHV])) { converged = true; break; }
    if (total >= maxit) break;
    const double floor32 = 2e-6 * rc0;
    double inner = 0.9 * target;  // a little below: the recursive residual drifts from the true one by O(eps32)
    bool restart = false;
    // the kernels of one iteration; they see `it` only as first / odd / even (which rho slot holds the new value)
    auto iteration = [&](int it) -> int {
      int irc;
      if (fused) {  // the update also leaves the restricted vector of the coarse-grid correction behind
        const CoarseGrid &c0 = s->cgl[0].c;
        DISPATCH_L(s, LAUNCH(s, FAM_VEC, (k_bicg32_update_restrict<L, 0>), nblk(c0.nclev, 128), 128, g, c0, it, (const float *)r, (const float *)v, p,
                              (const double *)s->sc, (const float *)s->cgl[0].coef, s->cgl[0].rbar, (float2 *)s->cgl[0].e));
      } else LAUNCH(s, FAM_VEC, k_bicg32_update_p, vec_grid(n), VEC_BLOCK, n, it, r, v, p, s->sc);
      if (prec) launch_prec32(s, p, ph, fused);
      if ((irc = launch_matmult<1, float>(s, ph, v, rh, s->sc + SC_RHV))) return irc;
      if (fused) {
        const CoarseGrid &c0 = s->cgl[0].c;
        DISPATCH_L(s, LAUNCH(s, FAM_VEC, (k_bicg32_update_restrict<L, 1>), nblk(c0.nclev, 128), 128, g, c0, it, (const float *)r, (const float *)v, sv,
                              (const double *)s->sc, (const float *)s->cgl[0].coef, s->cgl[0].rbar, (float2 *)s->cgl[0].e));
      } else LAUNCH(s, FAM_VEC, k_bicg32_update_s, vec_grid(n), VEC_BLOCK, n, it, r, v, sv, s->sc);
      if (prec) launch_prec32(s, sv, sh, fused);
      if ((irc = launch_matmult<2, float>(s, sh, t, sv, s->sc + SC_TS))) return irc;
      LAUNCH(s, FAM_VEC, k_bicg32_update_xr, vec_grid(n), VEC_BLOCK, n, it, ph, sh, sv, t, rh, xc, r, s->partials, s->counter, s->sc);
      return 0;
    };
    for (int it = 1; total < maxit && !restart && !converged; ++it) {
      if (use_graph && !s->iter_graph_off && (it > 1 || graphs_warm)) {
        // slot 0: first iteration of a cycle, 1: odd, 2: even; re-captured when the solver's preconditioner setting changed
        auto &ig = s->iter_graph[it == 1 ? 0 : ((it & 1) ? 1 : 2)];
        if (!ig.exec || ig.key != graph_key) {
          if (ig.exec) { cudaGraphExecDestroy(ig.exec); ig.exec = nullptr; }
          const long l0 = s->launches;
          long f0[FAM_COUNT];
          for (int q = 0; q < FAM_COUNT; ++q) f0[q] = s->fam_launch[q];
          cudaGraph_t gr = nullptr;
          bool ok = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
          if (ok) {
            const int irc = iteration(it);
            ok = cudaStreamEndCapture(s->stream, &gr) == cudaSuccess && irc == 0 && gr != nullptr;
          }
          if (ok) ok = cudaGraphInstantiate(&ig.exec, gr, 0) == cudaSuccess;
          if (gr) cudaGraphDestroy(gr);
          if (ok) {
            ig.key = graph_key;
            ig.launches = s->launches - l0;
            for (int q = 0; q < FAM_COUNT; ++q) ig.fam[q] = s->fam_launch[q] - f0[q];
          }
          // the capture only recorded the launches: undo their count, the graph launch below adds it
          s->launches = l0;
          for (int q = 0; q < FAM_COUNT; ++q) s->fam_launch[q] = f0[q];
          if (!ok) { cudaGetLastError(); ig.exec = nullptr; s->iter_graph_off = true; }
        }
        if (ig.exec) {
          CU(cudaGraphLaunch(ig.exec, s->stream));
          s->launches += ig.launches;
          for (int q = 0; q < FAM_COUNT; ++q) s->fam_launch[q] += ig.fam[q];
        } else if ((rc = iteration(it))) return rc;
      } else {
        if ((rc = iteration(it))) return rc;
        graphs_warm = true;  // everything an iteration sets up lazily (function attributes, block sizes) exists now
      }
      dirty = true;
      // The correction xc is an fp32 vector: left to grow over a whole cycle its rounding (6e-8 |x|) puts a floor of a few
      // 1e-5 r0 under the TRUE residual, right where the default tolerance sits (measured: recursive 1.3e-6, true 2.1e-5).
      // Flushed into the fp64 solution every few iterations it only ever holds the recent, small increments.  (Measured
      // on C2 with PPRTS_B200_FLUSH_EVERY = 2, 4: no change of the iteration counts -- the drift of the recursive residual
      // comes from the fp32 MatMults, not from xc -- so the flush stays at the true-residual checks by default.)
      if ((it % flush_every) == 0) { LAUNCH(s, FAM_VEC, k_mixed_flush, vec_grid(n), VEC_BLOCK, n, xc, x); dirty = false; }
      if ((rc = allreduce_scalars(s, s->sc + (((it + 1) & 1) ? SC_RR : SC_RHO0), 2))) return rc;
      if ((rc = read_scalars(s))) return rc;
      const double rn = std::sqrt(s->h_sc[SC_RR]);
      ++total;
      sol.niter_diff = total;
      sol.res_diff = rn;
      if (krylov_trace()) fprintf(stderr, "[krylov32] it %d rn/r0 %.3e\n", total, rn / r0);
      if (!(rn == rn)) return set_err(21, "diffuse solve broke down (NaN residual) at iteration %d", total);
      if (rn <= inner && inner > floor32) {  // the recursion claims convergence: check the true residual in fp64
        if (dirty) { LAUNCH(s, FAM_VEC, k_mixed_flush, vec_grid(n), VEC_BLOCK, n, xc, x); dirty = false; }
        if ((rc = launch_matmult<0, double>(s, x, ax, nullptr, nullptr))) return rc;
        LAUNCH(s, FAM_VEC, k_residual_norm2, vec_grid(n), VEC_BLOCK, n, s->b, ax, s->partials, s->counter, s->sc + SC_TMP);
        if ((rc = allreduce_scalars(s, s->sc + SC_TMP, 1))) return rc;
        if ((rc = read_scalars(s))) return rc;
        const double rtrue = std::sqrt(s->h_sc[SC_TMP]);
        sol.res_diff = rtrue;
        if (krylov_trace()) fprintf(stderr, "[krylov32] it %d true residual %.3e (recursive %.3e) target %.3e\n", total, rtrue / r0, rn / r0, target / r0);
        if (rtrue <= target) converged = true;
        else if (rtrue > 2.0 * rn && !replace_r) {
          // the recursive residual has drifted away from the true one (fp32 MatMults: measured on C2, recursive 1.3e-6 vs
          // true 2.1e-5 three iterations after recursive 6.1e-6 vs true 2.2e-5 -- the recursion no longer moves the true
          // residual): start the next cycle from the residual that was just evaluated instead of iterating on
          restart = true;
          have_ax = true;
        } else {
          inner = std::min(0.5 * inner, 0.5 * inner * target / rtrue);  // keep the Krylov space, aim lower
          if (replace_r) {  // ... and put the recursion back on the true residual (reliable update)
            LAUNCH(s, FAM_VEC, k_mixed_replace_r, vec_grid(n), VEC_BLOCK, n, it, (const double *)s->b, (const double *)ax,
                   (const float *)rh, r, s->partials, s->counter, s->sc);
            if ((rc = allreduce_scalars(s, s->sc + (((it + 1) & 1) ? SC_RR : SC_RHO0), 2))) return rc;
            inner = 0.9 * target;  // the recursive residual is the true one again
          }
        }
      }
      if (!converged && rn <= floor32) restart = true;  // refine from a fresh fp64 residual
      // breakdown: rho = (rh, r) vanished with r not converged -> new cycle, i.e. new shadow residual rh = r
      if (!converged && std::fabs(s->h_sc[((it + 1) & 1) ? SC_RHO1 : SC_RHO0]) <= 1e-30 * rn * rc0) restart = true;
    }
    if (dirty) { LAUNCH(s, FAM_VEC, k_mixed_flush, vec_grid(n), VEC_BLOCK, n, xc, x); dirty = false; }
    if (!restart && !converged) break;  // maxit
  }
  sol.first_diff = false;
  sol.lchanged = true;
  sol.lWm2_diff = false;
  if (!converged) return set_err(22, "***** SOLVER did NOT converge :( -- %d iterations, residual %g", sol.niter_diff, sol.res_diff);
  return 0;
}


// Restarted GMRES(m), right-preconditioned with the column solve, fp64: what the reference falls back to when the
// KSP solve did not converge -- "Resetted initial guess to zero and try again with gmres" (src/pprts.F90:4277-4296).
// Same tolerances and iteration limit as the failed solve (the KSP object keeps them), same convergence test
// (MyKSPConverged on the GMRES residual estimate, confirmed by the true residual at the end of every cycle).
static int solve_ediff_gmres(pprts_b200_solver *s, Solution &sol, double rtol, double atol, int maxit) {
  const Grid &g = s->g;
  const long n = (long)s->Nf * g.nlev, stride = (state_len(g, s->Nf) + 3) & ~3L;
  const int m = s->cfg.gmres_restart > 0 ? std::min(s->cfg.gmres_restart, 60) : 30;
  const bool prec = precond_available(s);
  double *x = sol.ediff, *w = s->kr[0], *z = s->kr[1], *u = s->kr[2];
  double *V = nullptr, *dh = nullptr;
  CU(cudaMalloc((void **)&V, (size_t)(m + 1) * stride * sizeof(double)));
  struct Guard { double *a, *b; ~Guard() { cudaFree(a); cudaFree(b); } } guard{V, nullptr};
  CU(cudaMemsetAsync(V, 0, (size_t)(m + 1) * stride * sizeof(double), s->stream));
  CU(cudaMalloc((void **)&dh, 128 * sizeof(double)));
  guard.b = dh;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gv(m + 1), y(m), hcol(m + 2);
  int rc, total = 0;
  double r0 = 0.0;
  bool converged = false;
  auto dots_to_host = [&](int cnt) -> int {  // dh[0..cnt) summed over the ranks -> hcol
    int e = allreduce_scalars(s, dh, cnt);
    if (e) return e;
    CU(cudaMemcpyAsync(hcol.data(), dh, cnt * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
  };
  sol.niter_diff = 0;
  for (int cycle = 0; !converged && total < maxit; ++cycle) {
    if ((rc = launch_matmult<0, double>(s, x, w, nullptr, nullptr))) return rc;
    LAUNCH(s, FAM_VEC, k_relax_residual, vec_grid(n), VEC_BLOCK, n, s->b, w);  // w = b - A x
    LAUNCH(s, FAM_VEC, (k_multi_dot<1>), vec_grid(n), VEC_BLOCK, n, (const double *)w, stride, 0, 1, (const double *)w, s->partials, s->counter, dh);
    if ((rc = dots_to_host(1))) return rc;
    const double beta = std::sqrt(hcol[0]);
    if (!(beta == beta)) return set_err(21, "GMRES retry broke down (NaN residual) in cycle %d", cycle);
    if (cycle == 0) r0 = std::max(DBL_MIN, beta);
    sol.res_diff = beta;
    if (beta / r0 <= rtol || beta <= atol) { converged = true; break; }
    LAUNCH(s, FAM_VEC, k_scale_to, vec_grid(n), VEC_BLOCK, n, 1.0 / beta, (const double *)w, V);
    std::fill(gv.begin(), gv.end(), 0.0);
    gv[0] = beta;
    int j = 0;
    for (; j < m && total < maxit; ++j) {
      double *vj = V + (size_t)j * stride, *vn = V + (size_t)(j + 1) * stride;
      if (prec) launch_prec(s, vj, z);
      if ((rc = launch_matmult<0, double>(s, prec ? z : vj, w, nullptr, nullptr))) return rc;
      std::vector<double> h(j + 2, 0.0);
      double wnorm2 = 0.0;
      for (int pass = 0; pass < 2; ++pass) {  // classical Gram-Schmidt, twice
        for (int j0 = 0; j0 <= j; j0 += 8) {
          const int nv = std::min(8, j + 1 - j0);
          LAUNCH(s, FAM_VEC, (k_multi_dot<8>), vec_grid(n), VEC_BLOCK, n, (const double *)V, stride, j0, nv, (const double *)w, s->partials, s->counter, dh + j0);
        }
        if ((rc = dots_to_host(j + 1))) return rc;
        for (int q = 0; q <= j; ++q) h[q] += hcol[q];
        LAUNCH(s, FAM_VEC, k_gmres_orth, vec_grid(n), VEC_BLOCK, n, (const double *)V, stride, j + 1, (const double *)dh, w, s->partials, s->counter, dh + 64);
        if ((rc = allreduce_scalars(s, dh + 64, 1))) return rc;
        CU(cudaMemcpyAsync(&wnorm2, dh + 64, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
      }
      h[j + 1] = std::sqrt(wnorm2);
      ++total;
      if (!(h[j + 1] == h[j + 1])) return set_err(21, "GMRES retry broke down (NaN) at iteration %d", total);
      if (h[j + 1] > 0.0) LAUNCH(s, FAM_VEC, k_scale_to, vec_grid(n), VEC_BLOCK, n, 1.0 / h[j + 1], (const double *)w, vn);
      for (int q = 0; q < j; ++q) {  // previous Givens rotations, then the new one
        const double t = cs[q] * h[q] + sn[q] * h[q + 1];
        h[q + 1] = -sn[q] * h[q] + cs[q] * h[q + 1];
        h[q] = t;
      }
      const double den = std::hypot(h[j], h[j + 1]);
      cs[j] = den > 0 ? h[j] / den : 1.0; sn[j] = den > 0 ? h[j + 1] / den : 0.0;
      h[j] = den; h[j + 1] = 0.0;
      gv[j + 1] = -sn[j] * gv[j];
      gv[j] = cs[j] * gv[j];
      for (int q = 0; q <= j; ++q) H[(size_t)q * m + j] = h[q];
      const double rn = std::fabs(gv[j + 1]);
      sol.niter_diff = total;
      sol.res_diff = rn;
      if (rn / r0 <= rtol || rn <= atol || den == 0.0) { ++j; break; }
    }
    // x += M^-1 V y,  H(0:j,0:j) y = g
    for (int q = j - 1; q >= 0; --q) {
      double acc = gv[q];
      for (int c = q + 1; c < j; ++c) acc -= H[(size_t)q * m + c] * y[c];
      y[q] = H[(size_t)q * m + q] != 0.0 ? acc / H[(size_t)q * m + q] : 0.0;
    }
    if (j > 0) {
      CU(cudaMemcpyAsync(dh, y.data(), j * sizeof(double), cudaMemcpyHostToDevice, s->stream));
      LAUNCH(s, FAM_VEC, k_lincomb, vec_grid(n), VEC_BLOCK, n, (const double *)V, stride, j, (const double *)dh, u);
      if (prec) launch_prec(s, u, z);
      LAUNCH(s, FAM_VEC, k_axpy, vec_grid(n), VEC_BLOCK, n, 1.0, (const double *)(prec ? z : u), x);
      CU(cudaStreamSynchronize(s->stream));  // y is host memory reused by the next cycle
    }
  }
  if (!converged) {  // the loop ended on the estimate or on maxit: the true residual decides
    if ((rc = launch_matmult<0, double>(s, x, w, nullptr, nullptr))) return rc;
    LAUNCH(s, FAM_VEC, k_residual_norm2, vec_grid(n), VEC_BLOCK, n, s->b, w, s->partials, s->counter, dh);
    if ((rc = dots_to_host(1))) return rc;
    sol.res_diff = std::sqrt(hcol[0]);
    converged = sol.res_diff / r0 <= rtol || sol.res_diff <= atol;
  }
  sol.first_diff = false;
  sol.lchanged = true;
  sol.lWm2_diff = false;
  return converged ? 0 : 22;
}

// Right-preconditioned BiCGStab on the matrix-free operator; convergence test of MyKSPConverged
// (src/pprts.F90:4437-4486): ||r||/||r0|| <= rtol or ||r|| <= atol, nonzero initial guess (:4343).
static int solve_ediff_krylov(pprts_b200_solver *s, Solution &sol, double rtol, double atol, int maxit) {
  const Grid &g = s->g;
  const long n = (long)s->Nf * g.nlev;
  double *x = sol.ediff, *r = s->kr[0], *rh = s->kr[1], *p = s->kr[2], *v = s->kr[3], *sv = s->kr[4], *t = s->kr[5];
  const bool prec = precond_available(s);
  double *ph = prec ? s->kr[6] : p, *sh = prec ? s->kr[7] : sv;
  int rc;
  if (s->fail_next_krylov > 0) {  // fault injection for the tests of the retry path (internal.h)
    --s->fail_next_krylov;
    sol.first_diff = false; sol.lchanged = true; sol.lWm2_diff = false;
    return set_err(22, "***** SOLVER did NOT converge :( -- injected failure");
  }
  if (s->cfg.krylov_fp32) return solve_ediff_mixed(s, sol, rtol, atol, maxit);
  if ((rc = launch_matmult<0, double>(s, x, r, nullptr, nullptr))) return rc;
  LAUNCH(s, FAM_VEC, k_bicg_init, vec_grid(n), VEC_BLOCK, n, s->b, r, rh, p, v, s->partials, s->counter, s->sc);
  if ((rc = allreduce_scalars(s, s->sc + SC_RR, 2))) return rc;
  if ((rc = read_scalars(s))) return rc;
  const double r0 = std::max(DBL_MIN, std::sqrt(s->h_sc[SC_RR]));
  sol.niter_diff = 0;
  sol.res_diff = r0;
  bool converged = false;
  if (r0 <= atol) converged = true;  // x already solves the system to tolerance (warm start)
  // `it` is the index inside the current Krylov recursion: a breakdown (rho = (rh, r) = 0 with r not converged, e.g.
  // when the source and its first images have disjoint supports in a vacuum) restarts it with rh = r
  for (int total = 1, it = 1; total <= maxit && !converged; ++total, ++it) {
    LAUNCH(s, FAM_VEC, k_bicg_update_p, vec_grid(n), VEC_BLOCK, n, it, r, v, p, s->sc);
    if (prec) launch_prec(s, p, ph);
    if ((rc = launch_matmult<1>(s, ph, v, rh, s->sc + SC_RHV))) return rc;
    LAUNCH(s, FAM_VEC, k_bicg_update_s, vec_grid(n), VEC_BLOCK, n, it, r, v, sv, s->sc);
    if (prec) launch_prec(s, sv, sh);
    if ((rc = launch_matmult<2>(s, sh, t, sv, s->sc + SC_TS))) return rc;
    LAUNCH(s, FAM_VEC, k_bicg_update_xr, vec_grid(n), VEC_BLOCK, n, it, ph, sh, sv, t, rh, x, r, s->partials, s->counter, s->sc);
    if ((rc = allreduce_scalars(s, s->sc + (((it + 1) & 1) ? SC_RR : SC_RHO0), 2))) return rc;
    if ((rc = read_scalars(s))) return rc;
    const double rn = std::sqrt(s->h_sc[SC_RR]);
    sol.niter_diff = total;
    sol.res_diff = rn;
    if (krylov_trace()) fprintf(stderr, "[krylov64] it %d rn/r0 %.3e\n", total, rn / r0);
    if (!(rn == rn)) return set_err(21, "diffuse solve broke down (NaN residual) at iteration %d", total);
    if (rn / r0 <= rtol || rn <= atol) converged = true;
    else if (std::fabs(s->h_sc[((it + 1) & 1) ? SC_RHO1 : SC_RHO0]) <= 1e-30 * rn * r0) {
      LAUNCH(s, FAM_VEC, k_bicg_restart, vec_grid(n), VEC_BLOCK, n, r, rh, s->sc);
      it = 0;
    }
  }
  sol.first_diff = false;
  sol.lchanged = true;
  sol.lWm2_diff = false;
  if (!converged) return set_err(22, "***** SOLVER did NOT converge :( -- %d iterations, residual %g", sol.niter_diff, sol.res_diff);
  return 0;
}


// solve (src/pprts.F90:4206-4303): the Krylov solve and what the reference does when it fails -- with
// -accept_incomplete_solve the iterate is taken as it is; otherwise the initial guess is reset to zero and the system
// is solved again with GMRES, and only if that fails too the error is raised.
static int solve_ediff(pprts_b200_solver *s, Solution &sol) {
  const Grid &g = s->g;
  const int maxit = s->cfg.maxit_diff > 0 ? s->cfg.maxit_diff : 1000;
  double rtol_d, atol_d;
  determine_ksp_tolerances(s, rtol_d, atol_d);
  double rtol = s->cfg.rtol >= 0 ? s->cfg.rtol : rtol_d, atol = s->cfg.atol >= 0 ? s->cfg.atol : atol_d;
  if (sol.first_diff) { rtol = std::min(rtol, rtol_d); atol = std::min(atol, atol_d); }
  int rc = need_comm(s);
  if (rc) return rc;
  if (s->cfg.diff_explicit) {  // explicit solvers: src/pprts_explicit.F90:692-696
    rc = solve_ediff_explicit(s, sol);
    return (rc == 22 && s->cfg.accept_incomplete_solve) ? 0 : rc;
  }
  rc = solve_ediff_krylov(s, sol, rtol, atol, maxit);
  if (rc != 21 && rc != 22) return rc;
  if (s->cfg.accept_incomplete_solve) return 0;
  s->krylov_retries++;
  CU(cudaMemsetAsync(sol.ediff, 0, (size_t)state_len(g, s->Nf) * sizeof(double), s->stream));  // VecSet(x, zero)
  rc = solve_ediff_gmres(s, sol, rtol, atol, maxit);
  if (rc == 22)
    return set_err(22, "***** SOLVER did NOT converge :( -- %d iterations, residual %g; if you know what you are doing, you "
                       "can use the option -accept_incomplete_solve to continue", sol.niter_diff, sol.res_diff);
  return rc;
}

static int calc_flx_div(pprts_b200_solver *s, Solution &sol) {
  const Grid &g = s->g;
  int rc = scale_flx(s, sol, false);
  if (rc) return rc;
  if (s->decomposed) {  // the cell balance reads the streams entering through the high faces
    if ((rc = halo_fill(s, sol.ediff, s->Nf, diff_halo(s), true, true))) return rc;
    if (sol.lsolar_rad && (rc = halo_fill(s, sol.edir, s->Nd, dir_halo(s), s->xinc == 0, s->yinc == 0))) return rc;
  }
  DISPATCH_L(s, LAUNCH(s, FAM_ABSO, (k_absorption<L>), nblk(g.ncell, 128), 128, g, s->xinc, s->yinc, s->n1d,
                        (int)sol.lsolar_rad, (int)sol.lthermal_rad, s->cfg.by_coeff_divergence, s->costheta, s->cdiv_dir,
                        s->cdiv_diff, s->a11, s->a12, s->kabs, s->dz, sol.edir, sol.ediff, s->b, sol.abso));
  return 0;
}

static int get_solution(pprts_b200_solver *s, int uid, bool lsolar, bool lthermal, Solution **out) {
  const Grid &g = s->g;
  const int key = s->cfg.keep_solutions ? uid : 0;
  // max_cached_solutions: a band loop over hundreds of g-points on a large domain cannot keep every solution in HBM
  // (3 GB each at 512x512x100); beyond the cap the least recently used ones are dropped -- never `uid` itself or
  // uid - 1, the warm start of a first solve.  (The reference's answer to the same problem is bf16 compression of the
  // stored solutions, -pprts_compress_solutions, src/pprts_base.F90:411-502.)
  if (s->cfg.keep_solutions && s->cfg.max_cached_solutions > 0 && !s->solutions.count(key)) {
    while ((int)s->solutions.size() >= s->cfg.max_cached_solutions) {
      int victim = 0; long oldest = -1; bool found = false;
      for (auto &kv : s->solutions)
        if (kv.first != uid && kv.first != uid - 1 && (!found || kv.second.last_use < oldest)) { victim = kv.first; oldest = kv.second.last_use; found = true; }
      if (!found) break;
      // the victim's arrays go to a spare list instead of back to the driver: freeing and re-allocating 3 GB per band
      // costs tens of milliseconds (cudaFree synchronises the device); work still queued on them is stream-ordered
      // before whatever the next owner enqueues
      Solution &v = s->solutions[victim];
      s->spare.push_back({v.edir, v.ediff, v.abso});
      v.edir = v.ediff = v.abso = nullptr;
      s->solutions.erase(victim);
    }
  }
  Solution &sol = s->solutions[key];
  sol.last_use = ++s->use_clock;
  auto prepare = [&]() -> int {  // prepare_solution (src/pprts_base.F90:297-344)
    bool fresh = false;
    if (!sol.edir) {
      if (!s->spare.empty()) { sol.edir = s->spare.back()[0]; sol.ediff = s->spare.back()[1]; sol.abso = s->spare.back()[2]; s->spare.pop_back(); }
      else { DA(sol.edir, (size_t)state_len(g, s->Nd)); DA(sol.ediff, (size_t)state_len(g, s->Nf)); DA(sol.abso, g.ncell); fresh = true; }
    }
    if (!fresh) {
      CU(cudaMemsetAsync(sol.edir, 0, (size_t)state_len(g, s->Nd) * sizeof(double), s->stream));
      CU(cudaMemsetAsync(sol.ediff, 0, (size_t)state_len(g, s->Nf) * sizeof(double), s->stream));
      CU(cudaMemsetAsync(sol.abso, 0, (size_t)g.ncell * sizeof(double), s->stream));
    }
    sol.lset = true; sol.lsolar_rad = lsolar; sol.lthermal_rad = lthermal; sol.lchanged = true;
    sol.ediff_zero = true;
    sol.lWm2_dir = sol.lWm2_diff = true; sol.first_dir = sol.first_diff = true;
    return 0;
  };
  if (!sol.lset || !s->cfg.keep_solutions) {
    int rc = prepare();
    if (rc) return rc;
    if (s->cfg.keep_solutions) {  // -initial_guess_from_last_uid (src/pprts.F90:2542-2558)
      auto it = s->solutions.find(uid - 1);
      if (it != s->solutions.end() && it->second.lset) {
        Solution &last = it->second;
        if (last.lsolar_rad && lsolar) {
          CU(cudaMemcpyAsync(sol.edir, last.edir, (size_t)s->Nd * g.nlev * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
          sol.lWm2_dir = last.lWm2_dir;
        }
        CU(cudaMemcpyAsync(sol.ediff, last.ediff, (size_t)s->Nf * g.nlev * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        sol.ediff_zero = false;
        sol.lWm2_diff = last.lWm2_diff;
      }
    }
  } else if (sol.lsolar_rad != lsolar) {  // :2578-2582
    int rc = prepare();
    if (rc) return rc;
  }
  *out = &sol;
  return 0;
}

static int solve_impl(pprts_b200_solver *s, int lthermal, int lsolar, double edirTOA, int uid, bool have_time, double time);
extern "C" int pprts_b200_solve(pprts_b200_solver *s, int lthermal, int lsolar, double edirTOA, int uid) {
  return solve_impl(s, lthermal, lsolar, edirTOA, uid, false, 0.0);
}
// solve_pprts(..., opt_solution_time = time): restore_solution also records the time and, with the error estimates
// enabled, the inf-norm of the change of the absorption (src/pprts.F90:4005-4049)
extern "C" int pprts_b200_solve_at_time(pprts_b200_solver *s, int lthermal, int lsolar, double edirTOA, int uid, double time) {
  return solve_impl(s, lthermal, lsolar, edirTOA, uid, true, time);
}
static int solve_impl(pprts_b200_solver *s, int lthermal, int lsolar, double edirTOA, int uid, bool have_time, double time) {
  if (!s) return set_err(1, "null solver");
  if (!s->have_optprop) return set_err(10, "atmosphere%%kabs is not allocated! - maybe you need to call set_optical_properties() first");
  if (uid < -1000 || uid > 1000) return set_err(11, "solution uid %d out of range", uid);
  if (lthermal && !s->have_planck)
    return set_err(12, "you asked to compute thermal radiation but did not provide planck emissions in set_optical_properties");
  CU(cudaSetDevice(s->device));
  const bool derived_lsolar = lsolar && edirTOA > 0.0 && s->theta >= 0.0;
  if (derived_lsolar && lthermal)
    return set_err(13, "Somehow ended up with a request to compute solar and thermal radiation in one call.");
  Solution *psol;
  int rc = get_solution(s, uid, derived_lsolar, lthermal != 0, &psol);
  if (rc) return rc;
  Solution &sol = *psol;
  const Grid &g = s->g;
  s->res_valid = false;
  if (!sol.lthermal_rad && !sol.lsolar_rad) {  // nothing to do (src/pprts.F90:2590-2596)
    CU(cudaMemsetAsync(sol.edir, 0, (size_t)s->Nd * g.nlev * sizeof(double), s->stream));
    CU(cudaMemsetAsync(sol.ediff, 0, (size_t)s->Nf * g.nlev * sizeof(double), s->stream));
    CU(cudaMemsetAsync(sol.abso, 0, (size_t)g.ncell * sizeof(double), s->stream));
    sol.lchanged = false;
    return 0;
  }
  if ((rc = need_comm(s))) return rc;
  if ((rc = ensure_coeffs(s, sol.lsolar_rad))) return rc;
  if ((rc = scale_flx(s, sol, false))) return rc;
  if (!sol.lsolar_rad) { sol.niter_dir = 0; sol.res_dir = 0.0; }  // a reused container must not report the previous band's direct solve
  if (sol.lsolar_rad && (rc = solve_edir(s, sol, edirTOA))) return rc;
  if ((rc = setup_b(s, sol))) return rc;
  // b_norm = sum(abs(b)); skip the diffuse solve if < atol (src/pprts.F90:2774-2785)
  const long n = (long)s->Nf * g.nlev;
  LAUNCH(s, FAM_VEC, k_norm1, vec_grid(n), VEC_BLOCK, n, s->b, s->partials, s->counter, s->sc + SC_BNORM1);
  if ((rc = allreduce_scalars(s, s->sc + SC_BNORM1, 1))) return rc;
  if ((rc = read_scalars(s))) return rc;
  double rtol, atol;
  determine_ksp_tolerances(s, rtol, atol);
  if (s->h_sc[SC_BNORM1] < atol) {
    CU(cudaMemcpyAsync(sol.ediff, s->b, n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    sol.niter_diff = 0;
    sol.res_diff = atol;
    sol.ediff_zero = false;
  } else {
    rc = solve_ediff(s, sol);
    sol.ediff_zero = false;
    if (rc) return rc;
  }
  sol.lchanged = true;
  sol.lWm2_diff = false;
  // restore_solution (src/pprts.F90:3991-4049)
  const bool lenable = s->cfg.max_solution_err > 0.0 && s->cfg.max_solution_time > 0.0;  // :456-457
  if (have_time) {  // eoshift(time, -1); time(1) = time
    for (int k = 29; k > 0; --k) sol.time[k] = sol.time[k - 1];
    sol.time[0] = time;
  }
  const bool track = have_time && lenable;
  if (track) CU(cudaMemcpyAsync(s->out_stage, sol.abso, (size_t)g.ncell * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  if ((rc = calc_flx_div(s, sol))) return rc;
  if ((rc = scale_flx(s, sol, true))) return rc;
  sol.lchanged = false;
  if (track) {
    CU(cudaMemsetAsync(s->sc + SC_TMP, 0, sizeof(double), s->stream));
    LAUNCH(s, FAM_ABSO, k_absdiff_max, vec_grid(g.ncell), VEC_BLOCK, g.ncell, (const double *)s->out_stage,
           (const double *)sol.abso, s->sc + SC_TMP);
    if (s->decomposed) NC(g_nccl.AllReduce(s->sc + SC_TMP, s->sc + SC_TMP, 1, NCCL_F64, NCCL_MAX, s->nccl_comm, s->stream));
    if ((rc = read_scalars(s))) return rc;
    for (int k = 29; k > 0; --k) sol.maxnorm[k] = sol.maxnorm[k - 1];
    sol.maxnorm[0] = s->h_sc[SC_TMP];
  }
  CU(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// error-growth tracking: need_new_solution / polyfit (src/adaptive_spectral_integration.F90:38-257), host logic
// ---------------------------------------------------------------------------------------------------------------
static bool approx_eq(double a, double b) {  // helper_functions.fypp:1272-1287, default precision 10 * epsilon
  const double eps = 10.0 * DBL_EPSILON;
  return a <= b + eps && a >= b - eps;
}
// least squares through the normal equations, (X^T X)^-1 by Gauss-Jordan with partial pivoting; 0 on success
static int polyfit(const double *vx, const double *vy, int n, int d, double *coef) {
  double X[16][4], A[4][4], inv[4][4];
  const int m = d + 1;
  if (n > 16 || m > 4) return 9;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (approx_eq(vx[i], vx[j])) return 1;  // equal abscissae: the matrix is singular (:198-207)
  for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) X[j][i] = std::pow(vx[j], (double)i);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += X[k][i] * X[k][j];
      A[i][j] = acc; inv[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int p = 0; p < m; ++p) {
    int piv = p;
    for (int i = p + 1; i < m; ++i) if (std::fabs(A[i][p]) > std::fabs(A[piv][p])) piv = i;
    if (A[piv][p] == 0.0) return 2;
    if (piv != p) for (int j = 0; j < m; ++j) { std::swap(A[p][j], A[piv][j]); std::swap(inv[p][j], inv[piv][j]); }
    const double ip = 1.0 / A[p][p];
    for (int j = 0; j < m; ++j) { A[p][j] *= ip; inv[p][j] *= ip; }
    for (int i = 0; i < m; ++i) {
      if (i == p) continue;
      const double f = A[i][p];
      for (int j = 0; j < m; ++j) { A[i][j] -= f * A[p][j]; inv[i][j] -= f * inv[p][j]; }
    }
  }
  for (int i = 0; i < m; ++i) {
    double acc = 0.0;
    for (int j = 0; j < n; ++j) {
      double mij = 0.0;
      for (int k = 0; k < m; ++k) mij += inv[i][k] * X[j][k];
      acc += mij * vy[j];
    }
    coef[i] = acc;
  }
  return 0;
}
// polyfit (src/adaptive_spectral_integration.F90:174-257) as a plain host function: usable without a device
extern "C" int pprts_b200_polyfit(const double *vx, const double *vy, int n, int d, double *coef) {
  if (!vx || !vy || !coef || n < 1 || d < 0) return 9;
  return polyfit(vx, vy, n, d, coef);
}
// need_new_solution(comm, solver%solutions(uid), time, lenable_solutions_err_estimates): *need = 1 if the band has to
// be solved again at `time`, 0 if the stored solution is still within -max_solution_err.  have_time = 0: time absent.
extern "C" int pprts_b200_need_new_solution(pprts_b200_solver *s, int uid, double time, int have_time, int *need) {
  if (!s || !need) return set_err(1, "null solver");
  constexpr int Nfit = 3, Nporder = 2;
  *need = 1;
  if (!have_time) return 0;  // 'time not provided'
  if (!(s->cfg.max_solution_err > 0.0 && s->cfg.max_solution_time > 0.0)) return 0;  // 'solutions_err_estimates disabled'
  auto it = s->solutions.find(s->cfg.keep_solutions ? uid : 0);
  if (it == s->solutions.end() || !it->second.lset) return 0;  // 'no solution yet -- not initialized'
  const Solution &sol = it->second;
  double t[Nfit], tm[Nfit], ex[2 * (Nfit - 1)], ey[2 * (Nfit - 1)], polyc[Nporder + 1], est = 0.0;
  for (int k = 1; k <= Nfit; ++k) t[k - 1] = sol.time[Nfit - k];
  for (int k = 1; k <= Nfit - 1; ++k) tm[k - 1] = (t[k - 1] + t[k]) * .5;
  tm[Nfit - 1] = (t[Nfit - 1] + time) * .5;
  for (int k = 1; k <= Nfit - 1; ++k) {  // error time series: zero in between two updates, the recorded norm at an update
    ex[2 * k - 2] = tm[k - 1]; ey[2 * k - 2] = 0.0;
    ex[2 * k - 1] = t[k];      ey[2 * k - 1] = sol.maxnorm[Nfit - k - 1];
  }
  for (int ipoly = 1; ipoly <= Nporder; ++ipoly) {
    if (polyfit(ex, ey, 2 * (Nfit - 1), ipoly, polyc)) return 0;  // 'problem fitting error curve'
    double e = 0.0, tp = 1.0;
    for (int k = 0; k <= ipoly; ++k) { e += polyc[k] * tp; tp *= time; }
    est = std::max(est, std::fabs(e));
  }
  int n = !(est <= s->cfg.max_solution_err);                 // ERR_TOL_IN_BOUND / ERR_TOL_EXCEEDED
  for (int k = 0; k < Nfit; ++k) if (t[k] < 0.0) n = 1;      // FEW_SOLUTIONS
  if (time - sol.time[0] > s->cfg.max_solution_time) n = 1;  // MIN_TIME_EXCEEDED
  *need = n;
  return 0;
}
extern "C" int pprts_b200_get_solution_times(pprts_b200_solver *s, int uid, double time[30], double maxnorm[30]) {
  if (!s) return set_err(1, "null solver");
  auto it = s->solutions.find(s->cfg.keep_solutions ? uid : 0);
  if (it == s->solutions.end()) return set_err(30, "no solution for uid %d", uid);
  for (int k = 0; k < 30; ++k) { time[k] = it->second.time[k]; maxnorm[k] = it->second.maxnorm[k]; }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------------------------
static int find_solution(pprts_b200_solver *s, int uid, Solution **out) {
  auto it = s->solutions.find(s->cfg.keep_solutions ? uid : 0);
  if (it == s->solutions.end() || !it->second.lset) return set_err(30, "no solution for uid %d", uid);
  if (it->second.lchanged) return set_err(31, "tried to get results from unrestored solution -- call restore_solution first");
  if (!it->second.lWm2_diff) return set_err(32, "tried to get result from a result vector(diff) which is not in [W/m2]");
  *out = &it->second;
  return 0;
}

static int compute_result(pprts_b200_solver *s, int uid, bool accumulate, double weight) {
  Solution *sol;
  int rc = find_solution(s, uid, &sol);
  if (rc) return rc;
  const Grid &g = s->g;
  double **dst = accumulate ? s->acc : s->res;
  DISPATCH_L(s, LAUNCH(s, FAM_ABSO, (k_result<L>), nblk(g.nlev, 256), 256, g, (int)sol->lsolar_rad, s->mu, sol->edir,
                        sol->ediff, sol->abso, dst[0], dst[1], dst[2], dst[3], (int)accumulate, weight));
  if (!accumulate) { s->res_valid = true; s->res_uid = uid; }
  return 0;
}

static int download_fields(pprts_b200_solver *s, double **src, double *edn, double *eup, double *abso, double *edir) {
  const Grid &g = s->g;
  double *hosts[4] = {edn, eup, edir, abso};
  for (int q = 0; q < 4; ++q) {
    if (!hosts[q]) continue;
    const int nz = q == 3 ? g.Nz : g.Nz + 1;
    dim3 grid((g.Nx + 31) / 32, (nz + 31) / 32, g.Ny), block(32, 8);
    LAUNCH(s, FAM_IO, k_kji_to_zxy<double>, grid, block, nz, g.Nx, g.Ny, src[q], s->out_stage, 1.0);
    CU(cudaMemcpyAsync(hosts[q], s->out_stage, (size_t)nz * g.ncol * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

extern "C" int pprts_b200_compute_result_device(pprts_b200_solver *s, int uid) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  return compute_result(s, uid, false, 1.0);
}
extern "C" int pprts_b200_synchronize(pprts_b200_solver *s) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int pprts_b200_get_result(pprts_b200_solver *s, int uid, double *edn, double *eup, double *abso, double *edir) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  int rc = compute_result(s, uid, false, 1.0);
  if (rc) return rc;
  return download_fields(s, s->res, edn, eup, abso, edir);
}
// pprts_get_result whose device->host copies run on the copy stream: the call returns as soon as the result reduction
// and the layout conversion are enqueued, the solver's stream is free for the next band while the four fields travel.
// The host arrays (page-locked for the copy to be asynchronous) are complete after pprts_b200_wait_results.
extern "C" int pprts_b200_get_result_async(pprts_b200_solver *s, int uid, double *edn, double *eup, double *abso, double *edir) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  int rc = ensure_copy_stream(s);
  if (rc) return rc;
  const Grid &g = s->g;
  if (!s->egress[0]) {
    for (int q = 0; q < 4; ++q) DA(s->egress[q], (size_t)g.nlev);
    CU(cudaEventCreateWithFlags(&s->egress_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->egress_done, cudaEventDisableTiming));
  }
  if ((rc = compute_result(s, uid, false, 1.0))) return rc;
  if (s->egress_pending) CU(cudaStreamWaitEvent(s->stream, s->egress_done, 0));  // the previous copies still read egress
  double *hosts[4] = {edn, eup, edir, abso};
  for (int q = 0; q < 4; ++q) {
    if (!hosts[q]) continue;
    const int nz = q == 3 ? g.Nz : g.Nz + 1;
    dim3 grid((g.Nx + 31) / 32, (nz + 31) / 32, g.Ny), block(32, 8);
    LAUNCH(s, FAM_IO, k_kji_to_zxy<double>, grid, block, nz, g.Nx, g.Ny, s->res[q], s->egress[q], 1.0);
  }
  CU(cudaEventRecord(s->egress_ready, s->stream));
  CU(cudaStreamWaitEvent(s->copy_stream, s->egress_ready, 0));
  for (int q = 0; q < 4; ++q) {
    if (!hosts[q]) continue;
    const int nz = q == 3 ? g.Nz : g.Nz + 1;
    CU(cudaMemcpyAsync(hosts[q], s->egress[q], (size_t)nz * g.ncol * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream));
  }
  CU(cudaEventRecord(s->egress_done, s->copy_stream));
  s->egress_pending = true;
  return 0;
}
extern "C" int pprts_b200_wait_results(pprts_b200_solver *s) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  if (s->copy_stream) CU(cudaStreamSynchronize(s->copy_stream));
  s->egress_pending = false;
  return 0;
}
extern "C" int pprts_b200_accumulate_result(pprts_b200_solver *s, int uid, double weight) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  return compute_result(s, uid, true, weight);
}
extern "C" int pprts_b200_reset_accumulator(pprts_b200_solver *s) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  for (int q = 0; q < 4; ++q) CU(cudaMemsetAsync(s->acc[q], 0, (size_t)s->g.nlev * sizeof(double), s->stream));
  return 0;
}
extern "C" int pprts_b200_get_accumulated(pprts_b200_solver *s, double *edn, double *eup, double *abso, double *edir) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  return download_fields(s, s->acc, edn, eup, abso, edir);
}

// ---------------------------------------------------------------------------------------------------------------
// band loop (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053, :621-708): upload of band ib+1 overlaps the solve of band ib
// ---------------------------------------------------------------------------------------------------------------
static const int BAND_SLOT0 = 1 << 30;  // staging slots reserved for the band loop (user slots are small integers)
static int stage_band_async(pprts_b200_solver *s, Staged &st, const pprts_b200_bands *b, int ib) {
  const Grid &gr = s->g;
  auto up = [&](double **dst, const double *src, size_t n) -> int {
    if (!*dst) CU(cudaMalloc((void **)dst, n * sizeof(double)));
    CU(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, s->copy_stream));
    return 0;
  };
  int rc;
  st.albedo = b->albedo;
  if ((rc = up(&st.kabs, b->kabs + (long)ib * b->stride_kabs, s->ncell_atm))) return rc;
  if ((rc = up(&st.ksca, b->ksca + (long)ib * b->stride_ksca, s->ncell_atm))) return rc;
  if ((rc = up(&st.g, b->g + (long)ib * b->stride_g, s->ncell_atm))) return rc;
  st.have_planck = b->planck != nullptr;
  if (b->planck && (rc = up(&st.planck, b->planck + (long)ib * b->stride_planck, s->nlev_atm))) return rc;
  st.have_Bsrfc = b->planck_srfc != nullptr;
  if (b->planck_srfc && (rc = up(&st.Bsrfc, b->planck_srfc + (long)ib * b->stride_planck_srfc, gr.ncol))) return rc;
  st.have_albedo2d = b->albedo_2d != nullptr;
  if (b->albedo_2d && (rc = up(&st.albedo2d, b->albedo_2d, gr.ncol))) return rc;
  return 0;
}

extern "C" int pprts_b200_solve_bands(pprts_b200_solver *s, const pprts_b200_bands *b, double *edn, double *eup,
                                      double *abso, double *edir, int *niter_diff) {
  if (!s) return set_err(1, "null solver");
  if (!b || b->nbands < 0) return set_err(2, "solve_bands: bad band description");
  if (b->nbands > 0 && (!b->kabs || !b->ksca || !b->g)) return set_err(2, "solve_bands: kabs, ksca and g are required");
  if (b->nbands > 0 && b->lsolar && !b->edirTOA) return set_err(2, "solve_bands: solar bands need edirTOA(nbands)");
  if (b->nbands > 0 && !b->lsolar && !b->planck)
    return set_err(12, "you asked to compute thermal radiation but did not provide planck emissions in set_optical_properties");
  if (b->uid0 < -1000 || b->uid0 + std::max(0, b->nbands - 1) > 1000) return set_err(11, "solution uids %d.. out of range", b->uid0);
  CU(cudaSetDevice(s->device));
  int rc = ensure_copy_stream(s);
  if (rc) return rc;
  if (!b->accumulate_only && (rc = pprts_b200_reset_accumulator(s))) return rc;
  // Adaptive spectral integration (rrtmg/rrtmg/pprts_rrtmg.F90:1001, src/adaptive_spectral_integration.F90:38): with a
  // time, a band whose stored solution is still within -max_solution_err is not solved again -- its stored result
  // goes into the sum as it is.  The decisions are taken up front: they only depend on each uid's own history.
  std::vector<char> need((size_t)std::max(0, b->nbands), 1);
  if (b->have_time) {
    if (!s->cfg.keep_solutions) return set_err(2, "solve_bands: a solution time needs keep_solutions = 1");
    for (int ib = 0; ib < b->nbands; ++ib) {
      int n = 1;
      if ((rc = pprts_b200_need_new_solution(s, b->uid0 + ib, b->time, 1, &n))) return rc;
      need[ib] = (char)n;
    }
  }
  auto next_needed = [&](int from) { while (from < b->nbands && !need[from]) ++from; return from; };
  int slot = 0, staged_band = next_needed(0);
  if (staged_band < b->nbands) {
    if ((rc = stage_band_async(s, s->staged[BAND_SLOT0], b, staged_band))) return rc;
    CU(cudaEventRecord(s->band_uploaded[0], s->copy_stream));
  }
  for (int ib = 0; ib < b->nbands; ++ib) {
    const int uid = b->uid0 + ib;
    if (!need[ib]) {  // reuse solver%solutions(uid): pprts_get_result of the stored solution
      if (niter_diff) niter_diff[ib] = -1;
      if ((rc = compute_result(s, uid, true, b->weights ? b->weights[ib] : 1.0))) return rc;
      continue;
    }
    Staged &st = s->staged[BAND_SLOT0 + slot];
    CU(cudaStreamWaitEvent(s->stream, s->band_uploaded[slot], 0));
    // returns after the ingest kernels have run: this slot and the other one (consumed one band earlier) are free
    rc = set_optprop_impl(s, st.albedo, st.kabs, st.ksca, st.g, st.have_planck ? st.planck : nullptr,
                          st.have_Bsrfc ? st.Bsrfc : nullptr, st.have_albedo2d ? st.albedo2d : nullptr,
                          b->ldelta_scaling, true);
    if (rc) return rc;
    staged_band = next_needed(ib + 1);
    slot ^= 1;
    if (staged_band < b->nbands) {  // upload of the next band to solve overlaps this band's solve
      if ((rc = stage_band_async(s, s->staged[BAND_SLOT0 + slot], b, staged_band))) return rc;
      CU(cudaEventRecord(s->band_uploaded[slot], s->copy_stream));
    }
    rc = solve_impl(s, !b->lsolar, b->lsolar, b->lsolar ? b->edirTOA[ib] : 0.0, uid, b->have_time != 0, b->time);
    if (rc) {
      char msg[sizeof(g_err)];
      snprintf(msg, sizeof(msg), "%s", g_err);
      cudaStreamSynchronize(s->copy_stream);
      return set_err(rc, "solve_bands: band %d (uid %d): %s", ib, uid, msg);
    }
    if (niter_diff) { Solution *sol; if (!find_solution(s, uid, &sol)) niter_diff[ib] = sol->niter_diff; }
    if ((rc = compute_result(s, uid, true, b->weights ? b->weights[ib] : 1.0))) return rc;
  }
  CU(cudaStreamSynchronize(s->copy_stream));
  if (edn || eup || abso || edir) return download_fields(s, s->acc, edn, eup, abso, edir);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// NCCL (dlopen'ed: single-GPU users do not need the library)
// ---------------------------------------------------------------------------------------------------------------
extern "C" int pprts_b200_nccl_unique_id(char id[128]) {
  int rc = load_nccl();
  if (rc) return rc;
  nccl_uid_t u;
  if ((rc = g_nccl.GetUniqueId(&u))) return set_err(41, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
  memcpy(id, u.internal, 128);
  return 0;
}
extern "C" int pprts_b200_comm_init(pprts_b200_solver *s, int nranks, int rank, const char id[128]) {
  if (!s) return set_err(1, "null solver");
  int rc = load_nccl();
  if (rc) return rc;
  CU(cudaSetDevice(s->device));
  nccl_uid_t u;
  memcpy(u.internal, id, 128);
  if (s->decomposed) {
    const int nxp = s->cfg.nxproc, nyp = s->cfg.nyproc, xi = s->cfg.xrank, yi = s->cfg.yrank;
    if (nranks != nxp * nyp) return set_err(45, "communicator has %d ranks, the process grid is %d x %d", nranks, nxp, nyp);
    if (rank != xi + yi * nxp)  // PETSc DMDA / MPI_Cart order, x fastest (src/pprts_base.F90:745-773)
      return set_err(45, "rank %d does not sit at (%d, %d) of a %d x %d grid (rank = xi + yi*nxproc)", rank, xi, yi, nxp, nyp);
    s->nb_west = (xi + nxp - 1) % nxp + yi * nxp;
    s->nb_east = (xi + 1) % nxp + yi * nxp;
    s->nb_south = xi + ((yi + nyp - 1) % nyp) * nxp;
    s->nb_north = xi + ((yi + 1) % nyp) * nxp;
  }
  if ((rc = g_nccl.CommInitRank(&s->nccl_comm, nranks, u, rank))) return set_err(42, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
  s->nranks = nranks; s->rank = rank;
  if (s->decomposed) {  // global column count for the solver tolerances; also checks the face lengths of the neighbours
    double h[3] = {(double)s->g.ncol, 0, 0};
    CU(cudaMemcpyAsync(s->sc + SC_TMP, h, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    NC(g_nccl.AllReduce(s->sc + SC_TMP, s->sc + SC_TMP, 1, NCCL_F64, NCCL_SUM, s->nccl_comm, s->stream));
    CU(cudaMemcpyAsync(h, s->sc + SC_TMP, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->glob_ncol = h[0];
    // determine_1d_layers (src/pprts.F90:708-724): with dz3d the ranks may find different numbers of 1-D layers; the
    // maximum over the ranks becomes 1-D on all of them, unconstrained_fraction is the global share of 3-D layers
    int n1d_loc = 0;
    for (unsigned char f : s->l1d_atm) n1d_loc += f;
    h[0] = (double)n1d_loc;
    CU(cudaMemcpyAsync(s->sc + SC_TMP, h, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    NC(g_nccl.AllReduce(s->sc + SC_TMP, s->sc + SC_TMP, 1, NCCL_F64, NCCL_MAX, s->nccl_comm, s->stream));
    CU(cudaMemcpyAsync(h, s->sc + SC_TMP, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if ((rc = finalize_1d_layers(s, (int)std::lround(h[0]), -1.0))) return rc;
    int c3 = 0;
    for (unsigned char f : s->l1d_atm) c3 += !f;
    h[0] = (double)c3; h[1] = (double)s->Nz_atm;
    CU(cudaMemcpyAsync(s->sc + SC_TMP, h, 2 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    NC(g_nccl.AllReduce(s->sc + SC_TMP, s->sc + SC_TMP, 2, NCCL_F64, NCCL_SUM, s->nccl_comm, s->stream));
    CU(cudaMemcpyAsync(h, s->sc + SC_TMP, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->unconstrained_fraction = h[0] / h[1];
    // halo_fill_5pt of atm%hhl (src/pprts.F90:603): the level heights of the neighbours' edge columns, for the centred
    // differences of atm%hgrad (:621-638) at the rank boundaries (slope_correction_fluxes); static, exchanged once
    if (s->have_dz3d && s->vh) {
      const Grid &g = s->g;
      const size_t nx = (size_t)(g.Nz + 1) * g.Ny, ny = (size_t)(g.Nz + 1) * g.Nx;
      double *send[4] = {nullptr, nullptr, nullptr, nullptr};
      for (int q = 0; q < 4; ++q) {
        const size_t n = q < 2 ? nx : ny;
        if (!s->vh_ghost[q]) DA(s->vh_ghost[q], n);
        CU(cudaMalloc((void **)&send[q], n * sizeof(double)));
        LAUNCH(s, FAM_HALO, k_pack_edge, nblk((long)n, 256), 256, g, q, (const double *)s->vh, send[q]);
      }
      // my first column is the EAST ghost (i = Nx) of my west neighbour, my last column the WEST ghost of my east neighbour
      NC(g_nccl.GroupStart());
      NC(g_nccl.Send(send[0], nx, NCCL_F64, s->nb_west, s->nccl_comm, s->stream));
      NC(g_nccl.Recv(s->vh_ghost[1], nx, NCCL_F64, s->nb_east, s->nccl_comm, s->stream));
      NC(g_nccl.Send(send[1], nx, NCCL_F64, s->nb_east, s->nccl_comm, s->stream));
      NC(g_nccl.Recv(s->vh_ghost[0], nx, NCCL_F64, s->nb_west, s->nccl_comm, s->stream));
      NC(g_nccl.Send(send[2], ny, NCCL_F64, s->nb_south, s->nccl_comm, s->stream));
      NC(g_nccl.Recv(s->vh_ghost[3], ny, NCCL_F64, s->nb_north, s->nccl_comm, s->stream));
      NC(g_nccl.Send(send[3], ny, NCCL_F64, s->nb_north, s->nccl_comm, s->stream));
      NC(g_nccl.Recv(s->vh_ghost[2], ny, NCCL_F64, s->nb_south, s->nccl_comm, s->stream));
      NC(g_nccl.GroupEnd());
      CU(cudaStreamSynchronize(s->stream));
      for (double *p : send) cudaFree(p);
    }
  }
  return 0;
}
extern "C" int pprts_b200_allreduce_accumulated(pprts_b200_solver *s) {
  if (!s) return set_err(1, "null solver");
  if (!s->nccl_comm) return s->nranks == 1 ? 0 : set_err(43, "communicator not initialised");
  if (s->decomposed) return set_err(46, "band sharding and x/y decomposition share one communicator: use one of them");
  CU(cudaSetDevice(s->device));
  for (int q = 0; q < 4; ++q) {
    int rc = g_nccl.AllReduce(s->acc[q], s->acc[q], (size_t)s->g.nlev, /*ncclFloat64*/ 8, /*ncclSum*/ 0, s->nccl_comm, s->stream);
    if (rc) return set_err(44, "ncclAllReduce: %s", g_nccl.GetErrorString(rc));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------------------------
extern "C" int pprts_b200_get_coeffs(pprts_b200_solver *s, int which, double *out) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const Grid &g = s->g;
  const int nc = which == 1 ? s->Nd * s->Nd : which == 2 ? s->Nd * s->Nf : s->Nf * s->Nf;
  const void *src = which == 1 ? s->d2d : which == 2 ? s->d2f : s->f2f;
  double *tmp;
  CU(cudaMalloc((void **)&tmp, (size_t)nc * g.ncell * sizeof(double)));
  if (s->cfg.coeff_fp32) LAUNCH(s, FAM_IO, k_coeff_to_host_layout<float>, nblk((long)nc * g.ncell, 256), 256, g, nc, (const float *)src, tmp);
  else LAUNCH(s, FAM_IO, k_coeff_to_host_layout<double>, nblk((long)nc * g.ncell, 256), 256, g, nc, (const double *)src, tmp);
  CU(cudaMemcpyAsync(out, tmp, (size_t)nc * g.ncell * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaFree(tmp));
  return 0;
}
extern "C" int pprts_b200_get_state(pprts_b200_solver *s, int uid, int which, double *out) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const Grid &g = s->g;
  Solution *sol = nullptr;
  if (which != 4) { int rc = find_solution(s, uid, &sol); if (rc) return rc; }
  if (which == 3) {
    dim3 grid((g.Nx + 31) / 32, (g.Nz + 31) / 32, g.Ny), block(32, 8);
    LAUNCH(s, FAM_IO, k_kji_to_zxy<double>, grid, block, g.Nz, g.Nx, g.Ny, sol->abso, s->out_stage, 1.0);
    CU(cudaMemcpyAsync(out, s->out_stage, (size_t)g.ncell * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  } else {
    const int nd = which == 1 ? s->Nd : s->Nf;
    const double *src = which == 1 ? sol->edir : which == 2 ? sol->ediff : s->b;
    LAUNCH(s, FAM_IO, k_state_to_host_layout, nblk((long)nd * g.nlev, 256), 256, g, nd, src, s->out_stage);
    CU(cudaMemcpyAsync(out, s->out_stage, (size_t)nd * g.nlev * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int pprts_b200_get_info(pprts_b200_solver *s, int uid, int *niter_dir, int *niter_diff, double *res_dir, double *res_diff) {
  if (!s) return set_err(1, "null solver");
  auto it = s->solutions.find(s->cfg.keep_solutions ? uid : 0);
  if (it == s->solutions.end()) return set_err(30, "no solution for uid %d", uid);
  *niter_dir = it->second.niter_dir; *niter_diff = it->second.niter_diff;
  *res_dir = it->second.res_dir; *res_diff = it->second.res_diff;
  return 0;
}
// opt_buildings of solve_pprts (src/pprts.F90:2487-2494): the faces stay attached to the solver until replaced or
// removed (nfaces = 0).  Coefficients are re-derived on the next solve, as alloc_coeff_* does on every call (:2682-2687).
extern "C" int pprts_b200_set_buildings(pprts_b200_solver *s, int nfaces, const int *iface, const double *albedo,
                                        const double *planck) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  if (nfaces > 0 && (!iface || !albedo)) return set_err(2, "set_buildings: iface and albedo are required");
  const Grid &g = s->g;
  for (int m = 0; m < nfaces; ++m)
    if (iface[m] < 1 || (long)iface[m] > 6L * g.ncell)
      return set_err(61, "building face %d: index %d outside the (6,%d,%d,%d) grid", m + 1, iface[m], g.Nz, g.Nx, g.Ny);
  CU(cudaStreamSynchronize(s->stream));
  cudaFree(s->bld_iface); cudaFree(s->bld_albedo); cudaFree(s->bld_planck); cudaFree(s->bld_out);
  s->bld_iface = nullptr; s->bld_albedo = s->bld_planck = s->bld_out = nullptr;
  if (s->nbld > 0 || nfaces > 0) s->coeffs_dirty = true;
  s->nbld = std::max(0, nfaces);
  s->bld_have_planck = false;
  if (nfaces <= 0) return 0;
  DA(s->bld_iface, (size_t)nfaces); DA(s->bld_albedo, (size_t)nfaces); DA(s->bld_planck, (size_t)nfaces);
  DA(s->bld_out, (size_t)3 * nfaces);
  CU(cudaMemcpy(s->bld_iface, iface, sizeof(int) * nfaces, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(s->bld_albedo, albedo, sizeof(double) * nfaces, cudaMemcpyHostToDevice));
  if (planck) {
    CU(cudaMemcpy(s->bld_planck, planck, sizeof(double) * nfaces, cudaMemcpyHostToDevice));
    s->bld_have_planck = true;
  }
  return 0;
}
// pprts_get_result(..., opt_buildings): fill_buildings_arr (src/pprts.F90:6011-6160); edir may be NULL
extern "C" int pprts_b200_get_buildings_result(pprts_b200_solver *s, int uid, double *edir, double *incoming,
                                               double *outgoing) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  if (s->nbld <= 0 && !s->decomposed) return set_err(62, "no buildings attached to the solver");
  Solution *sol;
  int rc = find_solution(s, uid, &sol);
  if (rc) return rc;
  const Grid &g = s->g;
  if (s->decomposed) {  // collective: the right / front faces on the rank edge read the neighbour's low faces (in W/m2 by now)
    if ((rc = need_comm(s))) return rc;
    if (sol->lsolar_rad && (rc = halo_fill(s, sol->edir, s->Nd, dir_halo(s), true, true))) return rc;
    if ((rc = halo_fill(s, sol->ediff, s->Nf, diff_halo(s), true, true))) return rc;
    if (s->nbld <= 0) return 0;  // a rank without faces only takes part in the exchange
  }
  const int n = s->nbld;
  DISPATCH_L(s, LAUNCH(s, FAM_ABSO, (k_buildings_result<L>), nblk(n, 128), 128, g, (int)sol->lsolar_rad, s->mu, n,
                        (const int *)s->bld_iface, (const double *)sol->edir, (const double *)sol->ediff, s->bld_out,
                        s->bld_out + n, s->bld_out + 2 * n));
  std::vector<double> h((size_t)3 * n);
  CU(cudaMemcpyAsync(h.data(), s->bld_out, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (edir) std::copy(h.begin(), h.begin() + n, edir);
  if (incoming) std::copy(h.begin() + n, h.begin() + 2 * n, incoming);
  if (outgoing) std::copy(h.begin() + 2 * n, h.end(), outgoing);
  return 0;
}
// set_abso_in_buildings (src/pprts.F90:5986-6009), what pprts_get_result(opt_buildings) does LAST under
// -pprts_set_abso_in_buildings <val>: solution%abso of the cells with building faces is overwritten after the result has
// been copied out, so the pprts_get_result calls that follow see val (x mu for a solar solution), the current one does not.
extern "C" int pprts_b200_set_abso_in_buildings(pprts_b200_solver *s, int uid, double val) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  if (s->nbld <= 0) return s->decomposed ? 0 : set_err(62, "no buildings attached to the solver");
  Solution *sol;
  int rc = find_solution(s, uid, &sol);
  if (rc) return rc;
  LAUNCH(s, FAM_ABSO, k_buildings_set_abso, nblk(s->nbld, 128), 128, s->g, s->nbld, (const int *)s->bld_iface, val, sol->abso);
  s->res_valid = false;
  return 0;
}
// slope_correction_fluxes(solver, edir) (src/pprts_postprocess.F90:130-196), the spectral callers' post-processing of
// the direct flux on a terrain-following grid (rrtmg/rrtmg/pprts_rrtmg.F90 after the band sum).  edir: host array
// (Nz+1,Nx,Ny) column-major, corrected in place.  Needs init with dz3d (atm%hgrad, src/pprts.F90:621) and one domain.
static void sun_vector(const pprts_b200_solver *s, double v[3]) {  // sun%sundir = spherical_2_cartesian(phi, theta)
  const double d2r = PB_PI / 180.0, th = s->theta < 0 ? 90.0 : s->theta;
  v[0] = -std::sin(th * d2r) * std::sin(s->phi * d2r);
  v[1] = -std::sin(th * d2r) * std::cos(s->phi * d2r);
  v[2] = -std::cos(th * d2r);
}
static int slope_correct_device(pprts_b200_solver *s, int lslope, int latm, double *field) {
  if (!lslope && !latm) return 0;
  if (!s->have_dz3d || !s->vh) return set_err(70, "slope correction needs a solver initialised with dz3d (atm%%hgrad not initialized!)");
  if (s->decomposed && !s->vh_ghost[0]) return set_err(71, "slope correction on a decomposed solver: call pprts_b200_comm_init first (it exchanges the level heights)");
  double sv[3];
  sun_vector(s, sv);
  const Grid &g = s->g;
  HeightGhosts gh{{s->vh_ghost[0], s->vh_ghost[1], s->vh_ghost[2], s->vh_ghost[3]}};
  LAUNCH(s, FAM_ABSO, k_slope_correction, nblk(g.nlev, 256), 256, g, lslope, latm, sv[0], sv[1], sv[2], (const double *)s->vh, gh, field);
  return 0;
}
extern "C" int pprts_b200_slope_correction_fluxes(pprts_b200_solver *s, int lslope, int latm, double *edir) {
  if (!s || !edir) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const Grid &g = s->g;
  const size_t n = (size_t)g.nlev;
  double *dev = s->out_stage, *tmp = s->kr[7];
  CU(cudaMemcpyAsync(tmp, edir, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  {
    dim3 block(32, 8), grid((unsigned int)((g.Nx + 31) / 32), (unsigned int)((g.Nz + 1 + 31) / 32), (unsigned int)g.Ny);
    LAUNCH(s, FAM_IO, k_zxy_to_kji<double>, grid, block, g.Nz + 1, g.Nx, g.Ny, (const double *)tmp, dev);
  }
  int rc = slope_correct_device(s, lslope, latm, dev);
  if (rc) return rc;
  {
    dim3 block(32, 8), grid((unsigned int)((g.Nx + 31) / 32), (unsigned int)((g.Nz + 1 + 31) / 32), (unsigned int)g.Ny);
    LAUNCH(s, FAM_IO, k_kji_to_zxy<double>, grid, block, g.Nz + 1, g.Nx, g.Ny, (const double *)dev, tmp, 1.0);
  }
  CU(cudaMemcpyAsync(edir, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
// the same on the device-side spectral sum of the direct flux (pprts_b200_accumulate_result / solve_bands)
extern "C" int pprts_b200_slope_correct_accumulated(pprts_b200_solver *s, int lslope, int latm) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  return slope_correct_device(s, lslope, latm, s->acc[2]);  // accumulator order: edn, eup, edir, abso
}
// smooth_surface_fluxes(solver, edn, eup) (src/pprts_postprocess.F90:44-84), the other post-processing step of the
// spectral callers (rrtmg/rrtmg/pprts_rrtmg.F90:319); the option -pprts_smooth_srfc_flx <radius> arrives as an argument.
// find_iter_and_kernelwidth (src/pprts_postprocess.F90:86-128) is pure host arithmetic.
extern "C" int pprts_b200_find_iter_and_kernelwidth(double radius, double dx, double dy, int xm, int ym, int *Niter,
                                                    int *kernel_width) {
  if (!Niter || !kernel_width) return set_err(1, "null output");
  const int test_Ni = 500, Ni_limit = 250;
  std::vector<double> test_k(test_Ni);
  const double rpix = radius / ((dx + dy) / 2);
  int min_iter = 1;
  for (int i = 1; i <= test_Ni; ++i) {
    test_k[i - 1] = (std::sqrt((12 * rpix * rpix + (double)i) / (double)i + 1) - 1) / 2;
    if (std::lround(test_k[i - 1]) >= std::min(xm, ym)) min_iter = i + 1;
  }
  if (min_iter > Ni_limit)
    return set_err(72, "the smoothing iteration count would be larger than %d... this could become expensive (%d)", Ni_limit, min_iter);
  int arg = min_iter;
  double best = std::fabs(test_k[arg - 1] - (double)std::lround(test_k[arg - 1]));
  for (int i = min_iter + 1; i <= test_Ni; ++i) {
    const double res = std::fabs(test_k[i - 1] - (double)std::lround(test_k[i - 1]));
    if (res < best) { best = res; arg = i; }
  }
  *Niter = arg;
  *kernel_width = (int)std::lround(test_k[arg - 1]);
  if (*kernel_width == 0) *Niter = 0;
  return 0;
}
// field: one level [Ny][Nx] on the device, smoothed in place; scratch: at least ncol doubles
// One pass of the box blur on a decomposed solver: the kw edge columns (rows) travel to the two neighbours along that
// direction (convolve_ediff_srfc's exchange, src/pprts_base.F90:1766-1795); a direction that is not split wraps locally.
static int blur_pass_decomposed(pprts_b200_solver *s, int kw, int along_y, const double *in, double *out) {
  const Grid &g = s->g;
  const size_t n = (size_t)kw * (along_y ? g.Nx : g.Ny);
  if (n > s->blur_cap) {
    for (double *&p : s->blur_buf) { cudaFree(p); p = nullptr; }
    s->blur_cap = 0;  // nothing usable until all four buffers exist again
    for (double *&p : s->blur_buf) DA(p, n);
    s->blur_cap = n;
  }
  double *snd_lo = s->blur_buf[0], *snd_hi = s->blur_buf[1], *halo_lo = s->blur_buf[2], *halo_hi = s->blur_buf[3];
  LAUNCH(s, FAM_HALO, k_blur_pack, nblk((long)n, 256), 256, g.Nx, g.Ny, kw, along_y, in, snd_lo, snd_hi);
  if (along_y ? g.gy : g.gx) {
    const int lo_nb = along_y ? s->nb_south : s->nb_west, hi_nb = along_y ? s->nb_north : s->nb_east;
    // my lowest columns are the halo ABOVE my low neighbour's range, my highest ones the halo BELOW my high neighbour's
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(snd_lo, n, NCCL_F64, lo_nb, s->nccl_comm, s->stream));
    NC(g_nccl.Recv(halo_hi, n, NCCL_F64, hi_nb, s->nccl_comm, s->stream));
    NC(g_nccl.Send(snd_hi, n, NCCL_F64, hi_nb, s->nccl_comm, s->stream));
    NC(g_nccl.Recv(halo_lo, n, NCCL_F64, lo_nb, s->nccl_comm, s->stream));
    NC(g_nccl.GroupEnd());
    s->halo_exchanges++;
  } else {  // periodic within the rank: my own opposite edge
    halo_lo = snd_hi; halo_hi = snd_lo;
  }
  LAUNCH(s, FAM_ABSO, k_box_blur_halo, nblk(g.ncol, 256), 256, g.Nx, g.Ny, kw, along_y, in, (const double *)halo_lo,
         (const double *)halo_hi, out);
  return 0;
}
static int smooth_level_device(pprts_b200_solver *s, double radius, double *field, double *scratch) {
  const Grid &g = s->g;
  int rc;
  if (s->decomposed) {  // collective over the ranks of the decomposition
    if ((rc = need_comm(s))) return rc;
    if (radius < 0) {   // imp_allreduce_mean: sum over all ranks / global number of columns
      LAUNCH(s, FAM_ABSO, k_field_sum, 1, 1024, (long)g.ncol, (const double *)field, s->sc + SC_TMP);
      if ((rc = allreduce_scalars(s, s->sc + SC_TMP, 1))) return rc;
      LAUNCH(s, FAM_ABSO, k_field_fill, nblk(g.ncol, 256), 256, (long)g.ncol, field, (const double *)(s->sc + SC_TMP), 1.0 / s->glob_ncol);
      return 0;
    }
    // find_iter_and_kernelwidth works on the LOCAL domain size; Niter = max, kernel_width = min over the ranks
    // (src/pprts_postprocess.F90:116-119).  A rank whose own search fails still takes part in the reduction.
    int niter = 0, kw = 0;
    char msg[sizeof(g_err)];
    const int lrc = pprts_b200_find_iter_and_kernelwidth(radius, g.dx, g.dy, g.Nx, g.Ny, &niter, &kw);
    if (lrc) snprintf(msg, sizeof(msg), "%s", g_err);
    double h[3] = {(double)niter, -(double)kw, (double)lrc};
    CU(cudaMemcpyAsync(s->sc + SC_TMP, h, sizeof(h), cudaMemcpyHostToDevice, s->stream));
    NC(g_nccl.AllReduce(s->sc + SC_TMP, s->sc + SC_TMP, 3, NCCL_F64, NCCL_MAX, s->nccl_comm, s->stream));
    CU(cudaMemcpyAsync(h, s->sc + SC_TMP, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (h[2] != 0.0) return lrc ? set_err(lrc, "%s", msg) : set_err(72, "smooth_surface_fluxes: find_iter_and_kernelwidth failed on another rank");
    niter = (int)std::lround(h[0]); kw = (int)std::lround(-h[1]);
    if (kw == 0) niter = 0;
    for (int it = 0; it < niter; ++it) {
      if ((rc = blur_pass_decomposed(s, kw, 0, field, scratch))) return rc;
      if ((rc = blur_pass_decomposed(s, kw, 1, scratch, field))) return rc;
    }
    return 0;
  }
  if (radius < 0) {
    LAUNCH(s, FAM_ABSO, k_fill_mean, 1, 1024, (long)g.ncol, field);
    return 0;
  }
  int niter = 0, kw = 0;
  rc = pprts_b200_find_iter_and_kernelwidth(radius, g.dx, g.dy, g.Nx, g.Ny, &niter, &kw);
  if (rc) return rc;
  for (int it = 0; it < niter; ++it) {
    LAUNCH(s, FAM_ABSO, k_box_blur, nblk(g.ncol, 256), 256, g.Nx, g.Ny, kw, 0, (const double *)field, scratch);
    LAUNCH(s, FAM_ABSO, k_box_blur, nblk(g.ncol, 256), 256, g.Nx, g.Ny, kw, 1, (const double *)scratch, field);
  }
  return 0;
}
// edn, eup: host arrays (Nz+1,Nx,Ny) column-major; only the surface level changes
extern "C" int pprts_b200_smooth_surface_fluxes(pprts_b200_solver *s, double radius, double *edn, double *eup) {
  if (!s || !edn || !eup) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const Grid &g = s->g;
  const size_t nc = (size_t)g.ncol, nz1 = (size_t)g.Nz + 1;
  std::vector<double> h(2 * nc);
  for (size_t c = 0; c < nc; ++c) { h[c] = edn[g.Nz + nz1 * c]; h[nc + c] = eup[g.Nz + nz1 * c]; }
  double *dev = s->out_stage;
  CU(cudaMemcpyAsync(dev, h.data(), 2 * nc * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  int rc;
  if ((rc = smooth_level_device(s, radius, dev, s->kr[7])) || (rc = smooth_level_device(s, radius, dev + nc, s->kr[7]))) return rc;
  CU(cudaMemcpyAsync(h.data(), dev, 2 * nc * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  for (size_t c = 0; c < nc; ++c) { edn[g.Nz + nz1 * c] = h[c]; eup[g.Nz + nz1 * c] = h[nc + c]; }
  return 0;
}
// the same on the device-side spectral sums of edn and eup (pprts_b200_accumulate_result / solve_bands)
extern "C" int pprts_b200_smooth_accumulated(pprts_b200_solver *s, double radius) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const size_t srfc = (size_t)s->g.Nz * s->g.ncol;
  int rc = smooth_level_device(s, radius, s->acc[0] + srfc, s->kr[7]);
  return rc ? rc : smooth_level_device(s, radius, s->acc[1] + srfc, s->kr[7]);
}
extern "C" int pprts_b200_mat_mult_ediff(pprts_b200_solver *s, const double *x, double *y) {
  if (!s) return set_err(1, "null solver");
  CU(cudaSetDevice(s->device));
  const Grid &g = s->g;
  int rc = ensure_coeffs(s, false);
  if (rc) return rc;
  const long n = (long)s->Nf * g.nlev;
  CU(cudaMemcpyAsync(s->out_stage, x, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  LAUNCH(s, FAM_IO, k_state_from_host_layout, nblk(n, 256), 256, g, s->Nf, s->out_stage, s->kr[6]);
  if ((rc = need_comm(s)) || (rc = launch_matmult<0, double>(s, s->kr[6], s->kr[7], nullptr, nullptr))) return rc;
  LAUNCH(s, FAM_IO, k_state_to_host_layout, nblk(n, 256), 256, g, s->Nf, s->kr[7], s->out_stage);
  CU(cudaMemcpyAsync(y, s->out_stage, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

static int lookup_probe(pprts_b200_solver *s, int mode, long n, const double *kabs, const double *ksca, const double *g,
                        const double *dz, const float *lutspace, const SymMaps &sym, double *out) {
  CU(cudaSetDevice(s->device));
  if (mode != 3 && s->angles_dirty) { int rc = contract_direct_tables(s); if (rc) return rc; }
  const int nc = mode == 1 ? s->Nd * s->Nd : mode == 2 ? s->Nd * s->Nf : s->Nf * s->Nf;
  double *d_in[4] = {nullptr, nullptr, nullptr, nullptr}, *d_out;
  float *d_ls = nullptr;
  const double *h_in[4] = {kabs, ksca, g, dz};
  if (lutspace) { CU(cudaMalloc((void **)&d_ls, n * 4 * sizeof(float))); CU(cudaMemcpy(d_ls, lutspace, n * 4 * sizeof(float), cudaMemcpyHostToDevice)); }
  else for (int q = 0; q < 4; ++q) { CU(cudaMalloc((void **)&d_in[q], n * sizeof(double))); CU(cudaMemcpy(d_in[q], h_in[q], n * sizeof(double), cudaMemcpyHostToDevice)); }
  CU(cudaMalloc((void **)&d_out, (size_t)n * nc * sizeof(double)));
  DISPATCH_L(s, LAUNCH(s, FAM_LOOKUP, (k_lookup_probe<L>), nblk(n, 128), 128, s->lutdev, sym, mode, n, s->g.dx, d_in[0],
                        d_in[1], d_in[2], d_in[3], d_ls, d_out));
  CU(cudaMemcpyAsync(out, d_out, (size_t)n * nc * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  for (int q = 0; q < 4; ++q) cudaFree(d_in[q]);
  cudaFree(d_ls); cudaFree(d_out);
  return 0;
}
extern "C" int pprts_b200_lookup_coeffs(pprts_b200_solver *s, int mode, long n, const double *kabs, const double *ksca,
                                        const double *g, const double *dz, double *out) {
  if (!s) return set_err(1, "null solver");
  if (mode < 1 || mode > 3) return set_err(1, "mode option %d not recognized", mode);
  return lookup_probe(s, mode, n, kabs, ksca, g, dz, nullptr, s->sym, out);
}
// used by the pprts_f2c_opp_* compatibility layer (f2c.cu)
int pprts_b200_internal_opp_probe(pprts_b200_solver *s, int mode, const float sample[4], int le, int ln, double *out) {
  SymMaps sym;
  build_sym_maps(s->cfg.solver_id, le != 0, ln != 0, sym);
  return lookup_probe(s, mode, 1, nullptr, nullptr, nullptr, nullptr, sample, sym, out);
}
extern "C" int pprts_b200_debug_fail_next_krylov(pprts_b200_solver *s, int n) { if (!s) return set_err(1, "null solver"); s->fail_next_krylov = n; return 0; }
extern "C" long pprts_b200_krylov_retries(pprts_b200_solver *s) { return s ? s->krylov_retries : 0; }
int pprts_b200_internal_force_angles(pprts_b200_solver *s, double symmetry_phi, double theta) {
  s->symmetry_phi = symmetry_phi; s->theta = theta; s->angles_dirty = true; s->coeffs_dirty = true;
  return 0;
}

extern "C" long pprts_b200_kernel_launches(pprts_b200_solver *s) { return s ? s->launches : 0; }
extern "C" void *pprts_b200_stream(pprts_b200_solver *s) { return s ? (void *)s->stream : nullptr; }
extern "C" int pprts_b200_enable_kernel_timing(pprts_b200_solver *s, int on) {
  if (!s) return set_err(1, "null solver");
  collect_events(s);
  s->timing = on != 0;
  return 0;
}
extern "C" int pprts_b200_kernel_stats(pprts_b200_solver *s, int reset, double ms[16], long launches[16]) {
  if (!s) return set_err(1, "null solver");
  collect_events(s);
  for (int q = 0; q < 16; ++q) { ms[q] = 0; launches[q] = 0; }
  for (int q = 0; q < FAM_COUNT; ++q) { ms[q] = s->fam_ms[q]; launches[q] = s->fam_launch[q]; }
  if (reset) for (int q = 0; q < FAM_COUNT; ++q) { s->fam_ms[q] = 0; s->fam_launch[q] = 0; }
  return 0;
}
int pprts_b200_internal_lut_ranges(const pprts_b200_lut *lut, float ranges[6][2], int *Nd, int *Nf) {
  for (int d = 0; d < 6; ++d) { ranges[d][0] = lut->ax[d].front(); ranges[d][1] = lut->ax[d].back(); }
  *Nd = lut->Nd; *Nf = lut->Nf;
  return 0;
}
