// common.cuh -- grid descriptors, stream layouts and the deterministic block reduction shared by all kernels.
//
// Data layout in HBM (DESIGN.md "Data layout"): everything is structure-of-arrays with x fastest so that a warp
// of 32 neighbouring cells/columns reads 128/256 contiguous bytes from every plane:
//   cell  field  f[(k*Ny + j)*Nx + i]                      k = 0..Nz-1 (TOA first)
//   level field  f[(k*Ny + j)*Nx + i]                      k = 0..Nz
//   state        x[d*nlev + (k*Ny + j)*Nx + i]             d = stream (dof), one plane per dof
//   coefficients c[(cell/128)*NP*128 + n*128 + cell%128]   n = dst*Nsrc + src (reference order), NP = planes of the array:
//                blocked by tiles of 128 consecutive cells, so the NP coefficients of a tile are ONE contiguous run
//                (51 KB for diff2diff 3_10) that a single bulk copy fetches with perfect DRAM page locality
//   divergence   cdiv[s*ncp + cell]                        fp64 planes, ncp = ncell padded to the tile
// A side dof with index (k,i,j) lives on the low-x / low-y face of cell (k,i,j), exactly as in the reference
// (src/pprts_explicit.F90:398-451); face i+1 of the last column is the periodic image (one GPU) or a ghost.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PB_MAX_DOF 24
#define PB_MAX_DIRDOF 8
#define PB_TILE 128        // cells per coefficient tile
#define PB_TILE_SHIFT 7

struct Grid {
  int Nz, Nx, Ny;
  long ncol;   // Nx*Ny
  long ncell;  // Nz*ncol
  long nlev;   // (Nz+1)*ncol
  long ncp;    // ncell padded to a multiple of PB_TILE
  double dx, dy;
  int gx, gy;  // 1: the domain is split across GPUs in x / y -> the high-x / high-y faces are ghost planes
  int kb1d;    // first layer of the 1-D block at the BOTTOM of the column (Nz if there is none), see is1d()
};
// determine_1d_layers (src/pprts.F90:649-725) yields a 1-D block of n1d layers contiguous from the top and, when only
// the lowest layer is thicker than twostr_ratio * dx, that layer on its own: l1d = [T..T, F..F, T].  Kernels get the top
// block as `n1d` and the bottom one through the grid.
__host__ __device__ __forceinline__ bool is1d(const Grid &g, int n1d, int k) { return k < n1d || k >= g.kb1d; }

// x/y decomposition (src/pprts_base.F90:721-828, halo exchange src/pprts_explicit.F90:232-328, :715-848).
// A state vector of ND dofs is [ND*nlev main][ND*(Nz+1)*Ny ghost-x][ND*(Nz+1)*Nx ghost-y]: the ghost-x block holds the
// dofs living on the face i = Nx (owned by the east neighbour as its i = 0 face), laid out [d][k][j]; ghost-y likewise
// for j = Ny, laid out [d][k][i].  Kernels reach the high-x / high-y face of a cell through (offset, dof stride)
// pairs, so the periodic single-GPU case, the ghost case and interior cells run the same branch-free code.
__host__ __device__ __forceinline__ long state_len(const Grid &g, int nd) {
  return (long)nd * g.nlev + (long)nd * (g.Nz + 1) * ((g.gx ? g.Ny : 0) + (g.gy ? g.Nx : 0));
}
__host__ __device__ __forceinline__ long ghost_x_base(const Grid &g, int nd) { return (long)nd * g.nlev; }
__host__ __device__ __forceinline__ long ghost_y_base(const Grid &g, int nd) {
  return (long)nd * g.nlev + (g.gx ? (long)nd * (g.Nz + 1) * g.Ny : 0);
}

// Stream layouts (src/pprts.F90:332-396). Top streams alternate [F,T,...] (E_up first); side streams alternate
// starting with F (3_10, 8_10) or T (3_24).
template <int NDT, int NDS, int NFT, int NFS, bool SIDE0_IN, int ADIV_TOP, int ADIV_SIDE>
struct Layout {
  static constexpr int Ndt = NDT, Nds = NDS, Nft = NFT, Nfs = NFS;
  static constexpr int Nd = NDT + 2 * NDS, Nf = NFT + 2 * NFS;
  static constexpr int dirtop_adiv = ADIV_TOP, dirside_adiv = ADIV_SIDE;
  __host__ __device__ static constexpr bool top_in(int d) { return (d & 1) != 0; }
  __host__ __device__ static constexpr bool side_in(int d) { return SIDE0_IN ? ((d & 1) == 0) : ((d & 1) != 0); }
  // inv_dof (src/pprts_explicit.F90:1001-1014) for top_in(0) == false
  __host__ __device__ static constexpr int inv_dof(int d) { return top_in(d) ? d - 1 : d + 1; }
};
using L310 = Layout<1, 1, 2, 4, false, 1, 1>;
using L810 = Layout<4, 2, 2, 4, false, 4, 2>;
using L324 = Layout<1, 1, 8, 8, true, 1, 1>;

// tile-blocked coefficient index: base of cell c in an array of NP planes; coefficient n lives at base + n*PB_TILE
template <int NP>
__device__ __forceinline__ long cbase(long c) { return (c >> PB_TILE_SHIFT) * (long)(NP * PB_TILE) + (c & (PB_TILE - 1)); }
__device__ __forceinline__ long cbase_rt(int np, long c) { return (c >> PB_TILE_SHIFT) * ((long)np * PB_TILE) + (c & (PB_TILE - 1)); }

// fp32 -> fp64 widening on the integer ALU.  cvt.f64.f32 (SASS F2F.F64.F32) runs on the XU pipe at ~4 results/clk/SM
// on sm_100a; with Nf*Nf conversions per cell it saturated that pipe in the diffuse MatMult (ncu: XU 87 % busy,
// profiles/r1_matmult_tma.txt).  For a normal, non-negative float the widening is a re-bias of the exponent and a
// 29-bit shift of the mantissa -- two ALU instructions (SHF + LEA.HI); everything else (zero, subnormal, negative,
// inf/nan) takes the exact but slow cvt path.  Bit-identical to (double)f for every input.
__device__ __forceinline__ double widen(float f) {
  const unsigned int b = __float_as_uint(f);
  if (__builtin_expect((b - 0x00800000u) < 0x7f000000u, 1))
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
  return (double)f;
}
__device__ __forceinline__ double widen(double f) { return f; }
// Per-tile fast paths of the diffuse MatMult.  The lookup kernel classifies every coefficient tile (k_lookup,
// `tile_class`): 0 = every entry is a normal positive finite float, 1 = additionally exact +0 entries, 2 = anything
// else (subnormal, negative, inf/nan).  Class 0 widens with ONE integer multiply-add, b * 2^29 + (896 << 52): the
// 64-bit product puts mantissa and exponent in place, the addend re-biases the exponent (SASS IMAD.WIDE.U32).
// Class 1 adds the bias only where b != 0 (ISETP + predicated IADD).  Class 2 uses cvt.  All three are bit-identical
// to (double)f on the values of their class.
__device__ __forceinline__ int widen_class(float f) {
  const unsigned int b = __float_as_uint(f);
  if ((b - 0x00800000u) < 0x7f000000u) return 0;
  return b == 0u ? 1 : 2;
}
__device__ __forceinline__ double widen_normal(float f) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 0x20000000, %2;" : "=l"(r) : "r"(__float_as_uint(f)), "l"(0x3800000000000000ull));
  return __longlong_as_double((long long)r);
}
__device__ __forceinline__ double widen_normal_or_zero(float f) {
  const unsigned int b = __float_as_uint(f);
  unsigned long long r;
  asm("mul.wide.u32 %0, %1, 0x20000000;" : "=l"(r) : "r"(b));
  unsigned int hi = (unsigned int)(r >> 32);
  if (b) hi += 0x38000000u;
  return __hiloint2double((int)hi, (int)(unsigned int)r);
}

// Signed variant for the fp32 Krylov work vectors (values of either sign, exact zeros are common): normal numbers and
// +-0 stay on the integer ALU, subnormals / inf / nan take cvt.  Bit-identical to (double)f.
__device__ __forceinline__ double widen_any(float f) {
  const unsigned int b = __float_as_uint(f), m = b & 0x7fffffffu;
  if (__builtin_expect((m - 0x00800000u) < 0x7f000000u || m == 0u, 1))
    return __hiloint2double((int)(((m >> 3) + (m ? 0x38000000u : 0u)) | (b & 0x80000000u)), (int)(b << 29));
  return (double)f;
}
__device__ __forceinline__ double widen_any(double f) { return f; }

__device__ __forceinline__ long lidx(const Grid &g, int k, int j, int i) { return ((long)k * g.Ny + j) * g.Nx + i; }
__device__ __forceinline__ int wrap_p(int i, int n) { return (i + 1 == n) ? 0 : i + 1; }

// element offsets of the face locations of cell (k, j, i) in a state vector of ND dofs: dof d on the cell's own
// (low) faces is at own + d*nlev, on the level below at dn + d*nlev, on the high-x face at xp + d*xps, on the high-y
// face at yp + d*yps
struct Faces {
  long own, dn, xp, xps, yp, yps;
  bool xg, yg;  // the high-x / high-y face is a ghost (its rows belong to the neighbouring rank)
  int i, j, k;
};
template <int ND>
__device__ __forceinline__ Faces cell_faces(const Grid &g, long c) {
  Faces f;
  f.i = (int)(c % g.Nx); f.j = (int)((c / g.Nx) % g.Ny); f.k = (int)(c / g.ncol);
  f.own = c; f.dn = c + g.ncol;
  f.xg = g.gx && f.i + 1 == g.Nx;
  f.yg = g.gy && f.j + 1 == g.Ny;
  if (f.xg) { f.xp = ghost_x_base(g, ND) + (long)f.k * g.Ny + f.j; f.xps = (long)(g.Nz + 1) * g.Ny; }
  else { f.xp = lidx(g, f.k, f.j, wrap_p(f.i, g.Nx)); f.xps = g.nlev; }
  if (f.yg) { f.yp = ghost_y_base(g, ND) + (long)f.k * g.Nx + f.i; f.yps = (long)(g.Nz + 1) * g.Nx; }
  else { f.yp = lidx(g, f.k, wrap_p(f.j, g.Ny), f.i); f.yps = g.nlev; }
  return f;
}

// offset of dof d's entry / exit face for a cell (entry: where the stream comes in, exit: its destination row)
template <class L>
__device__ __forceinline__ long diff_face(const Grid &g, const Faces &f, int d, bool exit_face) {
  constexpr int Nft = L::Nft, Nfs = L::Nfs;
  if (d < Nft) return (long)d * g.nlev + ((L::top_in(d) == exit_face) ? f.dn : f.own);
  if (d < Nft + Nfs) return (L::side_in(d - Nft) == exit_face) ? f.xp + (long)d * f.xps : f.own + (long)d * g.nlev;
  return (L::side_in(d - Nft - Nfs) == exit_face) ? f.yp + (long)d * f.yps : f.own + (long)d * g.nlev;
}
// true if dof d's destination row of this cell lies on a ghost face (owned by the neighbouring rank)
template <class L>
__device__ __forceinline__ bool diff_exit_is_ghost(const Faces &f, int d) {
  constexpr int Nft = L::Nft, Nfs = L::Nfs;
  if (d < Nft) return false;
  if (d < Nft + Nfs) return L::side_in(d - Nft) && f.xg;
  return L::side_in(d - Nft - Nfs) && f.yg;
}


// ---------------------------------------------------------------------------------------------------------------
// Deterministic grid-wide sum of NV doubles per thread: warp shuffles -> shared memory -> one partial per block;
// the last block to arrive (atomic ticket) adds the partials in a fixed order. No floating-point atomics, so the
// Krylov scalars are bit-reproducible run to run.  `counter` wraps back to 0 by itself (atomicInc).
// After the call, `is_last` is true in every thread of the finishing block and out[0..NV) holds the totals.
// ---------------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ bool grid_reduce_sum(double (&v)[NV], double *partials, unsigned int *counter,
                                                double *out) {
  __shared__ double sm[NV][32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int n = 0; n < NV; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[n] += __shfl_down_sync(0xffffffffu, v[n], o);
    if (lane == 0) sm[n][warp] = v[n];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      double t = (lane < nwarp) ? sm[n][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (lane == 0) partials[(long)blockIdx.x * NV + n] = t;
    }
    if (lane == 0) {
      __threadfence();
      unsigned int ticket = atomicInc(counter, gridDim.x - 1);
      s_last = (ticket == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) acc[n] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int n = 0; n < NV; ++n) acc[n] += __ldcg(&partials[(long)b * NV + n]);
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NV; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_down_sync(0xffffffffu, acc[n], o);
    if (lane == 0) sm[n][warp] = acc[n];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      double t = (lane < nwarp) ? sm[n][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (lane == 0) out[n] = t;
    }
  }
  __syncthreads();
  return true;
}
