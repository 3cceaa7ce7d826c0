// common.cuh -- grid descriptors, stream layouts and the deterministic block reduction shared by all kernels.
//
// Data layout in HBM (DESIGN.md "Data layout"): everything is structure-of-arrays with x fastest so that a warp
// of 32 neighbouring cells/columns reads 128/256 contiguous bytes from every plane:
//   cell  field  f[(k*Ny + j)*Nx + i]                      k = 0..Nz-1 (TOA first)
//   level field  f[(k*Ny + j)*Nx + i]                      k = 0..Nz
//   state        x[d*nlev + (k*Ny + j)*Nx + i]             d = stream (dof), one plane per dof
//   coefficients c[n*ncp + cell]                           n = dst*Nsrc + src (reference order), ncp = padded ncell
// A side dof with index (k,i,j) lives on the low-x / low-y face of cell (k,i,j), exactly as in the reference
// (src/pprts_explicit.F90:398-451); face i+1 of the last column is the periodic image (one GPU) or a ghost.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PB_MAX_DOF 24
#define PB_MAX_DIRDOF 8

struct Grid {
  int Nz, Nx, Ny;
  long ncol;   // Nx*Ny
  long ncell;  // Nz*ncol
  long nlev;   // (Nz+1)*ncol
  long ncp;    // ncell padded to a multiple of 32 (plane stride of coefficient arrays)
  double dx, dy;
};

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

__device__ __forceinline__ long lidx(const Grid &g, int k, int j, int i) { return ((long)k * g.Ny + j) * g.Nx + i; }
__device__ __forceinline__ int wrap_p(int i, int n) { return (i + 1 == n) ? 0 : i + 1; }

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
