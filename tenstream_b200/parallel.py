"""Work partitioning over the GPUs of one box (one process per GPU): band sharding and x/y domain decomposition.

Band sharding
-------------

Spectral bands / g-points are independent units of the pprts path (each ``solve_pprts(uid=ib)`` only touches
``solutions(uid)``, src/pprts.F90:2507-2509); the reference loops over them serially on one communicator
(rrtmg/rrtmg/pprts_rrtmg.F90:999-1053) and sums the weighted fluxes on the host (:1048-1051).  Here every rank takes
a contiguous block of bands, accumulates ``weight * (edn, eup, abso, edir)`` on its GPU and ONE all-reduce of the four
fields ends the spectral integration -- there is no collective on the data path.

x/y decomposition
-----------------
Large domains are split DMDA-style (src/pprts_base.F90:721-828): a periodic 2-D process grid, x fastest
(``rank = xi + yi*nxp``), even split ``xs = (xi*Nx)/nxp``; z is never split.  The helpers below are the host-side
index arithmetic of that decomposition (the device-side halo exchange lives in csrc/solver.cu) plus a reference halo
exchange over ``torch.distributed`` used by the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bands(nbands: int, world: int, rank: int) -> range:
    """Contiguous block distribution, remainder spread over the first ranks (same rule as the reference's
    cartesian split of columns, src/pprts_base.F90:776-782: xs = (xi*N)/nproc)."""
    lo = (rank * nbands) // world
    hi = ((rank + 1) * nbands) // world
    return range(lo, hi)


class BandShardedIntegrator:
    """Runs ``set_optical_properties -> solve_pprts -> accumulate`` for this rank's bands and reduces the sums.

    ``api`` is the module that provides the pprts calls (``tenstream_b200.pprts``); ``solver`` its solver object.
    ``reduce`` is "nccl" (the library's own NCCL all-reduce on the device-resident sums) or "torch"
    (``torch.distributed.all_reduce`` on the host copies, any backend -- used with gloo in the CPU tests)."""

    def __init__(self, api, solver, world: int = 1, rank: int = 0, reduce: str = "nccl"):
        self.api, self.solver, self.world, self.rank, self.reduce = api, solver, world, rank, reduce

    def integrate(self, bands):
        """bands: sequence of dicts with keys albedo,kabs,ksca,g,planck,planck_srfc,lsolar,edirTOA,uid,weight."""
        s = self.solver
        s.reset_accumulator()
        mine = shard_bands(len(bands), self.world, self.rank)
        for ib in mine:
            b = bands[ib]
            self.api.set_optical_properties(s, b["albedo"], b["kabs"], b["ksca"], b["g"], planck=b.get("planck"),
                                            planck_srfc=b.get("planck_srfc"))
            self.api.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            s.accumulate_result(b["uid"], b.get("weight", 1.0))
        if self.world > 1 and self.reduce == "nccl":
            s.allreduce_accumulated()
        out = s.get_accumulated()
        if self.world > 1 and self.reduce == "torch":
            import torch
            import torch.distributed as dist
            red = []
            for a in out:
                t = torch.from_numpy(np.ascontiguousarray(a))
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                red.append(t.numpy().reshape(a.shape))
            out = tuple(red)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# x/y decomposition
# ---------------------------------------------------------------------------------------------------------------------
def split_even(n: int, nproc: int) -> list[int]:
    """Columns per process along one axis: xs = (xi*n)/nproc, xm = ((xi+1)*n)/nproc - xs (pprts_base.F90:776-782)."""
    return [((i + 1) * n) // nproc - (i * n) // nproc for i in range(nproc)]


def proc_grid(world: int) -> tuple[int, int]:
    """(nxp, nyp) like MPI_Dims_create read as [nyp, nxp] with nyp >= nxp (pprts_base.F90:756-763)."""
    nxp = int(np.floor(np.sqrt(world)))
    while world % nxp:
        nxp -= 1
    return nxp, world // nxp


class CartDecomposition:
    """This rank's share of a periodic (Nx, Ny) domain on an nxp x nyp process grid (setup_coord_native)."""

    def __init__(self, Nx: int, Ny: int, nxp: int, nyp: int, rank: int):
        if not 0 <= rank < nxp * nyp:
            raise ValueError(f"rank {rank} outside a {nxp} x {nyp} grid")
        self.nxp, self.nyp, self.rank = nxp, nyp, rank
        self.xi, self.yi = rank % nxp, rank // nxp                      # x fastest, like PETSc DMDA
        self.nxproc, self.nyproc = split_even(Nx, nxp), split_even(Ny, nyp)
        self.xs, self.ys = sum(self.nxproc[:self.xi]), sum(self.nyproc[:self.yi])
        self.xm, self.ym = self.nxproc[self.xi], self.nyproc[self.yi]
        self.west = (self.xi - 1) % nxp + self.yi * nxp                 # neighbours(10 / 16 / 4 / 22)
        self.east = (self.xi + 1) % nxp + self.yi * nxp
        self.south = self.xi + ((self.yi - 1) % nyp) * nxp
        self.north = self.xi + ((self.yi + 1) % nyp) * nxp

    def local(self, a, xaxis=-2, yaxis=-1):
        """Local block of a global array whose ``xaxis`` / ``yaxis`` are the horizontal dimensions."""
        a = np.asarray(a)
        sl = [slice(None)] * a.ndim
        sl[xaxis] = slice(self.xs, self.xs + self.xm)
        sl[yaxis] = slice(self.ys, self.ys + self.ym)
        return a[tuple(sl)]


def halo_fill_reference(field, dec: CartDecomposition):
    """Reference ghost fill over torch.distributed (any backend): returns (ghost_x, ghost_y), the i = 0 plane of the
    east neighbour and the j = 0 plane of the north neighbour of ``field`` (..., xm, ym) -- what the device-side
    ``halo_fill`` writes into the ghost blocks of a state vector (src/pprts_explicit.F90:232-328 with the sun in
    the east / north)."""
    import torch
    import torch.distributed as dist
    f = np.ascontiguousarray(field)
    send_x = torch.from_numpy(np.ascontiguousarray(f[..., 0, :]))
    send_y = torch.from_numpy(np.ascontiguousarray(f[..., :, 0]))
    recv_x, recv_y = torch.empty_like(send_x), torch.empty_like(send_y)
    if dec.nxp == 1:
        recv_x.copy_(send_x)
    if dec.nyp == 1:
        recv_y.copy_(send_y)
    ops = []
    if dec.nxp > 1:
        ops += [dist.P2POp(dist.isend, send_x, dec.west), dist.P2POp(dist.irecv, recv_x, dec.east)]
    if dec.nyp > 1:
        ops += [dist.P2POp(dist.isend, send_y, dec.south), dist.P2POp(dist.irecv, recv_y, dec.north)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return recv_x.numpy(), recv_y.numpy()
