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


def gather_all_to_zero(local, dec: CartDecomposition):
    """``gather_all_toZero`` (src/pprts.F90:6325-6359): every rank announces (xs, ys, xm, ym) (the allgather of the
    six-integer meta record; z is never split), then ranks > 0 send their (nz, xm, ym) block to rank 0, which pastes
    it at its offsets.  Returns the global (nz, Nx, Ny) array on rank 0 and None elsewhere.  Host arrays over
    ``torch.distributed`` (any backend): the result arrays are already on the host after ``pprts_get_result``."""
    import torch
    import torch.distributed as dist
    a = np.ascontiguousarray(local, dtype=np.float64)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return a.copy()
    meta = torch.tensor([dec.xs, dec.ys, dec.xm, dec.ym], dtype=torch.int64)
    all_meta = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_meta, meta)
    rank = dist.get_rank()
    if rank != 0:
        dist.send(torch.from_numpy(a), dst=0, tag=77)
        return None
    nz = a.shape[0]
    Nx = max(int(m[0] + m[2]) for m in all_meta)
    Ny = max(int(m[1] + m[3]) for m in all_meta)
    out = np.zeros((nz, Nx, Ny))
    out[:, dec.xs:dec.xs + dec.xm, dec.ys:dec.ys + dec.ym] = a
    for src in range(1, world):
        xs, ys, xm, ym = (int(v) for v in all_meta[src])
        buf = torch.empty((nz, xm, ym), dtype=torch.float64)
        dist.recv(buf, src=src, tag=77)
        out[:, xs:xs + xm, ys:ys + ym] = buf.numpy()
    return out


def pprts_get_result_toZero(api, solver, dec: CartDecomposition, opt_solution_uid=0):
    """``pprts_get_result_toZero`` (src/pprts.F90:6265-6323): local results of every rank, gathered on rank 0 as
    global (Nz[+1], Nx, Ny) arrays -> (gedn, geup, gabso, gedir) there, four ``None`` elsewhere."""
    edn, eup, abso, edir = api.pprts_get_result(solver, opt_solution_uid)
    return tuple(gather_all_to_zero(a, dec) for a in (edn, eup, abso, edir))


def gather_all_to_all(local, dec: CartDecomposition):
    """``gather_all_to_all`` (src/pprts.F90:6361-6401): as ``gather_all_toZero`` but every rank ends up with the global
    (nz, Nx, Ny) array (one allgather of the meta records, one of the blocks)."""
    import torch
    import torch.distributed as dist
    a = np.ascontiguousarray(local, dtype=np.float64)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return a.copy()
    meta = torch.tensor([dec.xs, dec.ys, dec.xm, dec.ym], dtype=torch.int64)
    all_meta = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_meta, meta)
    nz = a.shape[0]
    nmax = max(int(m[2] * m[3]) for m in all_meta) * nz          # equal-sized slots (backends without allgatherv)
    slot = torch.zeros(nmax, dtype=torch.float64)
    slot[:a.size] = torch.from_numpy(a.reshape(-1))
    slots = [torch.zeros(nmax, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(slots, slot)
    Nx = max(int(m[0] + m[2]) for m in all_meta)
    Ny = max(int(m[1] + m[3]) for m in all_meta)
    out = np.zeros((nz, Nx, Ny))
    for m, buf in zip(all_meta, slots):
        xs, ys, xm, ym = (int(v) for v in m)
        out[:, xs:xs + xm, ys:ys + ym] = buf[:nz * xm * ym].numpy().reshape(nz, xm, ym)
    return out


def extend_arr(arr, minimal_dimension: int = 3):
    """``extend_arr`` of set_global_optical_properties (src/pprts.F90:2453-2483): a horizontal dimension of extent 1 is
    replicated to the minimal 3 columns the periodic boundary conditions need; anything else below 3 is an error."""
    a = np.asarray(arr)
    if a.shape[1] == 1:
        a = np.repeat(a, minimal_dimension, axis=1)
    if a.shape[2] == 1:
        a = np.repeat(a, minimal_dimension, axis=2)
    if min(a.shape) < minimal_dimension:
        raise ValueError("set_optprop -> extend_arr :: dimension is smaller than we support.. please think of something here")
    return a


def scatter_from_zero(global_arr, dec: CartDecomposition):
    """``scatterZerotoPetscGlobal`` + ``petscVecToF90`` as set_global_optical_properties uses them
    (src/pprts.F90:2397-2420): rank 0 holds the global (nz, Nx, Ny) array (``None`` elsewhere) and every rank receives
    its (nz, xm, ym) block.  Host arrays over ``torch.distributed`` (any backend)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return np.array(dec.local(global_arr), dtype=np.float64, order="F")
    rank = dist.get_rank()
    meta = torch.tensor([dec.xs, dec.ys, dec.xm, dec.ym], dtype=torch.int64)
    all_meta = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_meta, meta)
    nz = torch.tensor([0 if global_arr is None else np.shape(global_arr)[0]], dtype=torch.int64)
    dist.broadcast(nz, src=0)
    nz = int(nz[0])
    if rank == 0:
        g = np.asarray(global_arr, dtype=np.float64)
        for dst in range(1, world):
            xs, ys, xm, ym = (int(v) for v in all_meta[dst])
            dist.send(torch.from_numpy(np.ascontiguousarray(g[:, xs:xs + xm, ys:ys + ym])), dst=dst, tag=78)
        return np.array(dec.local(g), dtype=np.float64, order="F")
    buf = torch.empty((nz, dec.xm, dec.ym), dtype=torch.float64)
    dist.recv(buf, src=0, tag=78)
    return np.asfortranarray(buf.numpy())


def set_global_optical_properties(api, solver, dec: CartDecomposition, global_albedo=None, global_kabs=None,
                                  global_ksca=None, global_g=None, global_planck=None):
    """``set_global_optical_properties`` (src/pprts.F90:2341-2451): rank 0 holds the optical properties of the whole
    domain -- (Nz, Nx, Ny) resp. (Nz+1, Nx, Ny) for planck, ``None`` on the other ranks --, every rank receives its
    block (which fields exist and the albedo are broadcast from rank 0, :2359-2362, :2393) and calls
    ``set_optical_properties`` on it.  A horizontal dimension of extent 1 is replicated to 3 first (``extend_arr``)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    fields = [global_kabs, global_ksca, global_g, global_planck]
    head = torch.tensor([float(f is not None) for f in fields] + [0.0 if global_albedo is None else float(global_albedo)],
                        dtype=torch.float64)
    if world > 1:
        dist.broadcast(head, src=0)
    have = [bool(v) for v in head[:4]]
    if not all(have[:3]):
        raise ValueError("set_global_optical_properties: kabs, ksca and g are required (set_optical_properties needs them)")
    rank0 = world == 1 or dist.get_rank() == 0
    local = [scatter_from_zero(extend_arr(f) if rank0 else None, dec) if h else None for f, h in zip(fields, have)]
    api.set_optical_properties(solver, float(head[4]), local[0], local[1], local[2], planck=local[3])

