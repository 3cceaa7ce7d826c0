"""Band sharding over the GPUs of one box (one process per GPU).

Spectral bands / g-points are independent units of the pprts path (each ``solve_pprts(uid=ib)`` only touches
``solutions(uid)``, src/pprts.F90:2507-2509); the reference loops over them serially on one communicator
(rrtmg/rrtmg/pprts_rrtmg.F90:999-1053) and sums the weighted fluxes on the host (:1048-1051).  Here every rank takes
a contiguous block of bands, accumulates ``weight * (edn, eup, abso, edir)`` on its GPU and ONE all-reduce of the four
fields ends the spectral integration -- there is no collective on the data path.
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
