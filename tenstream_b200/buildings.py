"""Host mirror of the reference's m_buildings (src/buildings.F90): the `opt_buildings` argument of solve_pprts and
pprts_get_result.

A building is a set of faces of grid cells.  `iface[m]` is the 1-based flattened index of face m in an array of shape
`da_sizes = (6, Nz, Nx, Ny)` with the first axis fastest (`faceidx_by_cell_plus_offset`, src/buildings.F90:215-222),
the face id being one of PPRTS_TOP_FACE .. PPRTS_FRONT_FACE (src/boxmc_geometry.F90:45-51).  `albedo[m]` is the
Lambertian albedo of the face, `planck[m]` (optional) its emitted Planck radiance; after pprts_get_result
`edir / incoming / outgoing` hold the direct flux hitting the face and the diffuse flux entering / leaving the
building through it, in W/m2 (src/pprts.F90:6011-6160).
"""
from __future__ import annotations

import numpy as np

PPRTS_TOP_FACE = 1    # z+0
PPRTS_BOT_FACE = 2    # z+1
PPRTS_LEFT_FACE = 3   # x+0
PPRTS_RIGHT_FACE = 4  # x+1
PPRTS_REAR_FACE = 5   # y+0
PPRTS_FRONT_FACE = 6  # y+1

_NAMES = {1: "PPRTS_TOP_FACE", 2: "PPRTS_BOT_FACE", 3: "PPRTS_LEFT_FACE", 4: "PPRTS_RIGHT_FACE",
          5: "PPRTS_REAR_FACE", 6: "PPRTS_FRONT_FACE"}


def ndarray_offsets(sizes):
    """helper_functions.fypp:2432-2437: offsets(1) = 1, offsets(i) = offsets(i-1) * sizes(i-1)."""
    off = np.ones(len(sizes), dtype=np.int64)
    for i in range(1, len(sizes)):
        off[i] = off[i - 1] * sizes[i - 1]
    return off


def ind_nd_to_1d(offsets, idx):
    """helper_functions.fypp:2456-2466 (1-based in, 1-based out)."""
    return int(1 + sum((int(i) - 1) * int(o) for i, o in zip(idx, offsets)))


def ind_1d_to_nd(offsets, ind):
    """helper_functions.fypp:2439-2454 (1-based in, 1-based out)."""
    r = int(ind) - 1
    out = [0] * len(offsets)
    for d in range(len(offsets) - 1, -1, -1):
        out[d] = r // int(offsets[d]) + 1
        r = r % int(offsets[d])
    return out


def faceidx_by_cell_plus_offset(da_offsets, k, i, j, face_id):
    """src/buildings.F90:215-222; k, i, j are 1-based cell indices."""
    return ind_nd_to_1d(da_offsets, (face_id, k, i, j))


def buildingid2str(bid):
    if bid not in _NAMES:
        raise ValueError(f"unknown building face_idx {bid}")
    return _NAMES[bid]


class PprtsBuildings:
    """t_pprts_buildings (src/buildings.F90:40-69)."""

    def __init__(self, da_sizes, Nfaces):
        self.da_sizes = tuple(int(n) for n in da_sizes)
        self.da_offsets = ndarray_offsets(self.da_sizes)
        self.iface = np.full(Nfaces, -1, dtype=np.int32)
        self.albedo = np.full(Nfaces, -1.0)
        self.temp = None
        self.planck = None
        self.edir = None
        self.incoming = None
        self.outgoing = None


def init_buildings(da_sizes, Nfaces):
    """src/buildings.F90:73-93."""
    return PprtsBuildings(da_sizes, Nfaces)


def clone_buildings(buildings, l_copy_data=True):
    """src/buildings.F90:95-131: the face list is shared, the data arrays are copied on request."""
    b = PprtsBuildings(buildings.da_sizes, len(buildings.iface))
    b.iface = buildings.iface
    if l_copy_data:
        b.albedo = buildings.albedo.copy()
        for name in ("temp", "planck"):
            v = getattr(buildings, name)
            setattr(b, name, None if v is None else v.copy())
    else:
        b.albedo = None
    return b


def check_buildings_consistency(buildings, Nz, Nx, Ny):
    """src/buildings.F90:224-267: returns the number of semi-open cells (cells with some but not all six faces)."""
    off = buildings.da_offsets
    if off[1] != 6 or off[2] != 6 * Nz or off[3] != 6 * Nz * Nx:
        raise ValueError(f"da_offsets {off} do not belong to a (6,{Nz},{Nx},{Ny}) grid")
    sides = np.zeros((6, Nz, Nx, Ny), dtype=bool)
    flat = sides.reshape(-1, order="F")
    flat[np.asarray(buildings.iface, dtype=np.int64) - 1] = True
    sides = flat.reshape((6, Nz, Nx, Ny), order="F")
    anyf, allf = sides.any(axis=0), sides.all(axis=0)
    return int((anyf & ~allf).sum())


def box_buildings(Nz, Nx, Ny, cells, albedo, planck=None):
    """All six faces of every cell in `cells` (iterable of 1-based (k,i,j)), uniform albedo / planck: the construction
    of examples/pprts/pprts_buildings.F90:152-172."""
    cells = list(cells)
    b = init_buildings((6, Nz, Nx, Ny), 6 * len(cells))
    m = 0
    for (k, i, j) in cells:
        for face in range(1, 7):
            b.iface[m] = faceidx_by_cell_plus_offset(b.da_offsets, k, i, j, face)
            b.albedo[m] = albedo
            m += 1
    if planck is not None:
        b.planck = np.full(len(b.iface), float(planck))
    return b


def check_buildings_energy_balance(buildings, lsolar, balance_limit=-1.0):
    """pprts_get_result::check_buildings_energy_balance (src/pprts.F90:5942-5984) with the value of
    ``-pprts_check_buildings_energy_balance`` as ``balance_limit`` (negative: off, the default).  For a solar result
    every face has to reflect what hits it: ``(edir + incoming) * albedo - outgoing``; faces whose imbalance reaches
    the limit are reported and the check fails, as CHKERR(errcnt) does in the reference.  Thermal results are not
    checked there either.  Returns the per-face balance (None when the check is off or not solar)."""
    if balance_limit < 0 or not lsolar:
        return None
    b = buildings
    balance = (np.asarray(b.edir) + np.asarray(b.incoming)) * np.asarray(b.albedo) - np.asarray(b.outgoing)
    bad = np.flatnonzero(np.abs(balance) >= balance_limit)
    if len(bad):
        lines = []
        for m in bad:
            face, k, i, j = ind_1d_to_nd(b.da_offsets, b.iface[m])
            lines.append(f"face nr. {m + 1} faceidx {int(b.iface[m])} cell_kij {(k, i, j)} {buildingid2str(face)} "
                         f"edir {b.edir[m]} ediff {b.incoming[m]} albedo {b.albedo[m]} outgoing {b.outgoing[m]} "
                         f"balance {balance[m]}")
        raise ValueError(f"failed energy balance on buildings faces ({len(bad)}):\n" + "\n".join(lines))
    return balance
