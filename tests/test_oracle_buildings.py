"""Buildings (opt_buildings of solve_pprts / pprts_get_result) in the oracle and in the host mirror of m_buildings.

Restates the reference's own tests on the synthetic table: tests/test_buildings/test_buildings.F90 --
test_init_buildings / test_clone_buildings (:43-149) and the known-answer examples on the 6x6x3 vacuum domain with one
hovering box (:151-420, set up by examples/pprts/pprts_buildings.F90:46-200): overhead sun, box albedo, emission.
"""
import numpy as np
import pytest

from oracle import binding as ob
from tenstream_b200 import buildings as B


# ---- m_buildings host mirror -------------------------------------------------------------------------------------------
def test_init_buildings():
    """test_buildings.F90:43-77"""
    Nfaces = 6
    b = B.init_buildings((6, 1, 1, 1), Nfaces)
    assert len(b.iface) == Nfaces and len(b.albedo) == Nfaces
    b.iface[:] = np.arange(1, Nfaces + 1)
    b.albedo[:] = np.arange(1, Nfaces + 1)
    for i in range(Nfaces):
        assert b.iface[i] == i + 1 and b.albedo[i] == float(i + 1)
    assert list(b.da_offsets) == [1, 6, 6, 6]


def test_clone_buildings():
    """test_buildings.F90:79-149: the face list is shared, the data is copied only on request"""
    b1 = B.init_buildings((6, 1, 1, 1), 6)
    b1.iface[:] = np.arange(1, 7)
    b1.albedo[:] = np.arange(1, 7)
    b2 = B.clone_buildings(b1, True)
    b3 = B.clone_buildings(b1, False)
    assert b2.iface is b1.iface and b3.iface is b1.iface
    assert np.array_equal(b2.albedo, b1.albedo) and b2.albedo is not b1.albedo
    assert b3.albedo is None


def test_face_index_round_trip_and_consistency():
    Nz, Nx, Ny = 3, 4, 5
    off = B.ndarray_offsets((6, Nz, Nx, Ny))
    assert list(off) == [1, 6, 18, 72]
    seen = set()
    for j in range(1, Ny + 1):
        for i in range(1, Nx + 1):
            for k in range(1, Nz + 1):
                for f in range(1, 7):
                    m = B.faceidx_by_cell_plus_offset(off, k, i, j, f)
                    assert B.ind_1d_to_nd(off, m) == [f, k, i, j]
                    seen.add(m)
    assert seen == set(range(1, 6 * Nz * Nx * Ny + 1))
    b = B.box_buildings(Nz, Nx, Ny, [(2, 2, 3)], 0.3)
    assert B.check_buildings_consistency(b, Nz, Nx, Ny) == 0
    b.iface = b.iface[:5]                                     # a semi-open box (src/buildings.F90:252-262)
    assert B.check_buildings_consistency(b, Nz, Nx, Ny) == 1
    assert B.buildingid2str(B.PPRTS_REAR_FACE) == "PPRTS_REAR_FACE"
    with pytest.raises(ValueError):
        B.buildingid2str(7)


# ---- the reference's example domain ----------------------------------------------------------------------------------------
Nx = Ny = 6
Nlay = 3
BOX = (2, 3, 3)            # glob_box_k, glob_box_i, glob_box_j (1-based)
DX = DZ = 100.0
S0 = 1.0
ATOL = 1e-5


def ex_pprts_buildings(oracle_lut, lthermal, lsolar, box_albedo, box_planck, phi0, theta0, Ag, dtau, w0):
    """examples/pprts/pprts_buildings.F90:46-200 for one box cell"""
    sd = ob.sundir_from_angles(phi0, theta0)
    s = ob.OracleSolver(oracle_lut, Nlay, Nx, Ny, DX, DX, sd, np.full(Nlay, DZ), diff_solver=0, rtol=1e-10, atol=1e-13)
    kabs = np.full((Nlay, Nx, Ny), dtau * (1 - w0) / DZ / Nlay)
    ksca = np.full((Nlay, Nx, Ny), dtau * w0 / DZ / Nlay)
    g = np.zeros((Nlay, Nx, Ny))
    plck = np.full((Nlay + 1, Nx, Ny), 100.0 / np.pi) if lthermal else None
    bld = B.box_buildings(Nlay, Nx, Ny, [BOX], box_albedo, box_planck if lthermal else None)
    assert B.check_buildings_consistency(bld, Nlay, Nx, Ny) == 0
    s.set_optical_properties(Ag, kabs, ksca, g, planck=plck)
    s.set_buildings(bld.iface, bld.albedo, bld.planck)
    rc = s.solve(lthermal, lsolar, S0, 0)
    assert rc == 0
    edn, eup, abso, edir = s.get_result(0)
    bld.edir, bld.incoming, bld.outgoing = s.get_buildings_result(0)
    s.destroy()
    return edir, edn, eup, abso, bld


def test_buildings_example_overhead_sun(oracle_lut):
    """test_buildings.F90:151-207"""
    k, i, j = BOX[0] - 1, BOX[1] - 1, BOX[2] - 1
    edir, edn, eup, abso, bld = ex_pprts_buildings(oracle_lut, False, True, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    assert edir[Nlay, i, j] == pytest.approx(0, abs=ATOL)          # beneath the building
    assert edir[k + 1, i, j] == pytest.approx(0, abs=ATOL)         # at its bottom
    assert abso[k, i, j] == pytest.approx(S0 / DZ, abs=ATOL)       # the box absorbs all incoming solar radiation
    assert bld.edir[0] == pytest.approx(S0, abs=ATOL)
    assert np.abs(bld.edir[1:]).max() <= ATOL
    assert np.abs(bld.incoming).max() <= ATOL and np.abs(bld.outgoing).max() <= ATOL


def test_buildings_example_albedo(oracle_lut):
    """test_buildings.F90:209-267"""
    k, i, j = BOX[0] - 1, BOX[1] - 1, BOX[2] - 1
    box_albedo = 0.5
    edir, edn, eup, abso, bld = ex_pprts_buildings(oracle_lut, False, True, box_albedo, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0)
    assert edir[Nlay, i, j] == pytest.approx(0, abs=ATOL)
    assert edir[k + 1, i, j] == pytest.approx(0, abs=ATOL)
    assert edir[k, i, j] == pytest.approx(S0, abs=ATOL)
    assert eup[k, i, j] == pytest.approx(S0 * box_albedo, abs=ATOL)
    assert bld.edir[0] == pytest.approx(S0, abs=ATOL)
    assert np.abs(bld.edir[1:]).max() <= ATOL
    np.testing.assert_allclose((bld.edir + bld.incoming) * box_albedo, bld.outgoing, atol=ATOL)
    # -pprts_check_buildings_energy_balance (src/pprts.F90:5942-5984): passes at the test's tolerance, fails when a face
    # does not reflect what hits it, off by default and for thermal results
    bal = B.check_buildings_energy_balance(bld, True, ATOL)
    assert bal.shape == bld.iface.shape and np.abs(bal).max() < ATOL
    assert B.check_buildings_energy_balance(bld, True) is None and B.check_buildings_energy_balance(bld, False, ATOL) is None
    bld.outgoing[2] += 0.1
    with pytest.raises(ValueError, match=r"failed energy balance on buildings faces \(1\)[\s\S]*face nr. 3 .*PPRTS_LEFT_FACE"):
        B.check_buildings_energy_balance(bld, True, ATOL)


def test_buildings_example_emission(oracle_lut):
    """test_buildings.F90:269-320"""
    k, i, j = BOX[0] - 1, BOX[1] - 1, BOX[2] - 1
    box_planck = 1.0
    edir, edn, eup, abso, bld = ex_pprts_buildings(oracle_lut, True, False, 0.0, box_planck, 0.0, 180.0, 0.0, 0.0, 0.0)
    assert eup[k, i, j] == pytest.approx(box_planck * np.pi, abs=ATOL)
    assert edn[k + 1, i, j] == pytest.approx(box_planck * np.pi, abs=ATOL)
    np.testing.assert_allclose(bld.outgoing, box_planck * np.pi, atol=ATOL)


def test_buildings_block_and_shadow(oracle_lut):
    """test_buildings.F90:322-420 (45 degree sun): the synthetic table does not reproduce the exact clear-sky beam
    geometry of the real LUT, so only the table-independent statements are kept: nothing leaves the box, its lit faces
    receive the beam, the unlit faces and all diffuse fluxes are zero (box albedo 0, vacuum, black ground)."""
    k, i, j = BOX[0] - 1, BOX[1] - 1, BOX[2] - 1
    for phi0, lit, dark in ((0.0, 5, 4), (180.0, 4, 5)):   # phi = 0: sun in the north -> FRONT (y+1) face is lit
        edir, edn, eup, abso, bld = ex_pprts_buildings(oracle_lut, False, True, 0.0, 0.0, phi0, 45.0, 0.0, 0.0, 0.0)
        assert edir[k + 1, i, j] == pytest.approx(0, abs=ATOL)
        assert bld.edir[0] > 0.1 and bld.edir[lit] > 0.1
        assert abs(bld.edir[dark]) <= ATOL and np.abs(bld.edir[1:4]).max() <= ATOL
        assert np.abs(bld.incoming).max() <= ATOL and np.abs(bld.outgoing).max() <= ATOL
        # everything that enters the box is absorbed there: abso * dz = (top + lit side flux per horizontal area)
        assert abso[k, i, j] * DZ == pytest.approx(bld.edir[0] + bld.edir[lit] * DZ / DX, abs=ATOL)


def test_removing_buildings_restores_the_plain_solve(oracle_lut):
    sd = ob.sundir_from_angles(20.0, 30.0)
    s = ob.OracleSolver(oracle_lut, Nlay, Nx, Ny, DX, DX, sd, np.full(Nlay, DZ), diff_solver=0, rtol=1e-10, atol=1e-13)
    kabs = np.full((Nlay, Nx, Ny), 1e-4)
    ksca = np.full((Nlay, Nx, Ny), 2e-3)
    g = np.full((Nlay, Nx, Ny), 0.3)
    s.set_optical_properties(0.2, kabs, ksca, g)
    assert s.solve(False, True, 100.0, 0) == 0
    base = s.get_result(0)
    bld = B.box_buildings(Nlay, Nx, Ny, [BOX], 0.4)
    s.set_buildings(bld.iface, bld.albedo, None)
    assert s.solve(False, True, 100.0, 1) == 0
    withb = s.get_result(1)
    assert np.abs(withb[1] - base[1]).max() > 1e-3
    s.set_buildings(None)
    assert s.solve(False, True, 100.0, 2) == 0
    again = s.get_result(2)
    for a, b in zip(base, again):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)
