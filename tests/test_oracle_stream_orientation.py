"""The chain sun azimuth -> xinc / yinc -> stream permutations -> table orientation, checked on its physics.

The tables hold the beam in ONE orientation (towards +x / +y, pinned on the reference's boxmc known answers in
tests/test_lut_stream_conventions.py); for any other sun azimuth pprts mirrors the stream indices
(lswitch_east = xinc == 0, lswitch_north = yinc == 0, src/pprts.F90:1157-1167, :3161, src/optprop.F90:1009-1045).  A mirror
applied the wrong way round keeps every symmetry test happy -- the solution is then the mirror image of the right one in the
diffuse streams only -- but it sends the forward-scattered light of a cloud UPSTREAM.  One strongly forward-scattering cell in
a vacuum domain, sun from the four cardinal directions: the diffuse energy that leaves the cell sideways must leave it on
the side the beam leaves it, in the stream that travels that way, and land on the ground downstream of the cell.

Stream numbering of 3_10 (src/pprts.F90:256-443, 0-based): 2 / 3 = downward through an x face travelling -x / +x, 6 / 7 the same
in y; a stream travelling towards + leaves cell i through face i+1, one travelling towards - through face i.
(Checked to discriminate: with the dir2diff permutation of the oracle applied the wrong way round -- `xinc != 0` in
alloc_coeff_dir2diff -- all four cases fail.)
sundir = the direction the photons travel (spherical_2_cartesian): phi0 = 0 sun in the north (towards -y), 90 in the east
(towards -x)."""
import numpy as np
import pytest

from oracle import binding as ob

Nz, Nx, Ny = 5, 9, 9
K, I, J = 1, 4, 4
DX = DZ = 100.0


def solve(oracle_lut, phi0, theta0=60.0):
    s = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(phi0, theta0), np.full(Nz, DZ), diff_solver=0,
                        rtol=1e-8)
    kabs = np.zeros((Nz, Nx, Ny))
    ksca = np.zeros((Nz, Nx, Ny))
    g = np.zeros((Nz, Nx, Ny))
    ksca[K, I, J] = 1.0 / DZ
    g[K, I, J] = 0.85
    s.set_optical_properties(0.0, kabs, ksca, g)
    assert s.solve(False, True, 1000.0) == 0
    ediff = s.state(2)
    edn = s.get_result()[0]
    s.destroy()
    return ediff, edn


# phi0 -> (axis the beam runs along, its sign)
CASES = {90.0: ("x", -1), 270.0: ("x", +1), 0.0: ("y", -1), 180.0: ("y", +1)}


@pytest.mark.parametrize("phi0", sorted(CASES))
def test_forward_scattered_light_leaves_downstream(oracle_lut, phi0):
    axis, sign = CASES[phi0]
    ediff, edn = solve(oracle_lut, phi0)
    minus, plus = (2, 3) if axis == "x" else (6, 7)

    def face(dof, towards_plus):          # the cloud cell's face that a stream travelling that way leaves through
        if axis == "x":
            return ediff[dof, K, I + 1 if towards_plus else I, J]
        return ediff[dof, K, I, J + 1 if towards_plus else J]
    down = face(plus, True) if sign > 0 else face(minus, False)
    up = face(minus, False) if sign > 0 else face(plus, True)
    assert down > 0 and up > 0
    assert down > 3.0 * up, (phi0, down, up)            # isotropic part both ways, the forward peak only downstream
    # across the beam the two sides are mirror images
    if axis == "x":
        a, b = ediff[6, K, I, J], ediff[7, K, I, J + 1]
    else:
        a, b = ediff[2, K, I, J], ediff[3, K, I + 1, J]
    assert a == pytest.approx(b, rel=1e-6)
    # the diffuse light reaches the ground downstream of the cell, not upstream of it
    srf = edn[Nz]
    prof = srf.sum(axis=1) if axis == "x" else srf.sum(axis=0)
    c = I if axis == "x" else J
    ahead = sum(prof[(c + sign * d) % len(prof)] for d in (1, 2, 3))
    behind = sum(prof[(c - sign * d) % len(prof)] for d in (1, 2, 3))
    assert ahead > 1.2 * behind, (phi0, ahead, behind)


# ---- 8 direct streams: the sub-face permutations (dir2dir permutes sources AND destinations, src/optprop.F90:1268-1301) -------
@pytest.fixture(scope="module")
def lut810(tmp_path_factory):
    from tenstream_b200 import lut as L
    cfg = L.LutConfig(solver_id=810, tau=L.PRESET_TAU31[::3].copy(), w0=[0.0, 0.5, 0.99999], aspect=[0.25, 0.5, 1.0, 2.0],
                      g=[0.0, 0.5], phi=[0., 45., 90.], theta=[0., 45., 90.])
    paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut810o") / "LUT"))
    return ob.OracleLut(cfg, paths)


# phi0 -> (offset of the cell downstream of the cloud, top quadrants in its shadow, top quadrants in the light);
# quadrants (src/boxmc_8_10.inc:27-35): 0 x-lo/y-lo, 1 x-hi/y-lo, 2 x-lo/y-hi, 3 x-hi/y-hi
SHADOW = {270.0: ((1, 0), (0, 2), (1, 3)), 90.0: ((-1, 0), (1, 3), (0, 2)),
          180.0: ((0, 1), (0, 1), (2, 3)), 0.0: ((0, -1), (2, 3), (0, 1))}


@pytest.mark.parametrize("phi0", sorted(SHADOW))
def test_half_cell_shadow_falls_on_the_downstream_quadrants(lut810, phi0):
    """dz = dx / 2 and a 45 degree sun: the beam moves half a cell per layer, so one layer below an opaque cell the cell
    downstream of it has the half next to the cloud in the shadow and the far half in the light.  (Fails in all four cases
    when the oracle applies the dir2dir permutation the wrong way round.)"""
    (di, dj), dark, lit = SHADOW[phi0]
    s = ob.OracleSolver(lut810, Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(phi0, 45.0), np.full(Nz, DX / 2), solver_id=810,
                        diff_solver=0, rtol=1e-8)
    kabs = np.zeros((Nz, Nx, Ny))
    kabs[K, I, J] = 1.0                       # tau = 50: nothing gets through
    s.set_optical_properties(0.0, kabs, np.zeros((Nz, Nx, Ny)), np.zeros((Nz, Nx, Ny)))
    assert s.solve(False, True, 1000.0) == 0
    edir = s.state(1)
    s.destroy()
    free = edir[:4, K + 1, 0, 0]              # far away from the cloud
    assert free.min() > 0
    here = edir[:4, K + 1, I + di, J + dj]
    for q in dark:
        assert here[q] < 0.1 * free[q], (phi0, q, here, free)      # rays that only clip the cloud's edge get through
    for q in lit:
        assert here[q] == pytest.approx(free[q], rel=1e-6), (phi0, q, here, free)
    assert edir[:4, K + 1, I, J].max() < 0.1 * free.min()           # under the cloud itself: dark


# ---- thermal emission: which level's Planck value an optically thick layer shows -------------------------------------------
@pytest.mark.parametrize("dx", [100.0, 10.0])      # 3-D cells (dz / dx = 1) and 1-D layers (dz / dx = 10 > twostr_ratio)
def test_thick_layers_emit_the_planck_value_of_the_near_level(oracle_lut, dx):
    """set_thermal_source (src/pprts.F90:4848-4987): B_eff(far, near, tau) -- the upward stream leaving a layer through its
    top carries the layer's emission weighted towards the TOP level, the downward one towards the BOTTOM level.  With
    tau = 20 per layer the flux through a level is pi x its own Planck value plus a gradient term of order 1 / tau; mixing up
    the two levels would show the neighbouring level's value (a factor 2 at the top of this profile)."""
    nz = 4
    s = ob.OracleSolver(oracle_lut, nz, 3, 3, dx, dx, ob.sundir_from_angles(0., 0.), np.full(nz, 100.0), diff_solver=0,
                        rtol=1e-8)
    assert s.l1d.all() == (dx < 50)
    kabs = np.full((nz, 3, 3), 0.2)
    z = np.zeros((nz, 3, 3))
    planck = (np.array([10., 20., 30., 40., 50.]) / np.pi)[:, None, None] * np.ones((1, 3, 3))
    s.set_optical_properties(0.0, kabs, z, z, planck=planck)
    assert s.solve(True, False, 0.0) == 0
    edn, eup, abso, _ = s.get_result()
    s.destroy()
    assert eup[0, 1, 1] == pytest.approx(10.0, rel=0.12) and eup[0, 1, 1] >= 10.0       # the level's own value, from below
    assert edn[nz, 1, 1] == pytest.approx(50.0, rel=0.12) and edn[nz, 1, 1] <= 50.0
    for k in (1, 2, 3):                                   # inside: up from the warmer side, down from the colder one
        b = 10.0 * (k + 1)
        assert b <= eup[k, 1, 1] < b * 1.12 and b * 0.88 < edn[k, 1, 1] <= b


# ---- the direct beam's normalisation ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("phi0,theta0", [(0., 0.), (40., 30.), (200., 60.)])
def test_direct_flux_through_horizontal_faces(oracle_lut, phi0, theta0):
    """setup_incSolar puts edirTOA x dx x dy on the top faces (no cos, src/pprts_base.F90:1146-1181) and pprts_get_result
    multiplies by mu (src/pprts.F90:5882-5887): the flux through a horizontal face at the top is S0 cos(theta); a vacuum
    passes all of it down to the ground; absorbing 1-D layers follow Beer-Lambert along the slant path"""
    mu = np.cos(np.deg2rad(theta0))
    nz = 4
    z = np.zeros((nz, 5, 5))
    s = ob.OracleSolver(oracle_lut, nz, 5, 5, 100., 100., ob.sundir_from_angles(phi0, theta0), np.full(nz, 100.0),
                        diff_solver=0, rtol=1e-8)
    s.set_optical_properties(0.0, z, z, z)
    assert s.solve(False, True, 1364.0) == 0
    edir = s.get_result()[3]
    s.destroy()
    np.testing.assert_allclose(edir, 1364.0 * mu, rtol=1e-6)        # tau clamps to the table's 1e-10
    s = ob.OracleSolver(oracle_lut, nz, 5, 5, 10., 10., ob.sundir_from_angles(phi0, theta0), np.full(nz, 100.0),
                        diff_solver=0, rtol=1e-8)
    assert s.l1d.all()
    kabs = np.full((nz, 5, 5), 3e-3)
    s.set_optical_properties(0.0, kabs, z, z)
    assert s.solve(False, True, 1364.0) == 0
    edir = s.get_result()[3]
    s.destroy()
    want = 1364.0 * mu * np.exp(-0.3 * np.arange(nz + 1) / mu)
    np.testing.assert_allclose(edir, want[:, None, None] * np.ones((1, 5, 5)), rtol=1e-10)


def test_all_1d_layers_obey_the_twostream_statements(oracle_lut):
    """tests/test_twostr/test_twostr.F90:54-59 for a pprts domain whose layers are all 1-D (dz / dx > twostr_ratio): the
    beam at the ground is S0 exp(-tau / mu0) (x mu0 in pprts' flux through the horizontal face), nothing diffuse comes in at
    the top, the ground reflects albedo x (direct + diffuse down); dtau = 1 / ke, w0 = g = 0.5, mu0 = 0.5, albedo = 0.1.
    Delta scaling is switched off so that tau along the beam stays the optical depth of the test."""
    ke, mu0, S0, albedo = 9, 0.5, 1e2, 0.1
    dz = 100.0
    s = ob.OracleSolver(oracle_lut, ke, 3, 3, 10., 10., ob.sundir_from_angles(0., 60.), np.full(ke, dz), diff_solver=0,
                        rtol=1e-10)
    assert s.l1d.all()
    ext = (1.0 / ke) / dz
    k = np.full((ke, 3, 3), 0.5 * ext)
    s.set_optical_properties(albedo, k, k, np.full((ke, 3, 3), 0.5), ldelta_scaling=False)
    assert s.solve(False, True, S0) == 0
    edn, eup, abso, edir = s.get_result()
    s.destroy()
    assert edir[0, 1, 1] == pytest.approx(S0 * mu0, rel=1e-7)                            # float32-cast sun angle
    assert edir[ke, 1, 1] == pytest.approx(S0 * mu0 * np.exp(-1.0 / mu0), rel=1e-6)
    assert edn[0, 1, 1] == 0.0
    assert eup[ke, 1, 1] == pytest.approx(albedo * (edir[ke, 1, 1] + edn[ke, 1, 1]), rel=1e-8)
    # conservation: what enters at the top leaves at the top, is absorbed by the ground or by the layers.  Not to rounding:
    # in 1-D layers calc_flx_div takes the beam's absorption as edir x (1 - exp(-kabs dz / mu0)) (src/pprts.F90:5305-5309)
    # while the Eddington coefficients remove 1 - a33 - a13 - a23 from it; the two differ by a few 1e-4 here.
    absorbed = (abso[:, 1, 1] * dz).sum() + (1 - albedo) * (edir[ke, 1, 1] + edn[ke, 1, 1])
    assert edir[0, 1, 1] - eup[0, 1, 1] == pytest.approx(absorbed, rel=2e-3)


# ---- 24 diffuse streams -----------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def lut324(tmp_path_factory):
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(324)
    paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut324o") / "LUT"))
    return ob.OracleLut(cfg, paths)


@pytest.mark.parametrize("phi0", sorted(CASES))
def test_forward_scattered_light_leaves_downstream_24_streams(lut324, phi0):
    """3_24 (src/boxmc_3_24.inc:36-120, src/pprts.F90:386-390): the side faces carry four angular quadrants each, [inward,
    outward] per quadrant, sides start with an INWARD stream; 0-based 8 + 2 q + (0 towards +x | 1 towards -x) on x faces,
    16 + ... on y faces, quadrants q = 2, 3 point downwards.  The 3_24 permutation (src/optprop.F90:1098-1124) has to send the
    forward peak of the scattered beam into the downward streams that travel with the beam."""
    axis, sign = CASES[phi0]
    s = ob.OracleSolver(lut324, Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(phi0, 60.0), np.full(Nz, DZ), solver_id=324,
                        diff_solver=0, rtol=1e-8)
    z = np.zeros((Nz, Nx, Ny))
    ksca, g = z.copy(), z.copy()
    ksca[K, I, J] = 1.0 / DZ
    g[K, I, J] = 0.85
    s.set_optical_properties(0.0, z, ksca, g)
    assert s.solve(False, True, 1000.0) == 0
    ediff = s.state(2)
    s.destroy()
    base = 8 if axis == "x" else 16
    plus = [base + 4, base + 6]          # downward quadrants, travelling towards + (inward: leave through face index + 1)
    minus = [base + 5, base + 7]         # ... towards - (leave through the cell's own low face)

    def leaving(dofs, towards_plus):
        if axis == "x":
            return sum(ediff[d, K, I + 1 if towards_plus else I, J] for d in dofs)
        return sum(ediff[d, K, I, J + 1 if towards_plus else J] for d in dofs)
    down = leaving(plus, True) if sign > 0 else leaving(minus, False)
    up = leaving(minus, False) if sign > 0 else leaving(plus, True)
    assert down > 3.0 * up > 0, (phi0, down, up)
    # upward quadrants (q = 0, 1) get the isotropic part only, the same on both sides
    up_plus, up_minus = leaving([base + 0, base + 2], True), leaving([base + 1, base + 3], False)
    assert up_plus == pytest.approx(up_minus, rel=1e-6) and up_plus < down


# ---- the ground -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver_id", [310, 810, 324])
def test_ground_reflects_albedo_times_what_arrives(oracle_lut, lut810, lut324, solver_id):
    """the surface rows of the diffuse operator (src/pprts_shell.F90:500-505) and of the solar source (src/pprts.F90:4830-4842):
    per column, what goes up at the ground is albedo x (direct + diffuse down), also with a 2-D albedo and whatever the
    number of streams the upward flux is spread over"""
    lut = {310: oracle_lut, 810: lut810, 324: lut324}[solver_id]
    rng = np.random.default_rng(3)
    nz, nx, ny = 4, 5, 6
    s = ob.OracleSolver(lut, nz, nx, ny, 100., 100., ob.sundir_from_angles(150., 40.), np.full(nz, 100.0),
                        solver_id=solver_id, diff_solver=0, rtol=1e-10)
    kabs = rng.uniform(1e-4, 1e-3, (nz, nx, ny))
    ksca = rng.uniform(1e-3, 1e-2, (nz, nx, ny))
    g = rng.uniform(0.0, 0.5, (nz, nx, ny))
    a2d = rng.uniform(0.05, 0.6, (nx, ny))
    s.set_optical_properties(0.3, kabs, ksca, g, albedo_2d=a2d)
    assert s.solve(False, True, 1000.0) == 0
    edn, eup, abso, edir = s.get_result()
    s.destroy()
    np.testing.assert_allclose(eup[nz], a2d * (edir[nz] + edn[nz]), rtol=1e-7)
    assert edn[0].max() == 0.0 and eup[0].min() > 0.0
