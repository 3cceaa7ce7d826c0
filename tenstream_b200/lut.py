"""optprop LUT files: parameter-space axes, the ``.mmap4`` wire format, and a synthetic table generator.

The real TenStream LUTs (~4.8 GB for 3_10) are downloaded from a web server by the reference
(``misc/download_LUT.sh:23``) and are not available offline, so this module can also *write* tables in
the reference's own on-disk format.  Both the CPU oracle and the CUDA library read those same files.

Reference interfaces restated here (paths relative to the reference tree):

* axes / parameter space: ``src/optprop_parameters.F90:145-245`` (presets), ``src/optprop_base.F90:228-240``
  (LUT_3_10: tau31, w020, aspect_zx23, g6, phi19, theta19), ``:104-131`` (linspace vs preset axes)
* file naming: ``src/optprop_LUT.F90:364-374`` (``gen_lut_basename``), ``:503-518`` and ``:1345-1380``
  (``<basename>_direct_<Nd>_<Nf>.<dims>.ds1000.nc.{Tdir,Sdir}.mmap4``, ``<basename>_diffuse_<Nf>...Sdiff.mmap4``)
* ``.mmap4`` layout: ``src/mmap.F90:63-127`` (writer) and ``:129-203`` (reader): one page (4096 B) of
  ``size_t`` header ``[dtype_size, n_elements, n_bytes, dim1=Ncoeff, dim2=Nlattice, 0...]`` followed by raw
  native-endian fp32, column-major ``(Ncoeff, Nlattice)`` -> the Ncoeff values of a lattice point are contiguous.
* coefficient order inside a lattice point: ``dst*Nsrc + src`` (``src/optprop_LUT.F90:1491-1492``).
* lattice index: first axis fastest (``src/helper_functions.fypp:2419-2437``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

PAGESIZE = 4096  # sysconf(_SC_PAGESIZE) on every platform the reference's LUTs were written on

# ---------------------------------------------------------------------------------------------------------------
# axes (src/optprop_parameters.F90)
# ---------------------------------------------------------------------------------------------------------------
PRESET_TAU31 = np.array([
    1e-10, 3.62266272998e-07, 7.04565803675e-06, 4.47545500233e-05, 0.000172126759821, 0.000495994753047,
    0.00119161313679, 0.00251026980343, 0.00480799264297, 0.00856221891924, 0.0143961482731, 0.0231530284254,
    0.0358868239775, 0.0541358315379, 0.079959118223, 0.11623968405, 0.167882053841, 0.246414427244,
    0.350199325489, 0.502459974196, 0.759082408765, 1.08083180518, 1.5415157991, 2.19832932733, 3.04549626819,
    4.27145477454, 6.16953841432, 9.43719309835, 15.7335501106, 29.5819342206, 100.0], dtype=np.float32)

PRESET_W020 = np.array([
    0.0, 0.152960717624, 0.295085090042, 0.416951893959, 0.521358613652, 0.610087211908, 0.684967634054,
    0.747886390181, 0.800286677013, 0.84336972609, 0.878674797098, 0.906377786525, 0.928097831502,
    0.943463164595, 0.954135786554, 0.963824066888, 0.972632134967, 0.981529289348, 0.990759644674, 0.99999],
    dtype=np.float32)

PRESET_ASPECT23 = np.array([
    0.02, 0.032, 0.042, 0.056, 0.075, 0.1, 0.133, 0.178, 0.237, 0.316, 0.422, 0.562, 0.75, 1., 1.25, 1.562,
    1.953, 2.441, 3.052, 3.815, 4.768, 5.96, 7.451], dtype=np.float32)

PRESET_G6 = np.array([0.0, 0.2424, 0.4137, 0.5717, 0.7144, 0.85], dtype=np.float32)


def linspace_f32(lo: float, hi: float, n: int) -> np.ndarray:
    """``populate_op_dim_vrange`` (src/optprop_base.F90:104-118) with fp32 ``linspace`` (helper_functions.fypp:2735)."""
    lo32, hi32 = np.float32(lo), np.float32(hi)
    if n <= 1:
        return np.array([lo32], dtype=np.float32)
    k = np.arange(n, dtype=np.float32)
    return (lo32 + k * (hi32 - lo32) / np.float32(n - 1)).astype(np.float32)


# solver_id -> (Nd_top, Nd_side, Nf_top, Nf_side)  (src/pprts.F90:256-443, c_wrapper/f2c_solver_ids.h)
SOLVER_STREAMS = {
    310: (1, 1, 2, 4),
    810: (4, 2, 2, 4),
    324: (1, 1, 8, 8),
}


def solver_dofs(solver_id: int) -> tuple[int, int]:
    dt, ds, ft, fs = SOLVER_STREAMS[solver_id]
    return dt + 2 * ds, ft + 2 * fs


@dataclass
class LutConfig:
    """Parameter space of one solver's tables (``t_op_config``, src/optprop_base.F90:43-58)."""
    solver_id: int = 310
    tau: np.ndarray = field(default_factory=lambda: PRESET_TAU31.copy())
    w0: np.ndarray = field(default_factory=lambda: PRESET_W020.copy())
    aspect: np.ndarray = field(default_factory=lambda: PRESET_ASPECT23.copy())
    g: np.ndarray = field(default_factory=lambda: PRESET_G6.copy())
    phi: np.ndarray = field(default_factory=lambda: linspace_f32(0, 90, 19))
    theta: np.ndarray = field(default_factory=lambda: linspace_f32(0, 90, 19))

    def __post_init__(self):
        for n in ("tau", "w0", "aspect", "g", "phi", "theta"):
            setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float32))

    @property
    def Nd(self) -> int:
        return solver_dofs(self.solver_id)[0]

    @property
    def Nf(self) -> int:
        return solver_dofs(self.solver_id)[1]

    @property
    def dir_axes(self):
        return [self.tau, self.w0, self.aspect, self.g, self.phi, self.theta]

    @property
    def diff_axes(self):
        return [self.tau, self.w0, self.aspect, self.g]

    @property
    def dir_shape(self):
        return tuple(len(a) for a in self.dir_axes)

    @property
    def diff_shape(self):
        return tuple(len(a) for a in self.diff_axes)

    # -- file names (src/optprop_LUT.F90:364-374, :503-518) -------------------------------------------------
    def _descr(self, names, axes):
        return "".join(f".{n}{len(a)}" for n, a in zip(names, axes)) + ".ds1000"

    def direct_basename(self, lut_basename: str) -> str:
        ns = ["tau", "w0", "aspect_zx", "g", "phi", "theta"]
        return f"{lut_basename}_direct_{self.Nd}_{self.Nf}{self._descr(ns, self.dir_axes)}.nc"

    def diffuse_basename(self, lut_basename: str) -> str:
        ns = ["tau", "w0", "aspect_zx", "g"]
        return f"{lut_basename}_diffuse_{self.Nf}{self._descr(ns, self.diff_axes)}.nc"

    def paths(self, lut_basename: str) -> dict[str, str]:
        d, s = self.direct_basename(lut_basename), self.diffuse_basename(lut_basename)
        return {"Tdir": d + ".Tdir.mmap4", "Sdir": d + ".Sdir.mmap4", "Sdiff": s + ".Sdiff.mmap4"}


def small_test_config(solver_id: int = 310) -> LutConfig:
    """A reduced parameter space for unit tests: real axis *values* (sub-sampled), so clamping, the strongly
    non-uniform tau spacing and 6-D interpolation are all exercised, but the tables are ~100 kB."""
    return LutConfig(
        solver_id=solver_id,
        tau=PRESET_TAU31[::3].copy(),        # 11 points, 1e-10 .. 100
        w0=PRESET_W020[[0, 4, 9, 14, 19]].copy(),
        aspect=PRESET_ASPECT23[[0, 5, 9, 13, 16, 22]].copy(),
        g=PRESET_G6[[0, 2, 5]].copy(),
        phi=linspace_f32(0, 90, 4),
        theta=linspace_f32(0, 90, 7),
    )


def bench_config(solver_id: int = 310) -> LutConfig:
    """Full-size tau/w0/aspect/g axes of LUT_3_10 (85 560 lattice points per sun angle, as in the reference)
    with a reduced (phi, theta) lattice so the synthetic direct tables stay at ~160 MB instead of 4.8 GB.
    The solve only ever touches the <=4 (phi, theta) blocks that bracket the sun."""
    return LutConfig(solver_id=solver_id, phi=linspace_f32(0, 90, 3), theta=linspace_f32(0, 90, 4))


# ---------------------------------------------------------------------------------------------------------------
# .mmap4 wire format (src/mmap.F90:63-203)
# ---------------------------------------------------------------------------------------------------------------
def write_mmap4(path: str, table: np.ndarray) -> None:
    """``table`` has shape (Nlattice, Ncoeff) in C order == Fortran (Ncoeff, Nlattice) column-major."""
    table = np.ascontiguousarray(table, dtype=np.float32)
    nlat, ncoef = table.shape
    header = np.zeros(PAGESIZE // 8, dtype=np.uint64)
    header[0] = 4
    header[1] = table.size
    header[2] = table.size * 4
    header[3] = ncoef
    header[4] = nlat
    tmp = path + f".tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(header.tobytes())
        f.write(table.tobytes())
    os.replace(tmp, path)


def read_mmap4(path: str) -> np.ndarray:
    """Memory-map a table; returns a read-only (Nlattice, Ncoeff) fp32 view (``load_mmap_array_2d``)."""
    header = np.fromfile(path, dtype=np.uint64, count=PAGESIZE // 8)
    if header[0] != 4:
        raise ValueError(f"{path}: dtype size {header[0]} != 4 (irealLUT is real32, src/data_parameters.F90:94)")
    shape = [int(v) for v in header[3:] if v != 0]
    if len(shape) != 2:
        raise ValueError(f"{path}: expected 2 dimensions, header says {shape}")
    ncoef, nlat = shape
    if ncoef * nlat != int(header[1]) or int(header[2]) != 4 * ncoef * nlat:
        raise ValueError(f"{path}: inconsistent header")
    return np.memmap(path, dtype=np.float32, mode="r", offset=PAGESIZE, shape=(nlat, ncoef))


# ---------------------------------------------------------------------------------------------------------------
# synthetic tables
# ---------------------------------------------------------------------------------------------------------------
# A small deterministic ray quadrature through the unit box (x,y in [0,1], z in [0,aspect], z pointing DOWN):
#   unscattered part  : sum over rays of exp(-tau_z * path/aspect), binned by the stream the ray leaves in
#   scattered part    : w0 * (1 - unscattered), re-emitted as (1-g)*isotropic-by-face-area + g*forward
#   absorbed part     : (1-w0) * (1 - unscattered)   -> sum over dst of T+S <= 1 for every src, as the real tables.
# The sample sets are mirror-symmetric in x and y, so the tables obey the E/W and N/S symmetries the reference's
# stream permutations assume (src/optprop.F90:1009-1045).  Values are smooth in all six parameters.

def _diffuse_streams(solver_id):
    """Returns a list of streams; each is dict(face, s, ...) describing the open set of directions:
    face 'z' with s=+1 is E_dn (enters through the top, leaves through the bottom).  z points DOWN here.
    10-stream layouts: side streams carry ``sz`` (the vertical half-space).  24-stream layout
    (src/boxmc_3_24.inc:36-120): every face direction is split into four angular quadrants ``q`` = the signs of the two
    transverse direction components -- (x, y) on top/bottom faces, (y, z) on x faces, (x, z) on y faces."""
    _, _, ft, fs = SOLVER_STREAMS[solver_id]
    if (ft, fs) == (2, 4):
        return [
            dict(face="z", s=-1),            # 0 E_up      (is_inward = F)
            dict(face="z", s=+1),            # 1 E_dn      (T)
            dict(face="x", s=-1, sz=+1),     # 2 down, -x  (F)
            dict(face="x", s=+1, sz=+1),     # 3 down, +x  (T)
            dict(face="x", s=-1, sz=-1),     # 4 up,   -x  (F)
            dict(face="x", s=+1, sz=-1),     # 5 up,   +x  (T)
            dict(face="y", s=-1, sz=+1),     # 6 down, -y  (F)
            dict(face="y", s=+1, sz=+1),     # 7 down, +y  (T)
            dict(face="y", s=-1, sz=-1),     # 8 up,   -y  (F)
            dict(face="y", s=+1, sz=-1),     # 9 up,   +y  (T)
        ]
    if (ft, fs) == (8, 8):
        out = []
        for q in ((-1, +1), (+1, +1), (-1, -1), (+1, -1)):        # top/bottom: (x, y) quadrants 1/2, 3/4, 5/6, 7/8
            out += [dict(face="z", s=-1, q=q), dict(face="z", s=+1, q=q)]          # [up (F), down (T)]
        for face in ("x", "y"):                                    # sides: (transverse, z) quadrants, "up" first
            for q in ((-1, -1), (+1, -1), (-1, +1), (+1, +1)):     # z points down: up = -1
                out += [dict(face=face, s=+1, q=q), dict(face=face, s=-1, q=q)]    # [inward (T), outward (F)]
        return out
    raise NotImplementedError(f"no synthetic diffuse table for a {ft}+2*{fs} stream layout")


def _stream_dirs(st):
    """Directions (unit vectors, weights) representing the angular distribution of a diffuse stream."""
    out = []
    if st["face"] == "z":
        for mu, wmu in ((0.9, 0.45), (0.55, 0.55)):
            sin = np.sqrt(1 - mu * mu)
            if "q" in st:
                for az in (22.5, 67.5):
                    a = np.deg2rad(az)
                    out.append(((st["q"][0] * sin * np.cos(a), st["q"][1] * sin * np.sin(a), st["s"] * mu), wmu / 2))
            else:
                for az in (45., 135., 225., 315.):
                    a = np.deg2rad(az)
                    out.append(((sin * np.cos(a), sin * np.sin(a), st["s"] * mu), wmu / 4))
    else:
        # flux through a side face, restricted to one vertical half-space (and one transverse sign for quadrants)
        for un, wn in ((0.85, 0.5), (0.5, 0.5)):          # component along the face normal
            rest = np.sqrt(1 - un * un)
            for ct in (0.35, 0.85):                        # share of the remainder that is vertical
                sz = st["q"][1] if "q" in st else st["sz"]
                signs = (float(st["q"][0]),) if "q" in st else (-1.0, 1.0)   # transverse sign: mirror-symmetric
                for sgn in signs:
                    uz = sz * rest * ct
                    ut = sgn * rest * np.sqrt(1 - ct * ct)
                    d = (st["s"] * un, ut, uz) if st["face"] == "x" else (ut, st["s"] * un, uz)
                    out.append((d, wn / (2 * len(signs))))
    return out


def _entry_points(face, s, aspect, n=5, sub=None):
    """n x n midpoints on the face a stream enters through. ``sub`` = (lo0,hi0,lo1,hi1) restricts to a part."""
    m = (np.arange(n) + 0.5) / n
    lo0, hi0, lo1, hi1 = sub if sub is not None else (0., 1., 0., 1.)
    p, q = np.meshgrid(lo0 + m * (hi0 - lo0), lo1 + m * (hi1 - lo1), indexing="ij")
    p, q = p.ravel(), q.ravel()
    if face == "z":
        z = np.zeros_like(p) if s > 0 else np.full_like(p, aspect)
        return np.stack([p, q, z], 1)
    if face == "x":
        x = np.zeros_like(p) if s > 0 else np.ones_like(p)
        return np.stack([x, p, q * aspect], 1)
    y = np.zeros_like(p) if s > 0 else np.ones_like(p)
    return np.stack([p, y, q * aspect], 1)


def _trace(points, d, aspect):
    """Straight rays through the box: returns a list of hits (ray index, exit axis, sign, path length, share).
    A ray leaving exactly through an edge or corner is shared equally between the faces it touches, which keeps
    the tables exactly mirror-symmetric (an argmin tie-break would depend on rounding of the entry points)."""
    d = np.asarray(d, dtype=np.float64)
    hi = np.array([1.0, 1.0, aspect])
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d > 0, (hi - points) / d, np.where(d < 0, (0.0 - points) / d, np.inf))
    t = np.where(t < 0, np.inf, t)
    tmin = t.min(axis=1, keepdims=True)
    tied = t <= tmin * (1 + 1e-9) + 1e-12
    share = tied / tied.sum(axis=1, keepdims=True)
    hits = []
    for n in range(len(points)):
        for ax in range(3):
            if tied[n, ax]:
                hits.append((n, ax, float(np.sign(d[ax])), float(tmin[n, 0]), float(share[n, ax])))
    return hits


def _classify_diffuse(ax, sgn, d, streams):
    """Index of the diffuse stream a ray leaving through (axis, sign) with direction d joins."""
    sg = lambda v: 1 if v >= 0 else -1
    for k, st in enumerate(streams):
        if "xyz"[ax] != st["face"] or st["s"] != sgn:
            continue
        if "q" in st:
            tr = (d[0], d[1]) if st["face"] == "z" else (d[1], d[2]) if st["face"] == "x" else (d[0], d[2])
            if (sg(tr[0]), sg(tr[1])) == tuple(st["q"]):
                return k
        elif st["face"] == "z" or st["sz"] == sg(d[2]):
            return k
    raise RuntimeError("unclassified ray")


def _iso_weights(streams, aspect):
    """Isotropic redistribution by face area, shared equally between the streams of one face direction."""
    nshare = {}
    for st in streams:
        nshare[(st["face"], st["s"])] = nshare.get((st["face"], st["s"]), 0) + 1
    w = np.array([(1.0 if st["face"] == "z" else aspect) / nshare[(st["face"], st["s"])] for st in streams])
    return w / w.sum()


def _direct_streams(solver_id):
    """Direct streams in the canonical orientation (beam travels +x, +y, down): entry face + sub-face."""
    dt, ds, _, _ = SOLVER_STREAMS[solver_id]
    if (dt, ds) == (1, 1):
        return [dict(face="z", sub=None), dict(face="x", sub=None), dict(face="y", sub=None)]
    if (dt, ds) == (4, 2):
        h = 0.5
        tops = [dict(face="z", sub=(0, h, 0, h)), dict(face="z", sub=(h, 1, 0, h)),
                dict(face="z", sub=(0, h, h, 1)), dict(face="z", sub=(h, 1, h, 1))]
        # side faces: (transverse, vertical) halves; "lower" half = larger z (z points down)
        xs = [dict(face="x", sub=(0, 1, h, 1)), dict(face="x", sub=(0, 1, 0, h))]
        ys = [dict(face="y", sub=(0, 1, h, 1)), dict(face="y", sub=(0, 1, 0, h))]
        return tops + xs + ys
    raise NotImplementedError


def _classify_direct(p, ax, dstreams, aspect):
    """Index of the direct stream whose (sub-)face contains the exit point p on face `ax`."""
    face = "xyz"[ax]
    for k, st in enumerate(dstreams):
        if st["face"] != face:
            continue
        if st["sub"] is None:
            return k
        if face == "z":
            a, b = p[0], p[1]
        elif face == "x":
            a, b = p[1], p[2] / aspect
        else:
            a, b = p[0], p[2] / aspect
        lo0, hi0, lo1, hi1 = st["sub"]
        if lo0 - 1e-9 <= a <= hi0 + 1e-9 and lo1 - 1e-9 <= b <= hi1 + 1e-9:
            return k
    raise RuntimeError("unclassified direct ray")


def synth_diffuse_table(cfg: LutConfig) -> np.ndarray:
    """(Nlattice, Nf*Nf) fp32, coefficient index dst*Nf+src, lattice index tau fastest."""
    streams = _diffuse_streams(cfg.solver_id)
    Nf = len(streams)
    tau = cfg.tau.astype(np.float64)
    w0 = cfg.w0.astype(np.float64)
    g = cfg.g.astype(np.float64)
    out = np.zeros((len(g), len(cfg.aspect), len(w0), len(tau), Nf, Nf))  # [g, a, w0, tau, dst, src]
    for ia, a in enumerate(cfg.aspect.astype(np.float64)):
        iso = _iso_weights(streams, a)
        for s, st in enumerate(streams):
            unsc = np.zeros((len(tau), Nf))
            fwd = np.zeros(Nf)
            pts = _entry_points(st["face"], st["s"], a)
            for d, wd in _stream_dirs(st):
                wray = wd / len(pts)
                for n, ax, sgn, path, share in _trace(pts, d, a):
                    dst = _classify_diffuse(ax, sgn, d, streams)
                    unsc[:, dst] += wray * share * np.exp(-tau * (path / a))       # tau_z * path/dz
                    fwd[dst] += wray * share
            sc = 1.0 - unsc.sum(1)                             # extinguished fraction (tau)
            # coef[g,w0,tau,dst] = unsc + w0*sc*((1-g)*iso + g*fwd)
            redis = (1 - g)[:, None] * iso[None, :] + g[:, None] * fwd[None, :]      # [g, dst]
            out[:, ia, :, :, :, s] = unsc[None, None, :, :] + \
                w0[None, :, None, None] * sc[None, None, :, None] * redis[:, None, None, :]
    return out.reshape(-1, Nf * Nf).astype(np.float32)


def synth_direct_tables(cfg: LutConfig) -> tuple[np.ndarray, np.ndarray]:
    """(Tdir, Sdir): (Nlattice, Nd*Nd) and (Nlattice, Nd*Nf) fp32; lattice index tau fastest, theta slowest."""
    fstreams = _diffuse_streams(cfg.solver_id)
    dstreams = _direct_streams(cfg.solver_id)
    Nf, Nd = len(fstreams), len(dstreams)
    tau = cfg.tau.astype(np.float64)
    w0 = cfg.w0.astype(np.float64)
    g = cfg.g.astype(np.float64)
    asp = cfg.aspect.astype(np.float64)
    nth, nph = len(cfg.theta), len(cfg.phi)
    T = np.zeros((nth, nph, len(g), len(asp), len(w0), len(tau), Nd, Nd))
    S = np.zeros((nth, nph, len(g), len(asp), len(w0), len(tau), Nf, Nd))
    for it, th in enumerate(cfg.theta.astype(np.float64)):
        mu = max(np.cos(np.deg2rad(th)), 1e-4)
        sn = np.sqrt(max(0.0, 1 - mu * mu))
        for ip, ph in enumerate(cfg.phi.astype(np.float64)):
            # canonical orientation: beam towards +x, +y (sym_phi=0 -> along y, 90 -> along x), downwards
            d = np.array([sn * np.sin(np.deg2rad(ph)), sn * np.cos(np.deg2rad(ph)), mu])
            d = np.where(np.abs(d) < 1e-9, 1e-9, d)
            d = d / np.linalg.norm(d)
            for ia, a in enumerate(asp):
                iso = _iso_weights(fstreams, a)
                for s, st in enumerate(dstreams):
                    pts = _entry_points(st["face"], +1, a, n=6, sub=st["sub"])
                    wray = 1.0 / len(pts)
                    unsc = np.zeros((len(tau), Nd))
                    fwd = np.zeros(Nf)
                    for n, ax, sgn, path, share in _trace(pts, d, a):
                        pout = pts[n] + path * d
                        ddst = _classify_direct(pout, ax, dstreams, a)
                        fdst = _classify_diffuse(ax, sgn, d, fstreams)
                        unsc[:, ddst] += wray * share * np.exp(-tau * (path / a))
                        fwd[fdst] += wray * share
                    sc = 1.0 - unsc.sum(1)
                    redis = (1 - g)[:, None] * iso[None, :] + g[:, None] * fwd[None, :]   # [g, dst]
                    T[it, ip, :, ia, :, :, :, s] = unsc[None, None, :, :]
                    S[it, ip, :, ia, :, :, :, s] = \
                        w0[None, :, None, None] * sc[None, None, :, None] * redis[:, None, None, :]
    return (T.reshape(-1, Nd * Nd).astype(np.float32), S.reshape(-1, Nf * Nd).astype(np.float32))


def write_synthetic_lut(cfg: LutConfig, lut_basename: str, force: bool = False) -> dict[str, str]:
    """Write the three tables of ``cfg`` next to ``lut_basename`` (e.g. '/tmp/luts/LUT'); returns their paths."""
    paths = cfg.paths(lut_basename)
    os.makedirs(os.path.dirname(os.path.abspath(paths["Tdir"])), exist_ok=True)
    if force or not os.path.exists(paths["Sdiff"]):
        write_mmap4(paths["Sdiff"], synth_diffuse_table(cfg))
    if force or not (os.path.exists(paths["Tdir"]) and os.path.exists(paths["Sdir"])):
        T, S = synth_direct_tables(cfg)
        write_mmap4(paths["Tdir"], T)
        write_mmap4(paths["Sdir"], S)
    return paths
