"""The C-ABI library loads and exports every symbol include/pprts_b200.h declares (no compute calls, no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from tenstream_b200 import build
    return build.build()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pprts_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pprts_(?:f2c|b200)_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built_lib):
    L = ctypes.CDLL(built_lib)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/pprts_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from tenstream_b200 import pprts
    assert set(header_symbols()) == set(pprts.ABI_SYMBOLS)


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the handle API refuses to create a solver; nothing routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from tenstream_b200 import lut as L
    from tenstream_b200 import pprts as P
    import tempfile
    cfg = L.small_test_config(310)
    paths = L.write_synthetic_lut(cfg, os.path.join(tempfile.mkdtemp(), "LUT"))
    lut = P.Lut(cfg, paths)
    with pytest.raises(P.PprtsError, match="no CUDA device"):
        P.init_pprts(0, 4, 3, 3, 100., 100., P.sundir_from_angles(0., 20.), dz1d=np.full(4, 50.), lut=lut)
    lut.close()
    src = open(os.path.join(ROOT, "tenstream_b200", "pprts.py")).read()
    assert "oracle" not in src.replace("CPU oracle", "")


def test_product_does_not_import_oracle():
    import glob
    for f in glob.glob(os.path.join(ROOT, "tenstream_b200", "**", "*"), recursive=True):
        if f.endswith((".py", ".cu", ".cuh", ".h")):
            txt = open(f).read()
            assert "from oracle" not in txt and "import oracle" not in txt and "pprts_oracle" not in txt, f
    for f in glob.glob(os.path.join(ROOT, "tools", "**", "*"), recursive=True) + glob.glob(os.path.join(ROOT, "fortran", "*")):
        if os.path.isdir(f):
            continue
        txt = open(f).read()                # dev tools and the Fortran shim stay off the checker as well
        assert "from oracle" not in txt and "import oracle" not in txt and "pprts_oracle" not in txt, f


def test_host_polyfit_matches_oracle(built_lib):
    """polyfit of the error-growth tracking (src/adaptive_spectral_integration.F90:174-257) is plain host code in the
    library: checked on the CPU against the oracle's restatement and numpy"""
    import numpy as np
    from oracle import binding as ob
    from tenstream_b200 import pprts as P
    rng = np.random.default_rng(2)
    for d in (1, 2):
        x = np.sort(rng.uniform(0, 10, 4))
        y = rng.uniform(-1, 1, 4)
        c, ierr = P.polyfit(x, y, d)
        co, ierro = ob.polyfit(x, y, d)
        assert ierr == ierro == 0
        assert np.allclose(c, co, rtol=1e-9, atol=1e-12)
        assert np.allclose(c, np.polyfit(x, y, d)[::-1], rtol=1e-7, atol=1e-9)
    c, ierr = P.polyfit(np.array([0., 1., 1., 2.]), np.array([0., 1., 0., 1.]), 1)
    assert ierr == 1                                   # equal abscissae (:198-207)


def test_host_find_iter_and_kernelwidth_matches_oracle(built_lib):
    """find_iter_and_kernelwidth of smooth_surface_fluxes (src/pprts_postprocess.F90:86-128) is plain host code in the
    library: checked on the CPU against the oracle's restatement, including the reference's error for radii that
    would need more than 250 iterations"""
    import pytest
    from oracle import binding as ob
    from tenstream_b200 import pprts as P
    for radius, dx, dy, xm, ym in [(0., 100., 100., 8, 8), (120., 100., 100., 8, 6), (300., 100., 50., 32, 40),
                                   (1000., 100., 100., 256, 256), (2500., 25., 25., 1024, 1024), (40., 100., 100., 5, 5)]:
        assert P.find_iter_and_kernelwidth(radius, dx, dy, xm, ym) == ob.find_iter_and_kernelwidth(radius, dx, dy, xm, ym)
    with pytest.raises(P.PprtsError, match="smoothing iteration count"):
        P.find_iter_and_kernelwidth(1e5, 100., 100., 4, 4)
    with pytest.raises(ValueError):
        ob.find_iter_and_kernelwidth(1e5, 100., 100., 4, 4)


def test_header_is_c99_and_links_from_c(built_lib, tmp_path):
    """include/pprts_b200.h is the boundary a C host model compiles against: it must be plain C (no C++ / torch types)
    and a C program must link against the shared library and call it (no device needed for the config defaults)"""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "t.c"
    src.write_text('#include "pprts_b200.h"\n'
                   'int main(void) { pprts_b200_config c; pprts_b200_default_config(&c);\n'
                   '  return (c.solver_id == 310 && c.keep_solutions == 1 && c.geometric_coeffs == 0) ? 0 : 1; }\n')
    libdir = os.path.dirname(built_lib)
    exe = tmp_path / "t"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-L", libdir, "-lpprts_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


REF_C_WRAPPER = "/root/reference/c_wrapper"


@pytest.mark.skipif(not os.path.isdir(REF_C_WRAPPER), reason="the reference checkout is not on this machine")
@pytest.mark.parametrize("client", ["pprts.c", "opp.c", "set_global_optprop.c"])
def test_reference_c_clients_link_unchanged(built_lib, tmp_path, client):
    """The drop-in claim of include/pprts_b200.h, layer 1: the reference's own C programs (c_wrapper/pprts.c, the host-model
    example, and c_wrapper/opp.c, the coefficient probe) are compiled from where they lie, with the reference's own
    headers (f2c_pprts.h, f2c_solver_ids.h), and linked against libpprts_b200.so instead of libf2c_pprts -- every symbol they
    use must resolve.  MPI is not in this image: a four-function stub mpi.h stands in for it (the wrapper only passes the
    communicator handle through).  Linked, not run (no device here)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    stub = tmp_path / "mpistub"
    stub.mkdir()
    (stub / "mpi.h").write_text(
        "#ifndef MPI_STUB_H\n#define MPI_STUB_H\ntypedef int MPI_Comm;\n#define MPI_COMM_WORLD 0\n"
        "static inline int MPI_Init(int *c, char ***v) { (void)c; (void)v; return 0; }\n"
        "static inline int MPI_Finalize(void) { return 0; }\n"
        "static inline int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = 1; return 0; }\n"
        "static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }\n"
        "static inline int MPI_Comm_c2f(MPI_Comm c) { return c; }\n#endif\n")
    libdir = os.path.dirname(built_lib)
    exe = tmp_path / client.replace(".c", "")
    subprocess.check_call([cc, "-std=gnu99", "-I", str(stub), "-I", REF_C_WRAPPER, os.path.join(REF_C_WRAPPER, client),
                           "-L", libdir, "-lpprts_b200", f"-Wl,-rpath,{libdir}", "-Wl,--no-undefined", "-lm", "-o", str(exe)])
    undefined = subprocess.run(["nm", "-u", str(exe)], capture_output=True, text=True, check=True).stdout
    used = sorted({ln.split()[-1] for ln in undefined.splitlines() if "pprts_f2c" in ln})
    assert used, "the client calls no pprts_f2c_* symbol?"
    exported = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    for sym in used:
        assert f" T {sym}" in exported, sym


@pytest.mark.skipif(not os.path.isdir(REF_C_WRAPPER), reason="the reference checkout is not on this machine")
def test_prototypes_agree_with_the_reference_header(tmp_path):
    """One translation unit that sees the reference's c_wrapper/f2c_pprts.h AND include/pprts_b200.h: C rejects
    incompatible redeclarations, so every pprts_f2c_* prototype has to agree in all argument types.  The one deliberate
    difference is taken out: pprts_f2c_destroy takes its flag BY VALUE here, as the reference's implementation
    (c_wrapper/f2c_pprts.F90:458-459) and its own caller (pprts.c:99, `pprts_f2c_destroy(0)`) do, while the reference's
    header says int* (f2c_pprts.h:65)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "both.c"
    src.write_text('#define pprts_f2c_destroy pprts_f2c_destroy_as_declared_by_the_reference\n'
                   '#include <f2c_pprts.h>\n#undef pprts_f2c_destroy\n#include "pprts_b200.h"\nint main(void) { return 0; }\n')
    subprocess.check_call([cc, "-std=gnu99", "-Werror", "-I", REF_C_WRAPPER, "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "both.o")])


# ---- the three statements of the boundary's structs must stay in step ----------------------------------------------------------
def _c_struct_fields(name):
    """[(field, kind, count)] of a `typedef struct { ... } name;` in include/pprts_b200.h; kind in {int, double, ptr}"""
    import re
    txt = open(os.path.join(ROOT, "include", "pprts_b200.h")).read()
    end = re.search(r"\}\s*" + name + r"\s*;", txt).start()
    body = txt[txt.rfind("typedef struct {", 0, end) + len("typedef struct {"):end]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const )?(int|double|float|long)\b(.*)", decl)
        assert m, decl
        base = m.group(2)
        for item in m.group(3).split(","):
            item = item.strip()
            ptr = item.startswith("*")
            item = item.lstrip("* ").replace("const ", "").strip()
            am = re.match(r"(\w+)\[(\d+)\]", item)
            fname, count = (am.group(1), int(am.group(2))) if am else (item, 1)
            out.append((fname, "ptr" if ptr else base, count))
    return out


def _fortran_type_fields(name):
    import re
    txt = open(os.path.join(ROOT, "fortran", "m_pprts_b200.F90")).read()
    body = re.search(r"type, bind\(c\) :: " + name + r"\n(.*?)\n\s*end type", txt, re.S | re.I).group(1)
    body = re.sub(r"&\s*\n\s*&?", " ", body)
    out = []
    for line in body.splitlines():
        line = line.split("!")[0].strip()
        if not line:
            continue
        typ, names = line.split("::")
        typ = typ.strip().lower().replace(" ", "")
        kind = {"integer(c_int)": "int", "integer(c_long)": "long", "real(c_double)": "double", "type(c_ptr)": "ptr"}[typ]
        for item in re.findall(r"\w+(?:\(\d+\))?", names):
            am = re.match(r"(\w+)\((\d+)\)", item)
            out.append((am.group(1), kind, int(am.group(2))) if am else (item, kind, 1))
    return out


def _ctypes_fields(struct):
    import ctypes as C
    out = []
    for fname, typ in struct._fields_:
        count = 1
        if hasattr(typ, "_length_"):
            count, typ = typ._length_, typ._type_
        kind = {C.c_int: "int", C.c_double: "double", C.c_long: "long"}.get(typ, "ptr")
        out.append((fname, kind, count))
    return out


def test_config_struct_is_the_same_in_c_fortran_and_python():
    """pprts_b200_config crosses the boundary by reference: include/pprts_b200.h, the bind(C) type of the Fortran shim
    and the ctypes mirror have to list the same fields, in the same order, with the same types -- a field added to one of
    them only would shift everything behind it without any error"""
    from tenstream_b200 import pprts as P
    c = _c_struct_fields("pprts_b200_config")
    low = lambda fields: [(n.lower(), k, cnt) for n, k, cnt in fields]      # noqa: E731  (Fortran is case-insensitive)
    assert len(c) > 30
    assert low(_fortran_type_fields("c_config")) == low(c)
    assert _ctypes_fields(P._Config) == c
    bands = _c_struct_fields("pprts_b200_bands")
    assert _ctypes_fields(P._Bands) == bands
    assert low(_fortran_type_fields("c_bands")) == low(bands)


def test_fortran_interfaces_match_the_header():
    """every bind(C) interface of fortran/m_pprts_b200.F90 names a function include/pprts_b200.h declares and takes as many
    arguments as its prototype (the shim cannot be compiled in this image, so at least its binding table is held to the
    header)"""
    import re
    hdr = open(os.path.join(ROOT, "include", "pprts_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(pprts_b200_\w+)\s*\(([^;{]*?)\)\s*;", hdr, re.S):
        args = " ".join(m.group(2).split())
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    f90 = open(os.path.join(ROOT, "fortran", "m_pprts_b200.F90")).read()
    f90 = re.sub(r"&\s*\n\s*&?", " ", f90)
    found = re.findall(r"function\s+\w+\s*\(([^)]*)\)\s*bind\(c,\s*name='(\w+)'\)", f90, re.I)
    assert len(found) >= 20
    for args, sym in found:
        assert sym in protos, f"{sym} is bound in the Fortran shim but not declared in the header"
        n = len([a for a in args.split(",") if a.strip()])
        assert n == protos[sym], (sym, n, protos[sym])
