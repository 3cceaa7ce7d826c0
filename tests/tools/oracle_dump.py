"""Stand-in for tools/reference_dump/ex_pprts_dump (the real reference): reads a case file, solves it with the CPU oracle
on a given table set, writes the result file.  Lets the harness of tests/test_gpu_vs_reference_dump.py be exercised where the
reference cannot be built.  Test infrastructure (uses oracle/).

    python tests/tools/oracle_dump.py <case file> <result file> <LUT basename> [small|bench|full]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def lut_config(kind, solver_id):
    from tenstream_b200 import lut as L
    return {"small": L.small_test_config, "bench": L.bench_config, "full": lambda sid: L.LutConfig(solver_id=sid)}[kind](solver_id)


def solve_case(case, cfg, paths, rtol=None, atol=None):
    from oracle import binding as ob
    olut = ob.OracleLut(cfg, paths)
    o = ob.OracleSolver(olut, case["Nz"], case["Nx"], case["Ny"], case["dx"], case["dy"], case["sundir"], case["dz1d"],
                        solver_id=case["solver_id"], diff_solver=1, rtol=case["rtol"] if rtol is None else rtol,
                        atol=case["atol"] if atol is None else atol)
    o.set_optical_properties(case["albedo"], case["kabs"], case["ksca"], case["g"], planck=case["planck"])
    rc = o.solve(case["lthermal"], case["lsolar"], case["edirTOA"])
    if rc:
        raise RuntimeError(f"oracle solve -> {rc}")
    return o.get_result()


if __name__ == "__main__":
    import casefile
    fcase, fres, base = sys.argv[1:4]
    kind = sys.argv[4] if len(sys.argv) > 4 else "small"
    case = casefile.read_case(fcase)
    cfg = lut_config(kind, case["solver_id"])
    casefile.write_result(fres, *solve_case(case, cfg, cfg.paths(base)))
