#!/usr/bin/env python
"""CPU prototype for DESIGN section 10 item 1 (fewer Krylov iterations): BiCGStab iteration counts of the diffuse system
of a C2-like cloud field with (a) no preconditioner, (b) a column-block solve (the GPU path's preconditioner keeps
exactly the couplings inside one column of cells; here: the block of A with source and destination dof in the same
column index, a close stand-in), (c) the column solve followed by a Galerkin coarse-grid correction on a x a
horizontal aggregates (per stream and level).

The operator is taken from the oracle's shell MatMult (checker code, hence this script lives under tests/) by probing
with 10 x 27 coloured vectors: the stencil couples a dof only with dofs whose (k, i, j) differ by at most one.

    [AGGREGATES=2,4,8] python tests/tools/coarse_grid_prototype.py [N=24] [Nz=20]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import binding as ob  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
NZ = int(sys.argv[2]) if len(sys.argv) > 2 else 20
assert N % 3 == 0


def probe_operator(o, Nf, Nl, Nx, Ny):
    """A as CSR, unknowns ordered (dof, level, x, y) in C order"""
    shape = (Nf, Nl, Nx, Ny)
    n = int(np.prod(shape))
    idx = np.arange(n).reshape(shape)
    kk, ii, jj = np.meshgrid(np.arange(Nl), np.arange(Nx), np.arange(Ny), indexing="ij")
    rows, cols, vals = [], [], []
    for s in range(Nf):
        for a in range(3):
            for b in range(3):
                for c in range(3):
                    x = np.zeros(shape)
                    x[s, a::3, b::3, c::3] = 1.0
                    y = o.mat_mult_ediff(x)
                    for dk in (-1, 0, 1):
                        ks = kk + dk
                        for di in (-1, 0, 1):
                            isrc = (ii + di) % Nx
                            for dj in (-1, 0, 1):
                                jsrc = (jj + dj) % Ny
                                m = (ks >= 0) & (ks < Nl) & (ks % 3 == a) & (isrc % 3 == b) & (jsrc % 3 == c)
                                if not m.any():
                                    continue
                                for d in range(Nf):
                                    v = y[d][m]
                                    nz = v != 0
                                    if nz.any():
                                        rows.append(idx[d][m][nz])
                                        cols.append(idx[s][np.clip(ks, 0, Nl - 1)[m], isrc[m], jsrc[m]][nz])
                                        vals.append(v[nz])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return A


def bicgstab_iters(A, b, M=None, rtol=1e-5):
    it = [0]

    def cb(_):
        it[0] += 1
    x, info = spla.bicgstab(A, b, rtol=rtol, atol=0.0, maxiter=2000, M=M, callback=cb)
    res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
    return it[0], info, res


def main():
    cfg, paths = bench.ensure_lut(0)
    lut = ob.OracleLut(cfg, paths)
    sc, bands = bench.make_workload(N, N, NZ)
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    print(f"domain {Nx}x{Ny}x{Nz}, dx {sc['dx']}", flush=True)
    for band in bands:
        o = ob.OracleSolver(lut, Nz, Nx, Ny, sc["dx"], sc["dy"], ob.sundir_from_angles(sc["phi0"], sc["theta0"]), sc["dz1d"],
                            diff_solver=1, coeff_fp32=True)
        o.set_optical_properties(band["albedo"], band["kabs"], band["ksca"], band["g"], planck=band["planck"],
                                 planck_srfc=band["planck_srfc"])
        assert o.solve(not band["lsolar"], band["lsolar"], band["edirTOA"], band["uid"]) == 0
        Nf, Nl = o.Nf, Nz + 1
        b = o.state(4, band["uid"]).reshape(-1)
        t0 = time.time()
        A = probe_operator(o, Nf, Nl, Nx, Ny)
        x = np.random.default_rng(0).standard_normal(A.shape[0])
        err = np.abs(A @ x - o.mat_mult_ediff(x.reshape(Nf, Nl, Nx, Ny)).reshape(-1)).max()
        print(f"{'solar' if band['lsolar'] else 'thermal'}: A {A.shape[0]} x {A.shape[0]}, nnz {A.nnz}, probe check {err:.1e}, "
              f"{time.time() - t0:.0f} s", flush=True)
        o.destroy()
        n = A.shape[0]
        col_of = np.broadcast_to((np.arange(Nx)[:, None] * Ny + np.arange(Ny)[None, :])[None, None], (Nf, Nl, Nx, Ny)).reshape(-1)
        Acoo = A.tocoo()
        keep = col_of[Acoo.row] == col_of[Acoo.col]
        Mcol = sp.csc_matrix((Acoo.data[keep], (Acoo.row[keep], Acoo.col[keep])), shape=(n, n))
        perm = np.argsort(col_of, kind="stable")                 # column-major ordering makes the LU block diagonal
        lu_col = spla.splu(Mcol[perm][:, perm].tocsc(), permc_spec="NATURAL")
        inv_perm = np.empty_like(perm); inv_perm[perm] = np.arange(n)

        def col_solve(r):
            return lu_col.solve(r[perm])[inv_perm]

        print("  no preconditioner      :", bicgstab_iters(A, b), flush=True)
        print("  column block           :", bicgstab_iters(A, b, spla.LinearOperator((n, n), matvec=col_solve)), flush=True)
        for a in [int(v) for v in os.environ.get("AGGREGATES", "4,8").split(",")]:   # 2 is costly beyond 24 x 24
            if Nx % a or Ny % a:
                continue
            agg = np.broadcast_to(((np.arange(Nx)[:, None] // a) * (Ny // a) + np.arange(Ny)[None, :] // a)[None, None],
                                  (Nf, Nl, Nx, Ny))
            dl = np.arange(Nf * Nl).reshape(Nf, Nl, 1, 1)
            cidx = (dl * ((Nx // a) * (Ny // a)) + agg).reshape(-1)
            nc = Nf * Nl * (Nx // a) * (Ny // a)
            P = sp.csr_matrix((np.ones(n), (np.arange(n), cidx)), shape=(n, nc))
            lu_c = spla.splu((P.T @ A @ P).tocsc())

            def two_level(r, lu_c=lu_c, P=P):
                z = col_solve(r)
                return z + P @ lu_c.solve(P.T @ (r - A @ z))

            def coarse_first(r, lu_c=lu_c, P=P):
                z = P @ lu_c.solve(P.T @ r)
                return z + col_solve(r - A @ z)
            print(f"  column + coarse {a}x{a}    :", bicgstab_iters(A, b, spla.LinearOperator((n, n), matvec=two_level)),
                  "| coarse first:", bicgstab_iters(A, b, spla.LinearOperator((n, n), matvec=coarse_first)), flush=True)

            def additive(r, lu_c=lu_c, P=P):     # no fine-level MatMult inside the preconditioner
                return col_solve(r) + P @ lu_c.solve(P.T @ r)
            print(f"  additive {a}x{a}           :", bicgstab_iters(A, b, spla.LinearOperator((n, n), matvec=additive)), flush=True)
            # inexact coarse solves: m sweeps of the coarse-level column relaxation (the coarse analogue of col_solve)
            Ac = (P.T @ A @ P).tocsr()
            ccol = cidx.reshape(Nf, Nl, Nx, Ny)[0, 0] if False else None
            aggc = np.broadcast_to((np.arange(Nx // a)[:, None] * (Ny // a) + np.arange(Ny // a)[None, :])[None, None],
                                   (Nf, Nl, Nx // a, Ny // a)).reshape(-1)
            Acoo_c = Ac.tocoo()
            keepc = aggc[Acoo_c.row] == aggc[Acoo_c.col]
            Mc = sp.csc_matrix((Acoo_c.data[keepc], (Acoo_c.row[keepc], Acoo_c.col[keepc])), shape=Ac.shape)
            permc = np.argsort(aggc, kind="stable")
            lu_cc = spla.splu(Mc[permc][:, permc].tocsc(), permc_spec="NATURAL")
            inv_permc = np.empty_like(permc); inv_permc[permc] = np.arange(nc)

            def ccol_solve(rc):
                return lu_cc.solve(rc[permc])[inv_permc]
            for m in [int(v) for v in os.environ.get("COARSE_SWEEPS", "1,2,4,8").split(",")]:
                def two_level_inexact(r, m=m, P=P, Ac=Ac):
                    z = col_solve(r)
                    rc = P.T @ (r - A @ z)
                    xc = np.zeros(nc)
                    for _ in range(m):
                        xc = xc + ccol_solve(rc - Ac @ xc)
                    return z + P @ xc

                def additive_inexact(r, m=m, P=P, Ac=Ac):
                    rc = P.T @ r
                    xc = np.zeros(nc)
                    for _ in range(m):
                        xc = xc + ccol_solve(rc - Ac @ xc)
                    return col_solve(r) + P @ xc
                print(f"  {a}x{a}, {m} coarse column sweeps: multiplicative",
                      bicgstab_iters(A, b, spla.LinearOperator((n, n), matvec=two_level_inexact)), "| additive",
                      bicgstab_iters(A, b, spla.LinearOperator((n, n), matvec=additive_inexact)), flush=True)


if __name__ == "__main__":
    main()
