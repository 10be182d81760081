#!/usr/bin/env python
"""CPU prototype (scipy, NOT product code): how many CG iterations would an aggregation-based algebraic multigrid
preconditioner need on the benchmark's base mesh, against the Jacobi-PCG the GPU path runs today?  Evidence for
DESIGN.md section 7, item 6.  Plain (unsmoothed) aggregation + one damped-Jacobi pre/post smoothing per level,
Galerkin coarse operators, V-cycle; everything an HBM-bound GPU implementation would also do (SpMV, axpy, no sweeps).

    python scripts/amg_prototype.py [levels_of_refinement]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def aggregate(A, theta=0.08):
    """greedy aggregation on the strength graph |a_ij| >= theta sqrt(a_ii a_jj); returns the aggregate id of every row"""
    n = A.shape[0]
    d = A.diagonal()
    C = A.tocoo()
    strong = (C.row != C.col) & (np.abs(C.data) >= theta * np.sqrt(np.abs(d[C.row] * d[C.col])))
    S = sp.csr_matrix((np.ones(strong.sum()), (C.row[strong], C.col[strong])), shape=(n, n))
    agg = np.full(n, -1, np.int64)
    n_agg = 0
    indptr, indices = S.indptr, S.indices
    for i in range(n):                                   # pass 1: roots whose strong neighbours are all free
        if agg[i] >= 0:
            continue
        nb = indices[indptr[i]:indptr[i + 1]]
        if len(nb) and np.all(agg[nb] < 0):
            agg[i] = n_agg; agg[nb] = n_agg; n_agg += 1
    for i in range(n):                                   # pass 2: attach the rest to a neighbouring aggregate
        if agg[i] >= 0:
            continue
        nb = indices[indptr[i]:indptr[i + 1]]
        nb = nb[agg[nb] >= 0]
        if len(nb):
            agg[i] = agg[nb[0]]
        else:
            agg[i] = n_agg; n_agg += 1
    return agg, n_agg


class AMG:
    def __init__(self, A, max_levels=8, coarse=2000, omega=0.67):
        self.levels = []
        self.omega = omega
        self.over = 1.8
        while len(self.levels) < max_levels - 1 and A.shape[0] > coarse:
            agg, na = aggregate(A)
            P = sp.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], na))
            dinv = 1.0 / A.diagonal()
            self.levels.append((A, P, dinv))
            A = (P.T @ A @ P).tocsr()
        self.coarse = spla.splu(A.tocsc())
        self.sizes = [l[0].shape[0] for l in self.levels] + [A.shape[0]]
        self.nnz = [l[0].nnz for l in self.levels] + [A.nnz]

    def vcycle(self, r, lvl=0):
        if lvl == len(self.levels):
            return self.coarse.solve(r)
        A, P, dinv = self.levels[lvl]
        x = self.omega * dinv * r                                        # pre-smoothing from a zero guess
        rc = P.T @ (r - A @ x)
        x += self.over * (P @ self.vcycle(rc, lvl + 1))                  # plain aggregation: over-correction
        x += self.omega * dinv * (r - A @ x)                             # post-smoothing
        return x


def pcg(A, b, prec, tol=1e-9, maxit=20000):
    x = np.zeros_like(b); g = -b.copy(); h = prec(g); d = -h; gh = g @ h; it = 0
    while np.linalg.norm(g) > tol and it < maxit:
        it += 1
        h = A @ d; al = gh / (d @ h); x += al * d; g += al * h
        if np.linalg.norm(g) <= tol:
            break
        h = prec(g); gn = g @ h; be = gn / gh; gh = gn; d = be * d - h
    return it, x


def main():
    lev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    nodes, hexs, mk = bench.load_x_mesh(lev)
    o = Oracle(); o.import_mesh(nodes, hexs, mk)
    o.setup(-bench.E0, 0.0, False); o.assemble(True)
    rp, col, val, _ = o.csr()
    A = sp.csr_matrix((val, col, rp)); b = o.vectors()[0]
    n = A.shape[0]
    print("system: %d DoF, %d nnz" % (n, A.nnz), flush=True)
    dinv = 1.0 / A.diagonal()
    t = time.time(); itj, xj = pcg(A, b, lambda g: dinv * g); tj = time.time() - t
    print("Jacobi-PCG: %d iterations (%.1f s)" % (itj, tj), flush=True)
    # Dirichlet rows are diagonal after the elimination: Jacobi solves them exactly, the hierarchy is built on the rest
    A.eliminate_zeros()                                   # the elimination leaves explicit zeros in the pattern
    free = np.diff(A.indptr) > 1
    Af = A[free][:, free].tocsr()
    t = time.time(); M = AMG(Af); ts = time.time() - t
    work = sum(M.nnz) / A.nnz
    print("AMG setup %.1f s: %d free DoF, level sizes %s, operator complexity %.2f" % (ts, free.sum(), M.sizes, work), flush=True)

    def prec(g):
        z = dinv * g
        z[free] = M.vcycle(g[free])
        return z
    t = time.time(); ita, xa = pcg(A, b, prec); ta = time.time() - t
    # a V-cycle costs ~ (3 SpMV-equivalents) x operator complexity on top of the CG's own SpMV
    spmv_equiv = ita * (1 + 3 * work)
    print("AMG-PCG: %d iterations (%.1f s), ~%.0f fine-grid SpMV equivalents vs %d for Jacobi -> %.1fx fewer; |x - x_jacobi| = %.2e"
          % (ita, ta, spmv_equiv, itj, itj / spmv_equiv, np.abs(xa - xj).max()), flush=True)


if __name__ == "__main__":
    main()
