#!/usr/bin/env python
"""CPU prototype (numpy/scipy, NOT product code) of the multigrid preconditioner the library builds on the device
(csrc/amg.cu): repeated pairwise "handshake" matching on the most negative coupling (aggregates of <= 2^passes rows),
Galerkin operators of the piecewise-constant prolongator (a sum over aggregate pairs), K-cycle (two flexible-CG steps
per coarse level) or V-cycle, Chebyshev/Jacobi smoothing, flexible outer CG.  It answers, on the CPU and per mesh,
how many iterations / fine-grid SpMV equivalents each variant needs against plain Jacobi-PCG.

    python scripts/amg_kcycle_prototype.py [refinement_level] [passes] [rounds]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def handshake(A, rounds=6):
    """one pairwise matching pass; returns aggregate ids (pairs and left-over singletons), numbered by lower row"""
    n = A.shape[0]
    indptr, indices, data = A.indptr, A.indices, A.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    mate = np.full(n, -1, np.int64)
    offd = (indices != rows) & (data < 0)
    for _ in range(rounds):
        free = mate < 0
        ok = offd & free[rows] & free[indices]
        w = np.where(ok, -data, -1.0)
        # segment arg-max with the lowest column winning ties
        best = np.full(n, -1, np.int64)
        nz = np.diff(indptr) > 0
        mx = np.full(n, -1.0)
        mx[nz] = np.maximum.reduceat(w, indptr[:-1][nz])
        hit = ok & (w == mx[rows])
        idx = np.flatnonzero(hit)
        # first hit per row (columns are sorted)
        r = rows[idx]
        first = np.ones(len(idx), bool); first[1:] = r[1:] != r[:-1]
        best[r[first]] = indices[idx[first]]
        i = np.flatnonzero(best >= 0)
        mutual = i[best[best[i]] == i]
        if len(mutual) == 0:
            break
        mate[mutual] = best[mutual]
    lead = (mate < 0) | (np.arange(n) < mate)
    agg = np.empty(n, np.int64)
    agg[lead] = np.arange(lead.sum())
    agg[~lead] = agg[mate[~lead]]
    return agg, int(lead.sum())


def coarsen(A, passes=3, rounds=6):
    n = A.shape[0]
    agg = np.arange(n)
    Ac = A
    for _ in range(passes):
        a, na = handshake(Ac, rounds)
        P = sp.csr_matrix((np.ones(Ac.shape[0]), (np.arange(Ac.shape[0]), a)), shape=(Ac.shape[0], na))
        Ac = (P.T @ Ac @ P).tocsr(); Ac.sort_indices()
        agg = a[agg]
    return agg, Ac


def lam_max(A, dinv, it=15):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(A.shape[0])
    for _ in range(it):
        v /= np.linalg.norm(v)
        w = dinv * (A @ v)
        lam = v @ w
        v = w
    return 1.1 * np.abs(lam)


class AMG:
    def __init__(self, A, passes=3, rounds=6, coarse=1500, max_levels=12, cycle="K", smoother="cheb2"):
        self.levels = []
        self.cycle = cycle
        self.smoother = smoother
        self.spmv = 0.0   # fine-grid SpMV equivalents spent (by nnz)
        self.nnz0 = A.nnz
        while len(self.levels) < max_levels - 1 and A.shape[0] > coarse:
            agg, Ac = coarsen(A, passes, rounds)
            dinv = 1.0 / A.diagonal()
            lm = lam_max(A, dinv)
            self.levels.append((A, agg, dinv, lm, Ac.shape[0]))
            A = Ac
        self.coarse = spla.splu(A.tocsc())
        self.sizes = [l[0].shape[0] for l in self.levels] + [A.shape[0]]
        self.nnzs = [l[0].nnz for l in self.levels] + [A.nnz]

    def mv(self, A, x):
        self.spmv += A.nnz / self.nnz0
        return A @ x

    def smooth0(self, A, dinv, lm, r):
        """smoothing from a zero initial guess"""
        if self.smoother == "jac":
            return (1.0 / lm * 1.5) * dinv * r if False else (4.0 / (3.0 * lm)) * dinv * r
        # Chebyshev degree 2 on [lm/4, lm]
        a, b = lm / 4.0, lm
        th, de = (a + b) / 2, (b - a) / 2
        s1 = th / de
        x = dinv * r / th
        rho = 1.0 / s1
        rho1 = 1.0 / (2 * s1 - rho)
        res = r - self.mv(A, x)
        dlt = rho1 * rho * (x) + 2 * rho1 / de * (dinv * res)
        return x + dlt

    def smooth(self, A, dinv, lm, x, r):
        """smoothing of A x = r starting from x"""
        if self.smoother == "jac":
            return x + (4.0 / (3.0 * lm)) * dinv * (r - self.mv(A, x))
        a, b = lm / 4.0, lm
        th, de = (a + b) / 2, (b - a) / 2
        s1 = th / de
        res = r - self.mv(A, x)
        d = dinv * res / th
        x = x + d
        rho = 1.0 / s1
        rho1 = 1.0 / (2 * s1 - rho)
        res = res - self.mv(A, d)
        d = rho1 * rho * d + 2 * rho1 / de * (dinv * res)
        return x + d

    def cyc(self, r, lvl=0):
        if lvl == len(self.levels):
            return self.coarse.solve(r)
        A, agg, dinv, lm, nc = self.levels[lvl]
        x = self.smooth0(A, dinv, lm, r)
        res = r - self.mv(A, x)
        rc = np.bincount(agg, weights=res, minlength=nc)
        if self.cycle == "V" or lvl + 1 == len(self.levels):
            ec = self.cyc(rc, lvl + 1)
        else:
            Ac = self.levels[lvl + 1][0]
            c1 = self.cyc(rc, lvl + 1); v1 = self.mv(Ac, c1)
            rho1 = c1 @ v1; al1 = c1 @ rc
            rt = rc - (al1 / rho1) * v1
            if self.cycle == "K" and np.linalg.norm(rt) <= 0.25 * np.linalg.norm(rc) and False:
                ec = (al1 / rho1) * c1
            else:
                c2 = self.cyc(rt, lvl + 1); v2 = self.mv(Ac, c2)
                gam = c2 @ v1; bet = c2 @ v2; al2 = c2 @ rt
                rho2 = bet - gam * gam / rho1
                ec = (al1 / rho1 - gam * al2 / (rho1 * rho2)) * c1 + (al2 / rho2) * c2
        x = x + ec[agg]
        return self.smooth(A, dinv, lm, x, r)


def fcg(A, b, prec, tol=1e-9, maxit=3000):
    """flexible CG (Polak-Ribiere beta), cold start"""
    x = np.zeros_like(b); g = b.copy(); z = prec(g); d = z.copy(); gz = g @ z; it = 0
    while np.linalg.norm(g) > tol and it < maxit:
        it += 1
        h = A @ d; al = gz / (d @ h); x += al * d
        gold = g.copy(); g = g - al * h
        if np.linalg.norm(g) <= tol:
            break
        z = prec(g); gn = g @ z; be = (gn - z @ gold) / gz; gz = gn; d = z + be * d
    return it, x


def main():
    import bench
    from oracle.oracle import Oracle
    lev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    nodes, hexs, mk = bench.load_x_mesh(lev)
    o = Oracle(); o.import_mesh(nodes, hexs, mk)
    o.setup(-bench.E0, 0.0, False); o.assemble(True)
    rp, col, val, _ = o.csr()
    A = sp.csr_matrix((val, col, rp)); b = o.vectors()[0]
    n = A.shape[0]
    print("system: %d DoF, %d nnz" % (n, A.nnz), flush=True)
    dinv = 1.0 / A.diagonal()
    t = time.time(); itj, xj = fcg(A, b, lambda g: dinv * g); tj = time.time() - t
    print("Jacobi-PCG: %d iterations (%.1f s)" % (itj, tj), flush=True)
    A.eliminate_zeros()
    free = np.diff(A.indptr) > 1
    Af = A[free][:, free].tocsr(); Af.sort_indices()
    for cycle, sm in (("K", "cheb2"), ("K", "jac"), ("V", "cheb2")):
        t = time.time(); M = AMG(Af, passes, rounds, cycle=cycle, smoother=sm); ts = time.time() - t
        print("%s-cycle/%s setup %.1f s: sizes %s, nnz/row %s, operator complexity %.2f" % (
            cycle, sm, ts, M.sizes, ["%.1f" % (a / b_) for a, b_ in zip(M.nnzs, M.sizes)], sum(M.nnzs) / Af.nnz), flush=True)

        def prec(g):
            z = dinv * g
            z[free] = M.cyc(g[free])
            return z
        M.spmv = 0.0
        t = time.time(); ita, xa = fcg(A, b, prec); ta = time.time() - t
        tot = ita + M.spmv * Af.nnz / A.nnz
        print("   AMG-FCG: %d iterations (%.1f s), %.0f fine SpMV equivalents vs %d Jacobi -> %.1fx fewer; |x - x_jacobi| = %.2e"
              % (ita, ta, tot, itj, itj / tot, np.abs(xa - xj).max()), flush=True)


if __name__ == "__main__":
    main()
