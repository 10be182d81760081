#!/usr/bin/env python
"""CPU prototype (numpy/scipy, NOT product code): Jacobi + aggregation coarse correction as CG preconditioner,
   M^-1 = D^-1 + P (P^T A P)^-1 P^T   with P = piecewise constants over groups of AGG consecutive free rows.
   python scripts/two_level_prototype.py [level] [agg sizes...]"""
import os, sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pcg(A, b, prec, tol=1e-9, maxit=5000):
    x = np.zeros_like(b); g = -b.copy(); z = prec(g); d = -z; gz = g @ z; it = 0
    while it < maxit:
        it += 1
        h = A @ d; al = gz / (d @ h); x += al * d; g += al * h
        if np.linalg.norm(g) <= tol: break
        z = prec(g); gn = g @ z; be = gn / gz; gz = gn; d = be * d - z
    return it, x


def main():
    import bench
    from oracle.oracle import Oracle
    lev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    aggs = [int(a) for a in sys.argv[2:]] or [512, 2048, 4096]
    nodes, hexs, mk = bench.load_x_mesh(lev)
    o = Oracle(); o.import_mesh(nodes, hexs, mk); o.setup(-bench.E0, 0.0, False); o.assemble(True)
    rp, col, val, _ = o.csr(); A = sp.csr_matrix((val, col, rp)); b = o.vectors()[0]; n = A.shape[0]
    dinv = 1.0 / A.diagonal()
    free = np.diff(A.indptr) > 1
    print("system %d DoF, %d nnz, %d free" % (n, A.nnz, free.sum()), flush=True)
    t = time.time(); itj, xj = pcg(A, b, lambda g: dinv * g); print("Jacobi-PCG %d its %.1fs" % (itj, time.time() - t), flush=True)
    # aggregates = runs of AGG dofs along the Morton (Z-order) curve through the dof coordinates: compact boxes
    _, _, v2d, v2n = o.vectors()
    xyz = np.zeros((n, 3)); xyz[v2d] = nodes[v2n]
    q = ((xyz - xyz.min(0)) / (xyz.max(0) - xyz.min(0) + 1e-300) * (2 ** 20 - 1)).astype(np.uint64)
    def spread(v):
        v = v & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    rank = np.empty(n, np.int64); rank[np.argsort(key, kind="stable")] = np.arange(n)
    for AGG in aggs:
        agg = rank // AGG; nc = agg.max() + 1
        P = sp.csr_matrix((free.astype(float), (np.arange(n), agg)), shape=(n, nc))
        Ac = (P.T @ A @ P).tocsc()
        empty = np.asarray(Ac.diagonal() == 0)
        Ac = Ac + sp.diags(empty.astype(float))
        lu = spla.splu(Ac)
        def prec(g):
            return dinv * g + P @ lu.solve(P.T @ g)
        t = time.time(); it, x = pcg(A, b, prec)
        print("AGG %d: nc %d, coarse nnz/row %.1f: %d its (%.1fx fewer) %.1fs, |x-xj| %.2e" % (AGG, nc, Ac.nnz / nc, it, itj / it, time.time() - t, np.abs(x - xj).max()), flush=True)


if __name__ == "__main__":
    main()
