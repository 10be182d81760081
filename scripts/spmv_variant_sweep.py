#!/usr/bin/env python
"""Per-variant SpMV / iteration times of the Jacobi-PCG on an X mesh (default: refined once = the size of a rank's share at
8 GPUs).  python scripts/spmv_variant_sweep.py [level] [variant[:occ]...]   (occ = option spmv_occ, CTAs per SM of the grid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import femocs_b200 as fb

lev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
variants = sys.argv[2:] or ["304", "305", "302", "300", "301", "201", "8"]
nodes, hexs, mk = bench.load_x_mesh(lev)
for v in variants:
    k, occ, split = (v.split(":") + ["6", "32"])[:3]
    k = int(k); occ = int(occ); split = int(split)
    ctx = fb.Context(0)
    ctx.set_option("spmv_kernel", k); ctx.set_option("cg_profile", 64); ctx.set_option("spmv_occ", occ); ctx.set_option("spmv_split", split)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0, cg_tolerance=1e-9, n_cg=10000))
    assert s.import_mesh(nodes, hexs, mk)
    for _ in range(2):
        s.setup(-bench.E0, 0.0); s.assemble(True); it = s.solve()
    sp, vec, ns = s.solve_profile(); ms, it2, _ = s.solve_stats()
    print("level %d variant %d occ %d split %d (ran %d): %d it, SpMV+dot %.1f us, vectors %.1f us (events, %d samples), whole solve %.1f us / iteration"
          % (lev, k, occ, split, s.solve_kernel(), it, 1e3 * sp, 1e3 * vec, ns, 1e3 * ms / max(1, it2)), flush=True)
    ctx.close()
