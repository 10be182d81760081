#!/usr/bin/env python
"""GPU sweep of the two-level preconditioner's aggregate size on the X meshes: iterations, cold-step time (set-up
included: the matrix is re-assembled, hence the coarse matrix re-inverted) and warm-matrix time.  python scripts/two_level_sweep.py [levels...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import femocs_b200 as fb

for lev in [int(a) for a in sys.argv[1:]] or [1, 2]:
    nodes, hexs, mk = bench.load_x_mesh(lev)
    ctx = fb.Context(0)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0, cg_tolerance=1e-9, n_cg=10000))
    assert s.import_mesh(nodes, hexs, mk)
    del nodes, hexs, mk
    n = s.n_dofs
    def cold():
        s.setup(-bench.E0, 0.0); s.assemble(True)
        t = time.perf_counter(); it = s.solve(); ctx.synchronize(); return it, time.perf_counter() - t
    it, t = cold(); it, t = cold()
    print("level %d (%d DoF): Jacobi %d it %.3f s" % (lev, n, it, t), flush=True)
    s.conf.precond = fb.PRECOND_TWOLEVEL
    for nc in (1024, 2048, 4096, 8192):
        ctx.set_option("tl_agg", (n + nc - 1) // nc)
        cold(); it, t = cold()
        s.setup(-bench.E0, 0.0); s.assemble(False)          # same matrix: the coarse inverse is kept
        t0 = time.perf_counter(); it2 = s.solve(); ctx.synchronize(); t2 = time.perf_counter() - t0
        print("   n_c ~%d: %d it, cold step %.3f s, same-matrix solve %.3f s (set-up %.3f s)" % (nc, it, t, t2, t - t2), flush=True)
    ctx.close()
