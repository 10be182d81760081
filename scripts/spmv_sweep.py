#!/usr/bin/env python
"""Tuning sweep on the GPU box: SpMV / vector-kernel time per CG iteration on the X mesh for the
SpMV kernel variants and DoF orderings.  Prints one line per point; not a bench number."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import femocs_b200 as fb  # noqa: E402

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kernels = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [100, 300, 301, 302]
orders = [int(k) for k in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
# native mesh: single-launch persistent CG vs CUDA-graph multi-kernel CG
m = bench.load_native()
for pers in ((1, 0) if os.environ.get('SWEEP_NATIVE', '1') == '1' else ()):
    ctx = fb.Context(0)
    ctx.set_option("cg_persistent", pers)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-9))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    for rep in range(4):
        s.setup(0.5, 0.0); s.assemble(True)
        it = s.solve()
        ms, its, _ = s.solve_stats()
    print("native mdbig persistent=%d: %d iterations, solve %.3f ms, %.2f us/iteration" % (pers, it, ms, 1e3 * ms / max(1, its)), flush=True)
    ctx.close()
if levels < 0:
    sys.exit(0)
nodes, hexs, mk = bench.load_x_mesh(levels)
for order in orders:
    ctx = fb.Context(0)
    ctx.set_option("dof_order", order)
    ctx.set_option("cg_profile", 24)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-9, n_cg=40))
    t = time.time(); s.import_mesh(nodes, hexs, mk); t_imp = time.time() - t
    n, nnz = s.n_dofs, s.nnz
    b_spmv = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    b_iter = b_spmv + 88.0 * n
    s.setup(0.5, 0.0); s.assemble(True)
    for k in kernels:
        ctx.set_option("spmv_kernel", k)
        res = []
        for rep in range(2):
            s.setup(0.5, 0.0); s.assemble(True)
            it = s.solve()
            sp, ve, ns = s.solve_profile()
            ms, its, _ = s.solve_stats()
            res.append((sp, ve, ms / max(1, its)))
        sp, ve, per = res[-1]
        print("order %d kernel %3d: spmv %.3f ms (%.0f GB/s, %.3f of 6542) vec %.3f ms  iter %.3f ms (%.3f)  res %.3e import %.1fs"
              % (order, k, sp, b_spmv / sp / 1e6, b_spmv / sp / 1e6 / 6542.1, ve, per, b_iter / per / 1e6 / 6542.1, s.last_residual, t_imp), flush=True)
    ctx.close()
