#!/usr/bin/env python
"""Diagnostics on the GPU box: persistent CG per-phase cycles, time vs iteration count, CTA count sweep."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import femocs_b200 as fb
m = bench.load_native()
for ctas in (148, 74):
    ctx = fb.Context(0)
    ctx.set_option("cg_persistent_ctas", ctas)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-9))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    for ncg in (10000, 10000, 50, 100, 150):
        s.setup(0.5, 0.0); s.assemble(True)
        t = time.perf_counter(); it = s.solve(n_cg=ncg); wall = time.perf_counter() - t
        ms, its, _ = s.solve_stats()
        print("ctas %d n_cg %d: it %d solve %.3f ms (wall %.3f ms)" % (ctas, ncg, it, ms, wall * 1e3), flush=True)
    ctx.set_option("cg_debug", 1)
    s.setup(0.5, 0.0); s.assemble(True); s.solve()
    ctx.close()
