#!/usr/bin/env python
"""GPU-box diagnostics: wall-clock of every ABI call of the re-meshed native MD step (host arrays in, atoms out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import femocs_b200 as fb
m = bench.load_native()
ctx = fb.Context(0)
atoms = np.ascontiguousarray(m["surf_atoms"])
interp = fb.Interpolator(ctx)
for rep in range(4):
    T = []
    def lap(name, t0):
        ctx.synchronize(); T.append((name, 1e3 * (time.perf_counter() - t0)))
    t = time.perf_counter(); s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0, cg_tolerance=1e-9)); s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); lap("import_mesh", t)
    t = time.perf_counter(); s.setup(-bench.E0, 0.0); s.assemble(True); lap("setup+assemble", t)
    t = time.perf_counter(); it = s.solve(); lap("solve", t)
    t = time.perf_counter(); interp.initialize(m); lap("interp_initialize", t)
    t = time.perf_counter(); interp.extract_solution(s, True); lap("extract", t)
    t = time.perf_counter(); f = fb.FieldReader(interp); f.set_preferences(False, 2, 1); f.interpolate(atoms); lap("locate_interpolate", t)
print("  ".join("%s %.2f" % kv for kv in T), " total %.2f ms" % sum(v for _, v in T))
ctx.close()
