#!/usr/bin/env python
"""GPU-box diagnostics of the native (nanotip_big) field step: run under
   ncu --metrics gpu__time_duration.sum --clock-control none  for the per-kernel launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import femocs_b200 as fb
m = bench.load_native()
ctx = fb.Context(0)
s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0, cg_tolerance=1e-9))
s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
interp = fb.Interpolator(ctx); interp.initialize(m)
atoms = np.ascontiguousarray(m["surf_atoms"])
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    s.setup(-bench.E0, 0.0); s.assemble(True)
    it = s.solve()
    interp.extract_solution(s, True)
    f = fb.FieldReader(interp); f.set_preferences(False, 2, 1); f.interpolate(atoms)
    f3 = fb.FieldReader(interp); f3.set_preferences(False, 3, 1); f3.interpolate(atoms)
print("iterations", it, "solve ms", s.solve_stats()[0])
ctx.close()
