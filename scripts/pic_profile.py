#!/usr/bin/env python
"""GPU-box diagnostics of the PIC step (config 3): run under
   ncu --metrics gpu__time_duration.sum --clock-control none   for the per-kernel launch list."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import femocs_b200 as fb
with np.load(os.path.join(bench.ROOT, "tests", "golden", "mesh_mdsmall.npz")) as z:
    m = {k: z[k] for k in z.files}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pos, vel, cells = bench.pic_particles(m, n)
ctx = fb.Context(0)
s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0, cg_tolerance=1e-9, mode="transient"))
s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
it = fb.Interpolator(ctx); it.initialize(m)
lo = m["nodes"].min(0); hi = m["nodes"].max(0)
box = np.array([lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]])
d_pos, d_vel, d_cell = torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), torch.from_numpy(cells).cuda()
s.setup(-bench.E0, 0.0); s.assemble(True); s.solve()
for step in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    lost = C.c_long(0)
    ctx.check(ctx.L.fb_pic_update_positions_dev(ctx.h, n, d_pos.data_ptr(), d_vel.data_ptr(), d_cell.data_ptr(), 0.5, box.ctypes.data, 1, C.byref(lost)))
    n -= lost.value
    s.assemble_dev(False, d_pos.data_ptr(), d_cell.data_ptr(), n, bench.Q_OVER_EPS0 * bench.WSP * 1e-3)
    its = s.solve(); s.check_limits(-1e30, 1e30)
    it.extract_solution(s, True)
    ctx.check(ctx.L.fb_pic_update_velocities_dev(ctx.h, n, d_pos.data_ptr(), d_cell.data_ptr(), d_vel.data_ptr(), 0.5, -17.5882))
    ctx.synchronize()
    print("step", step, "alive", n, "lost", lost.value, "cg", its, flush=True)
ctx.close()
