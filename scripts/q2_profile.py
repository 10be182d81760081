#!/usr/bin/env python
"""One FE_Q(2) field step on nanotip_big with a capped CG (for `ncu --metrics gpu__time_duration.sum`: per-kernel times of
k_q2_stiffness, k_q2_neumann, the value permutation and the CG kernels).  python scripts/q2_profile.py [n_cg]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femocs_b200 as fb

m = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mesh_mdbig.npz")))
c = fb.Context(0)
c.set_option("fe_degree", 2)
s = fb.PoissonSolver(c, fb.FieldConfig(E0=-0.5, cg_tolerance=1e-9, n_cg=int(sys.argv[1]) if len(sys.argv) > 1 else 20))
assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
for _ in range(2):
    s.setup(0.5, 0.0); s.assemble(True); it = s.solve()
print("FE_Q(2): %d DoF, nnz %d, %d iterations (capped), kernel %d" % (s.n_dofs, s.nnz, it, s.solve_kernel()))
