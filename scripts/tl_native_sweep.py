import sys
sys.path.insert(0,'/root/repo')
import numpy as np, femocs_b200 as fb
for name in ("mdsmall","mdbig","tip110"):
    m=dict(np.load('/root/repo/tests/golden/mesh_%s.npz'%name))
    ctx=fb.Context(0); ctx.set_option("cg_persistent",0)
    s=fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-9)); s.import_mesh(m["nodes"],m["hexs"],m["hex_markers"])
    s.setup(0.5,0.0); s.assemble(True); itj=s.solve()
    s.conf.precond=fb.PRECOND_TWOLEVEL
    out=[]
    for nc in (74,148,296,592,1184):
        ctx.set_option("tl_agg", (s.n_dofs+nc-1)//nc)
        s.setup(0.5,0.0); s.assemble(True); it=s.solve(); out.append((nc,it))
    print(name, s.n_dofs, "Jacobi", itj, "two-level (n_c, it):", out, flush=True)
    ctx.close()
