import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, bench, femocs_b200 as fb
for opt in (1, 0):
    nodes, hexs, mk = bench.load_x_mesh(2)
    ctx = fb.Context(0); ctx.set_option("asm_map", opt)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=bench.E0))
    assert s.import_mesh(nodes, hexs, mk); del nodes, hexs, mk
    for k in range(4):
        s.setup(-bench.E0, 0.0)
        t = time.perf_counter(); s.assemble(True); dt = time.perf_counter() - t
        print("asm_map", opt, "assemble(True) #%d: %.1f ms" % (k, 1e3 * dt), flush=True)
    if opt == 1:
        g1 = s.get_system()["val_save"][::1000].copy()
    else:
        g0 = s.get_system()["val_save"][::1000].copy()
    ctx.close()
print("max rel diff mapped vs row-walk:", np.abs(g1 - g0).max() / np.abs(g0).max())
