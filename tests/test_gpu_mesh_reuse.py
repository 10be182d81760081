"""Mesh hand-off with unchanged topology on the GPU: after nodes have moved, a second import_mesh with the
same connectivity keeps numbering / sparsity / SpMV tables / CG graph and only refreshes the geometry; the field it then
computes must be the one a fresh context computes on the moved mesh, and the one the oracle computes."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from test_host_setup import _interior_jitter

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("persistent", [1, 0], ids=["persistent_cg", "graph_cg"])
def test_reused_import_solves_the_moved_mesh(persistent, golden):
    import femocs_b200 as fb
    m = golden("mesh", "mdsmall")
    nodes, hexs, mk = m["nodes"], m["hexs"], m["hex_markers"]
    moved = _interior_jitter(nodes, 0.02)
    ctx = fb.Context(0)
    ctx.set_option("cg_persistent", -1 if persistent else 0)
    s = fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-11))
    assert s.import_mesh(nodes, hexs, mk) and not ctx.last_import_reused
    s.setup(0.5, 0.0); s.assemble(True); assert s.solve() > 0
    phi0 = s.export_solution().copy()
    assert s.import_mesh(moved, hexs, mk) and ctx.last_import_reused
    with pytest.raises(fb.FemocsB200Error):
        s.assemble(True)                                    # as after any import: setup first
    s.setup(0.5, 0.0); s.assemble(True); assert s.solve() > 0
    phi1 = s.export_solution().copy()
    assert _rel(phi1, phi0) > 1e-6                          # the geometry did change the field
    o = Oracle(); o.import_mesh(moved, hexs, mk); o.setup(0.5, 0.0, False); o.assemble(True); assert o.solve(10000, 1e-11, 1.2, 0) > 0
    assert _rel(phi1, o.export_solution()) < 1e-8
    fresh = fb.Context(0)
    f = fb.PoissonSolver(fresh, fb.FieldConfig(cg_tolerance=1e-11))
    assert f.import_mesh(moved, hexs, mk)
    f.setup(0.5, 0.0); f.assemble(True); assert f.solve() > 0
    assert _rel(phi1, f.export_solution()) < 1e-10
    vol = s.get_cell_volumes()
    assert _rel(vol, f.get_cell_volumes()) < 1e-13
    # the interpolator follows: tables are rebuilt from the new coordinates by initialize()
    it = fb.Interpolator(ctx); it.initialize(dict(m, nodes=moved)); it.extract_solution(s, True)
    ft = fb.Interpolator(fresh); ft.initialize(dict(m, nodes=moved)); ft.extract_solution(f, True)
    assert _rel(it.get_solutions(), ft.get_solutions()) < 1e-9
    # option: never reuse
    ctx.set_option("mesh_reuse", 0)
    assert s.import_mesh(moved, hexs, mk) and not ctx.last_import_reused
    fresh.close(); ctx.close()


def test_reused_bulk_import(golden):
    import femocs_b200 as fb
    from test_gpu_current_heat import TAB_T, TAB_RHO, T_AMB, emission_like
    m = golden("mesh", "mdsmall")
    nodes, hexs, mk = m["nodes"], m["hexs"], m["hex_markers"]
    moved = _interior_jitter(nodes, 0.02)
    ctx = fb.Context(0)
    s = fb.CurrentHeatSolver(ctx)
    assert s.import_mesh(nodes, hexs, mk) and not ctx.last_import_reused
    assert s.import_mesh(moved, hexs, mk) and ctx.last_import_reused
    s.set_dependencies(TAB_T, TAB_RHO); s.setup(T_AMB)
    o = Oracle(); o.import_bulk_mesh(moved, hexs, mk); o.ch_set_physics(TAB_T, TAB_RHO); o.ch_setup(T_AMB)
    cen = s.export_surface_centroids()
    assert np.array_equal(cen, o.surface_centroids())
    J, nott = emission_like(cen)
    s.current.set_bcs(J); s.current.assemble(); assert s.current.solve(5000, 1e-14) > 0
    o.current_assemble(J); assert o.ch_solve(0, 5000, 1e-14) > 0
    v2d = o.vectors()[2]
    assert _rel(s.current.export_solution(), o.ch_solution(0)[v2d]) < 1e-8
    ctx.close()
