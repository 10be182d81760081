"""FB_PRECOND_TWOLEVEL (Jacobi + aggregation coarse-grid correction, femocs_b200/csrc/twolevel.cu): the same system to
the same absolute residual as every other preconditioner -- potential within 1e-8 of the CPU oracle (SSOR-CG as the
reference) -- in markedly fewer iterations than Jacobi-PCG, with the reference's return conventions."""
import numpy as np
import pytest

from femocs_b200 import synth
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu
REL = 1e-8


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope="module")
def fb():
    import femocs_b200
    return femocs_b200


@pytest.mark.parametrize("name", ["hemicone", "mdsmall", "mdbig"])
def test_twolevel_matches_oracle(name, fb, golden):
    m = golden("mesh", name)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    o.setup(0.5, 0.0, False); o.assemble(True); assert o.solve(10000, 1e-11, 1.2, 0) > 0
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11))
    assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    c.set_option("cg_persistent", 0)
    s.setup(0.5, 0.0); s.assemble(True); it_j = s.solve()
    s.conf.precond = fb.PRECOND_TWOLEVEL
    c.set_option("tl_agg", 64)                                 # small meshes: ~300-500 aggregates
    s.setup(0.5, 0.0); s.assemble(True); it_t = s.solve()
    assert 0 < it_t < 0.8 * it_j, (it_t, it_j)
    assert _rel(s.export_solution(), o.export_solution()) < REL
    assert s.solve(cg_tolerance=1e-8) == 0                     # warm start from the converged potential
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve(n_cg=5) == -5                               # -#it when the cap is hit (DealSolver.cpp:455-457)
    c.close()


def test_twolevel_dirichlet_anode_and_space_charge(fb, golden):
    """the coarse matrix follows the Dirichlet mask (anode mode changes the constrained set); assemble(false) with new
    particles keeps the matrix and its coarse inverse"""
    m = golden("mesh", "mdsmall"); g = golden("interp", "mdsmall")
    ok = g["pic_ok"]; cf = -180.9512268 * 0.01
    c = fb.Context(0)
    c.set_option("cg_persistent", 0); c.set_option("tl_agg", 64)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11, anode_BC="dirichlet", mode="transient", precond=fb.PRECOND_TWOLEVEL))
    assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    s.setup(0.0, 100.0); s.assemble(True); assert s.solve() > 0
    o.setup(0.0, 100.0, True); o.assemble(True); assert o.solve(10000, 1e-11, 1.2, 0) > 0
    assert _rel(s.export_solution(), o.export_solution()) < REL
    launches = c.kernel_launches
    s.set_particles(g["points"][ok], g["pic_cells"][ok], cf)
    s.assemble(False); assert s.solve() > 0
    o.assemble(False, g["points"][ok], g["pic_cells"][ok], cf); assert o.solve(10000, 1e-11, 1.2, 0) > 0
    assert _rel(s.export_solution(), o.export_solution()) < REL
    assert c.kernel_launches - launches < 5 * 400              # no set-up kernels in between: iterations only
    c.close()


def test_twolevel_on_hbm_sized_system(fb, golden):
    m = golden("mesh", "mdsmall")
    nodes, hexs, mk = synth.refine_vacuum(m["nodes"], m["hexs"], m["hex_markers"], 2)
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-9))
    assert s.import_mesh(nodes, hexs, mk)
    s.setup(0.5, 0.0); s.assemble(True)
    itj = s.solve(); phi = s.export_solution().copy(); ms_j = s.solve_stats()[0]
    s.conf.precond = fb.PRECOND_TWOLEVEL
    s.setup(0.5, 0.0); s.assemble(True)
    itt = s.solve(); ms_t = s.solve_stats()[0]                 # includes the one-off set-up (sort, Galerkin matrix, Cholesky)
    assert 0 < itt < 0.5 * itj and s.solve_kernel() == 306, (itt, itj)
    assert _rel(s.export_solution(), phi) < 1e-7               # both stop at |r| <= 1e-9 with different preconditioners
    s.setup(0.5, 0.0); s.assemble(False)                       # second solve: set-up re-used
    it2 = s.solve(); ms_2 = s.solve_stats()[0]
    assert it2 == itt
    print("hbm-sized: Jacobi %d it %.1f ms | two-level %d it %.1f ms (first, with set-up) %.1f ms (again)" % (itj, ms_j, itt, ms_t, ms_2))
    c.close()
