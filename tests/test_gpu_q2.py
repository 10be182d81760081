"""FE_Q(2) on the GPU (option "fe_degree" = 2, csrc/q2.cu; north star "Q1/Q2 stiffness assembly", BASELINE config 2
"nanotip_big Q2 Laplace solve + surface-atom field interpolation") through the C ABI against the oracle's FE_Q(2)
restatement (tests/test_oracle_q2.py pins that one to an independent derivation).  Bars: numbering / sparsity bit-exact,
matrix and right-hand side 1e-12, potential and field on atoms 1e-8 with both sides solved to the same residual."""
import numpy as np
import pytest

from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu
REL = 1e-8


@pytest.fixture(scope="module")
def fb():
    import femocs_b200
    return femocs_b200


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _oracle(m):
    o = Oracle(); o.set_fe_degree(2)
    o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    return o


@pytest.mark.parametrize("name", ["hemicone", "mdbig"])
@pytest.mark.parametrize("anode", [False, True])
def test_q2_system_and_solution_match_oracle(name, anode, fb, golden):
    m = golden("mesh", name); o = _oracle(m)
    tol = 1e-11
    c = fb.Context(0)
    c.set_option("fe_degree", 2)
    s = fb.PoissonSolver(c, fb.FieldConfig(E0=-0.5, cg_tolerance=tol, anode_BC="dirichlet" if anode else "neumann"))
    assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    assert (s.n_dofs, s.n_cells, s.nnz, s.n_vertices, s.n_bfaces) == (o.n_dofs, o.n_cells, o.nnz, o.n_vertices, o.n_bfaces)
    s.setup(0.5, 7.5); s.assemble(True)
    o.setup(0.5, 7.5, anode); o.assemble(True)
    g = s.get_system()
    rp, col, val, save = o.csr()
    rhs, _, v2d, v2n = o.vectors()
    assert np.array_equal(g["rowptr"], rp) and np.array_equal(g["col"], col)
    assert np.array_equal(g["vertex2dof"], v2d) and np.array_equal(g["vertex2node"], v2n)
    scale = np.abs(save).max()
    assert np.abs(g["val_save"] - save).max() <= 1e-12 * scale
    assert np.abs(g["val"] - val).max() <= 1e-12 * scale
    assert np.abs(g["rhs"] - rhs).max() <= 1e-12 * max(np.abs(rhs).max(), 1e-300)
    it = s.solve()
    assert it > 0 and s.last_residual <= tol
    assert o.solve(20000, tol, 1.2, 0) > 0
    assert _rel(s.export_solution(), o.export_solution()) < REL
    assert not s.check_limits(-1.0, 1e4)
    _, lo, hi = o.check_limits(-1.0, 1e4)
    assert s.stat_sol_min == lo and abs(s.stat_sol_max - hi) <= REL * hi
    c.close()


def test_q2_preconditioners_and_kernels_agree(fb, golden):
    m = golden("mesh", "hemicone"); o = _oracle(m)
    o.setup(0.5, 0.0, False); o.assemble(True)
    assert o.solve(20000, 1e-11, 1.2, 0) > 0
    ref = o.export_solution()
    its = {}
    for precond, kernel in ((fb.PRECOND_JACOBI, -1), (fb.PRECOND_JACOBI, 32), (fb.PRECOND_JACOBI, 100), (fb.PRECOND_CHEBYSHEV, -1), (fb.PRECOND_TWOLEVEL, -1)):
        c = fb.Context(0)
        c.set_option("fe_degree", 2); c.set_option("spmv_kernel", kernel)
        s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11))
        s.conf.precond = precond
        s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        s.setup(0.5, 0.0); s.assemble(True)
        its[(precond, kernel)] = s.solve()
        assert its[(precond, kernel)] > 0
        assert _rel(s.export_solution(), ref) < REL, (precond, kernel)
        c.close()
    assert its[(fb.PRECOND_TWOLEVEL, -1)] < its[(fb.PRECOND_JACOBI, -1)]


def test_q2_field_step_matches_oracle(fb, golden):
    """config 2 with the quadratic element: solve -> extract_solution (vertex dofs, DealSolver.cpp:317-341) -> field on the
    surface atoms; the interpolator half is the FE_Q(1) one and sees only the vertex potentials"""
    m = golden("mesh", "mdbig"); o = _oracle(m); o.interp_initialize(m)
    c = fb.Context(0)
    c.set_option("fe_degree", 2)
    s = fb.PoissonSolver(c, fb.FieldConfig(E0=-0.5, cg_tolerance=1e-11))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > 0
    it = fb.Interpolator(c); it.initialize(m); it.extract_solution(s, True)
    f = fb.FieldReader(it); f.set_preferences(False, 2, 1); f.interpolate(m["surf_atoms"], m["surf_ids"])
    o.setup(0.5, 0.0, False); o.assemble(True); assert o.solve(20000, 1e-11, 1.2, 0) > 0
    nod = o.extract_solution(True)
    cells, sol = o.locate_interpolate(2, 1, m["surf_atoms"])
    assert np.array_equal(f.markers, cells)
    assert _rel(it.get_solutions(), nod) < REL
    assert _rel(f.interpolation, sol) < REL
    # the quadratic element changes the answer (it is not the Q1 system in disguise) ...
    o1 = Oracle(); o1.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    o1.setup(0.5, 0.0, False); o1.assemble(True); o1.solve(20000, 1e-11, 1.2, 0)
    assert 1e-4 < _rel(s.export_solution(), o1.export_solution()) < 0.2
    # ... and the Gauss-point gradient export (a Q1 formula in the reference) is refused loudly
    with pytest.raises(fb.FemocsB200Error):
        s.export_solution_grad()
    # back to FE_Q(1) on the same context
    c.set_option("fe_degree", 1)
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.5, 0.0); s.assemble(True); assert s.solve() > 0
    assert _rel(s.export_solution(), o1.export_solution()) < REL
    c.close()


@pytest.mark.parametrize("with_tables", [True, False])
def test_q2_space_charge_matches_oracle(with_tables, fb, golden):
    """Poisson with the quadratic element: PoissonSolver.cpp:276-296 (the path taken when shape_degree != 1) = scatter of the
    27 shape values at each particle's unit-cell point; right-hand side 1e-11, potential 1e-8; with the interpolator's
    hexahedron table and with the coefficients rebuilt from the cell vertices"""
    m = golden("mesh", "mdsmall"); g = golden("interp", "mdsmall")
    ok = g["pic_ok"]
    pts = g["points"][ok]; cells = g["pic_cells"][ok]
    cf = -180.9512268 * 0.01
    o = _oracle(m); o.interp_initialize(m)
    c = fb.Context(0)
    c.set_option("fe_degree", 2)
    s = fb.PoissonSolver(c, fb.FieldConfig(mode="transient", cg_tolerance=1e-11))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    if with_tables:
        it = fb.Interpolator(c); it.initialize(m)
    s.set_particles(pts, cells, cf)
    s.setup(0.5, 0.0); s.assemble(True)
    o.setup(0.5, 0.0, False); o.assemble(True, pts, cells, cf)
    rhs = o.vectors()[0]
    assert np.abs(s.get_system()["rhs"] - rhs).max() <= 1e-11 * np.abs(rhs).max()
    assert s.solve() > 0 and o.solve(20000, 1e-11, 1.2, 0) > 0
    assert _rel(s.export_solution(), o.export_solution()) < REL
    # the charge changes the potential (the particles were not lost on the way)
    o.setup(0.5, 0.0, False); o.assemble(True); o.solve(20000, 1e-11, 1.2, 0)
    assert _rel(s.export_solution(), o.export_solution()) > 1e-6
    c.close()
