"""GPU parity of the current / heat solvers on the bulk mesh (src/CurrentHeatSolver.cpp) against the CPU
oracle restatement: boundary classification and face order, the two assembled systems (matrix, right-hand side), the
solutions of the coupled loop of ProjectRunaway::solve_heat (src/ProjectRunaway.cpp:535-571), gradients and limits.
Both sides are driven to the same absolute residual (over-converged), tolerance 1e-8 relative as BASELINE.json states."""
import numpy as np
import pytest

from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

MESHES = ["hemicone", "mdsmall", "mdbig"]
# PhysicalQuantities::hc_resistivity_data is the host code's table; any monotone table exercises the same arithmetic
TAB_T = np.array([200., 250., 273.15, 300., 350., 400., 450., 500., 600., 800., 1000., 1200., 1357.])
TAB_RHO = np.array([10.49, 13.87, 15.43, 17.23, 20.58, 23.95, 27.34, 30.76, 37.72, 52.6, 69.1, 88.2, 104.3])
T_AMB = 300.0


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def emission_like(cen):
    """synthetic per-face emission data peaked at the apex (stands in for EmissionReader::get_current_densities /
    get_nottingham, which come from GETELEC on the host)"""
    z = cen[:, 2]
    J = 3e-4 * np.exp((z - z.max()) / 6.0)            # [A / Ang^2]
    nott = -2e-9 * J / 3e-4 * (1.0 + 0.3 * np.sin(cen[:, 0]))
    return J, nott


@pytest.fixture(scope="module")
def fb():
    import femocs_b200
    return femocs_b200


@pytest.fixture(scope="module")
def ctx(fb):
    c = fb.Context(0)
    yield c
    c.close()


def _pair(fb, ctx, m):
    o = Oracle(); o.import_bulk_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    o.ch_set_physics(TAB_T, TAB_RHO); o.ch_setup(T_AMB)
    s = fb.CurrentHeatSolver(ctx)
    assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.set_dependencies(TAB_T, TAB_RHO); s.setup(T_AMB)
    return s, o


@pytest.mark.parametrize("name", MESHES)
def test_bulk_mesh_and_faces_match_oracle(name, fb, ctx, golden):
    m = golden("mesh", name)
    s, o = _pair(fb, ctx, m)
    assert (s.n_dofs, s.n_cells, s.nnz, s.n_vertices, s.n_bfaces) == (o.n_dofs, o.n_cells, o.nnz, o.n_vertices, o.n_bfaces)
    assert s.n_cells == int((m["hex_markers"] < 0).sum())
    cen = s.export_surface_centroids()
    assert np.array_equal(cen, o.surface_centroids())                # same faces in the same (cell, face) order
    assert len(cen) == s.n_surface_faces
    if "quads" in m:
        assert len(cen) == len(m["quads"])                           # every surface quadrangle is one copper_surface face


@pytest.mark.parametrize("name", MESHES)
def test_current_and_heat_systems_match_oracle(name, fb, ctx, golden):
    m = golden("mesh", name)
    s, o = _pair(fb, ctx, m)
    J, nott = emission_like(o.surface_centroids())
    # --- current: matrix, rhs
    s.current.set_bcs(J); s.current.assemble()
    o.current_assemble(J)
    g = s.get_system(); rp, col, val, _ = o.csr(); rhs = o.vectors()[0]
    assert np.array_equal(g["rowptr"], rp) and np.array_equal(g["col"], col)
    assert np.abs(g["val"] - val).max() <= 1e-12 * np.abs(val).max()
    assert np.abs(g["rhs"] - rhs).max() <= 1e-12 * np.abs(rhs).max()
    it_g = s.current.solve(5000, 1e-14); it_o = o.ch_solve(0, 5000, 1e-14)
    assert it_g > 0 and it_o > 0
    v2d = o.vectors()[2]
    phi_o = o.ch_solution(0)[v2d]
    assert _rel(s.current.export_solution(), phi_o) < 1e-8
    # --- heat: temperature-dependent matrix and Joule source need a non-trivial previous temperature
    rng = np.random.default_rng(5)
    T0 = T_AMB + 400.0 * rng.random(o.n_vertices)
    Td = np.zeros(o.n_dofs); Td[v2d] = T0
    o.ch_set_solution(1, Td); s.heat.import_solution(T0)
    o.ch_set_solution(0, o.ch_solution(0))                           # (oracle keeps its own current potential)
    s.current.import_solution(phi_o)                                 # identical potentials on both sides for the source term
    dt = 4e-13
    s.heat.set_bcs(nott); s.heat.assemble(dt)
    o.heat_assemble(dt, nott)
    g = s.get_system(); _, _, val, _ = o.csr(); rhs = o.vectors()[0]
    assert np.abs(g["val"] - val).max() <= 1e-12 * np.abs(val).max()
    assert np.abs(g["rhs"] - rhs).max() <= 1e-11 * np.abs(rhs).max()
    it_g = s.heat.solve(5000, 1e-16); it_o = o.ch_solve(1, 5000, 1e-16)
    assert it_g > 0 and it_o > 0
    T_o = o.ch_solution(1)[v2d]
    assert _rel(s.heat.export_solution(), T_o) < 1e-8
    assert not s.heat.check_limits(0.0, 1e5)
    assert abs(s.heat.stat_sol_min - T_o.min()) < 1e-8 * T_o.max() and abs(s.heat.stat_sol_max - T_o.max()) < 1e-8 * T_o.max()
    assert s.heat.check_limits(0.0, T_o.max() * 0.5)                 # True = out of limits, as the reference


def test_coupled_loop_matches_oracle(fb, ctx, golden):
    """ProjectRunaway::solve_heat repeated: current.assemble/solve -> heat.assemble(dt)/solve with warm starts, three
    steps; temperatures, current potential and current density (export_temp_rho) against the oracle"""
    m = golden("mesh", "mdsmall")
    s, o = _pair(fb, ctx, m)
    J, nott = emission_like(o.surface_centroids())
    v2d = o.vectors()[2]
    for step in range(3):
        scale = 1.0 + 0.5 * step
        s.current.set_bcs(scale * J); s.current.assemble(); assert s.current.solve(5000, 1e-14) > 0
        o.current_assemble(scale * J); assert o.ch_solve(0, 5000, 1e-14) > 0
        s.heat.set_bcs(scale * nott); s.heat.assemble(2e-13); assert s.heat.solve(5000, 1e-16) >= 0
        o.heat_assemble(2e-13, scale * nott); assert o.ch_solve(1, 5000, 1e-16) >= 0
        assert _rel(s.current.export_solution(), o.ch_solution(0)[v2d]) < 1e-8
        assert _rel(s.heat.export_solution(), o.ch_solution(1)[v2d]) < 1e-8
    temp, rho = s.export_temp_rho()
    o.ch_select(0)
    assert _rel(rho, o.export_solution_grad()) < 1e-8
    assert temp.max() > T_AMB + 1e-3                                 # the tip did heat up


def test_call_order_is_enforced(fb, ctx, golden):
    m = golden("mesh", "hemicone")
    s, o = _pair(fb, ctx, m)
    J, nott = emission_like(o.surface_centroids())
    s.current.set_bcs(J); s.current.assemble()
    with pytest.raises(fb.FemocsB200Error):
        s.heat.solve()                                               # the heat system has not been assembled
    with pytest.raises(fb.FemocsB200Error):
        s.current.set_bcs(J[:-1]); s.current.assemble()              # wrong number of face values
    p = fb.PoissonSolver(ctx)
    with pytest.raises(fb.FemocsB200Error):
        p.setup(0.1)                                                 # the context holds a bulk mesh
