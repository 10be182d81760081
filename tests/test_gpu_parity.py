"""GPU parity tests: the CUDA path (through the C ABI, via femocs_b200.solver) against
  * the golden vectors produced by the reference's own code (tests/golden, bit-exact cells),
  * the CPU oracle on the same seeded inputs,
  * size-independent properties at large sizes.
Tolerances: cell indices bit-exact; potential / field 1e-8 relative (BASELINE.json north star),
with both solves driven to the same absolute residual tolerance."""
import numpy as np
import pytest
import scipy.sparse as sp

from femocs_b200 import synth
from oracle.fields import hash_field
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

MESHES = ["hemicone", "mdsmall", "mdbig"]
REL = 1e-8


@pytest.fixture(scope="module")
def fb():
    import femocs_b200
    return femocs_b200


@pytest.fixture(scope="module")
def ctx(fb):
    c = fb.Context(0)
    yield c
    c.close()


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _oracle(m):
    o = Oracle()
    o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    o.interp_initialize(m)
    return o


@pytest.fixture(scope="module")
def oracles(golden):
    return {n: _oracle(golden("mesh", n)) for n in MESHES}


# ------------------------------------------------------------------------------------------
# solver half vs the oracle
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MESHES)
def test_assembled_system_matches_oracle(name, fb, ctx, golden, oracles):
    m = golden("mesh", name); o = oracles[name]
    E0 = -0.3
    s = fb.PoissonSolver(ctx)
    assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    assert (s.n_dofs, s.n_cells, s.nnz, s.n_vertices, s.n_bfaces) == (o.n_dofs, o.n_cells, o.nnz, o.n_vertices, o.n_bfaces)
    s.setup(-E0, 0.0); s.assemble(True)
    o.setup(-E0, 0.0, False); o.assemble(True)
    g = s.get_system()
    rp, col, val, save = o.csr()
    rhs, _, v2d, v2n = o.vectors()
    assert np.array_equal(g["rowptr"], rp) and np.array_equal(g["col"], col)        # same numbering & pattern
    assert np.array_equal(g["vertex2dof"], v2d) and np.array_equal(g["vertex2node"], v2n)
    scale = np.abs(save).max()
    assert np.abs(g["val_save"] - save).max() <= 1e-12 * scale
    assert np.abs(g["val"] - val).max() <= 1e-12 * scale
    assert np.abs(g["rhs"] - rhs).max() <= 1e-12 * np.abs(rhs).max()
    vol = s.get_cell_volumes()
    assert _rel(vol[::97], [o.cell_vol(k) for k in range(0, o.n_cells, 97)]) < 1e-12
    assert vol.min() > 0


@pytest.mark.parametrize("name", MESHES)
def test_solution_matches_oracle(name, fb, ctx, golden, oracles):
    m = golden("mesh", name); o = oracles[name]
    E0 = -0.5
    tol = 1e-11
    s = fb.PoissonSolver(ctx, fb.FieldConfig(E0=E0, cg_tolerance=tol))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(-E0, 0.0); s.assemble(True)
    ncg = s.solve()
    assert ncg > 0 and s.last_residual <= tol
    o.setup(-E0, 0.0, False); o.assemble(True)
    assert o.solve(10000, tol, 1.2, 0) > 0                          # the reference's SSOR-CG
    assert _rel(s.export_solution(), o.export_solution()) < REL
    # same algorithm on the CPU (Jacobi-CG): iteration counts agree closely
    o.setup(-E0, 0.0, False); o.assemble(True)
    it_cpu = o.solve(10000, tol, 1.2, 1)
    assert abs(it_cpu - ncg) <= max(3, it_cpu // 20)
    # independent residual check of the GPU solution on the host
    g = s.get_system()
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]))
    # (the recurrence residual that CG monitors drifts from the true one by O(eps |A| |x|))
    drift = 1e-13 * abs(A).sum(1).max() * np.linalg.norm(g["sol"])
    assert np.linalg.norm(A @ g["sol"] - g["rhs"]) <= 2 * tol + drift
    assert not s.check_limits(-1.0, 1e4)
    _, lo, hi = o.check_limits(-1.0, 1e4)
    assert s.stat_sol_min == lo == 0.0 and abs(s.stat_sol_max - hi) <= REL * hi
    # warm start: already converged -> 0 iterations (deal.II SolverControl); the restart recomputes
    # the TRUE residual, which sits at the drift level above, hence the looser tolerance here
    assert s.solve(cg_tolerance=1e-8) == 0
    # iteration cap -> negative count (DealSolver.cpp:455-457)
    s.setup(-E0, 0.0); s.assemble(True)
    assert s.solve(n_cg=5) == -5


@pytest.mark.parametrize("kernel", [0, 100, 101, 102, 103, 200, 201, 202, 203, 204, 300, 301, 302, 303, 304, 305, 306, 307, 308, 310, 311, 2, 8, 32])
def test_spmv_kernel_variants_agree(kernel, fb, golden, oracles):
    """every SpMV kernel of the multi-kernel CG (windowed / plain row-block streaming variants, lanes-per-row
    variants) gives the oracle's solution"""
    m = golden("mesh", "mdsmall"); o = oracles["mdsmall"]
    c = fb.Context(0)
    c.set_option("cg_persistent", 0)
    c.set_option("spmv_kernel", kernel)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > 0
    if kernel >= 310 or kernel in (306, 307, 308):
        assert s.solve_kernel() == kernel       # no silent fall-back to the full-matrix layout
    o.setup(0.5, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
    assert _rel(s.export_solution(), o.export_solution()) < REL
    if kernel in (306, 307, 308):
        # segmented layout: other caps (chains of up to 6 segments on this mesh), and under the other preconditioners
        for cap, precond in ((12, fb.PRECOND_JACOBI), (20, fb.PRECOND_CHEBYSHEV), (32, fb.PRECOND_TWOLEVEL)):
            c.set_option("spmv_split", cap); s.conf.precond = precond
            s.setup(0.5, 0.0); s.assemble(True)
            assert s.solve() > 0 and s.solve_kernel() == kernel
            assert _rel(s.export_solution(), o.export_solution()) < REL, (cap, precond)
    c.close()


@pytest.mark.parametrize("degree", [2, 3, 5])
def test_chebyshev_preconditioner_matches_oracle(degree, fb, golden, oracles):
    """FB_PRECOND_CHEBYSHEV (polynomial of degree k in Dinv A, Gershgorin bound) solves the same system to the same
    absolute residual: potential within 1e-8 of the oracle, in fewer CG iterations than Jacobi"""
    m = golden("mesh", "mdbig"); o = oracles["mdbig"]
    c = fb.Context(0)
    c.set_option("cheb_degree", degree)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.5, 0.0); s.assemble(True)
    it_jacobi = s.solve()
    phi_jacobi = s.export_solution()
    s.conf.precond = fb.PRECOND_CHEBYSHEV
    s.setup(0.5, 0.0); s.assemble(True)
    it_cheb = s.solve()
    assert 0 < it_cheb < it_jacobi
    o.setup(0.5, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
    assert _rel(s.export_solution(), o.export_solution()) < REL
    assert _rel(s.export_solution(), phi_jacobi) < REL
    assert s.solve(cg_tolerance=1e-8) == 0                     # warm start
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve(n_cg=4) == -4                               # iteration cap keeps the reference's return convention
    c.close()


def test_chebyshev_on_hbm_sized_system(fb, golden):
    m = golden("mesh", "mdsmall")
    nodes, hexs, mk = synth.refine_vacuum(m["nodes"], m["hexs"], m["hex_markers"], 2)
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-9))
    assert s.import_mesh(nodes, hexs, mk)
    s.setup(0.5, 0.0); s.assemble(True)
    itj = s.solve(); phi = s.export_solution()
    s.conf.precond = fb.PRECOND_CHEBYSHEV
    c.set_option("cheb_degree", 3)
    s.setup(0.5, 0.0); s.assemble(True)
    itc = s.solve()
    assert 0 < itc < itj and s.solve_kernel() == 306
    assert _rel(s.export_solution(), phi) < 1e-7
    c.close()


@pytest.mark.parametrize("name", MESHES)
def test_persistent_and_multikernel_cg_agree(name, fb, golden):
    """the single-launch cooperative CG and the CUDA-graph multi-kernel CG run the same algorithm"""
    m = golden("mesh", name)
    out = []
    for persistent in (1, 0):
        c = fb.Context(0)
        c.set_option("cg_persistent", persistent)
        s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-10))
        s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        s.setup(0.5, 0.0); s.assemble(True)
        it = s.solve()
        assert it > 0
        launches = c.kernel_launches
        assert s.solve(cg_tolerance=1e-7) == 0           # warm start inside the same path
        out.append((it, s.export_solution(), launches))
        c.close()
    assert abs(out[0][0] - out[1][0]) <= 2
    assert _rel(out[0][1], out[1][1]) < 1e-9
    assert out[0][2] < 40 < out[1][2]                    # one cooperative launch vs 3 kernels per iteration


def test_uniform_field_exact(fb, ctx):
    nodes, hexs, mk = synth.box_mesh(7, 6, 9, 3.0, 2.5, 4.0, jitter=0.2)
    F = 0.37
    s = fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-12))
    s.import_mesh(nodes, hexs, mk)
    s.setup(F, 0.0); s.assemble(True)
    assert s.solve() > 0
    g = s.get_system()
    assert np.abs(s.export_solution() - F * nodes[g["vertex2node"], 2]).max() < 1e-11


def test_dirichlet_anode(fb, ctx, golden, oracles):
    m = golden("mesh", "hemicone"); o = oracles["hemicone"]
    conf = fb.FieldConfig(anode_BC="dirichlet", cg_tolerance=1e-11)
    s = fb.PoissonSolver(ctx, conf)
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.0, 55.0); s.assemble(True)
    assert s.solve() > 0
    o.setup(0.0, 55.0, True); o.assemble(True)
    assert o.solve(10000, 1e-11, 1.2, 0) > 0
    assert _rel(s.export_solution(), o.export_solution()) < REL
    g = s.get_system(); rhs = o.vectors()[0]
    assert np.abs(g["rhs"] - rhs).max() <= 1e-12 * np.abs(rhs).max()


def test_morton_numbering_gives_same_solution(fb, golden, oracles):
    m = golden("mesh", "mdsmall"); o = oracles["mdsmall"]
    c2 = fb.Context(0)
    c2.set_option("dof_order", 1)
    s = fb.PoissonSolver(c2, fb.FieldConfig(cg_tolerance=1e-11))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > 0
    o.setup(0.5, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
    assert _rel(s.export_solution(), o.export_solution()) < REL
    c2.close()


# ------------------------------------------------------------------------------------------
# interpolation half vs golden vectors of the reference's own code
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu_interp(fb, golden):
    out = {}
    for name in MESHES:
        m = golden("mesh", name)
        c = fb.Context(0)
        s = fb.PoissonSolver(c)
        s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        it = fb.Interpolator(c); it.initialize(m)
        out[name] = (c, s, it)
    yield out
    for c, _, _ in out.values():
        c.close()


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("rank", [1, 2, 3])
def test_locate_interpolate_golden(name, dim, rank, fb, golden, gpu_interp):
    g = golden("interp", name)
    c, s, it = gpu_interp[name]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    r = fb.SolutionReader(it); r.set_preferences(False, dim, rank)
    r.interpolate(g["points"])
    assert np.array_equal(r.markers, g["cells_d%dr%d" % (dim, rank)])            # bit-exact cell indices
    ref = g["sol_d%dr%d" % (dim, rank)]
    assert np.abs(r.interpolation - ref).max() <= 1e-12 * np.abs(ref).max()
    # cached-cell re-interpolation (SolutionReader.cpp:167-190)
    first = r.interpolation.copy()
    r.calc_interpolation()
    assert np.array_equal(first, r.interpolation)
    # SoA entry (the layout of femocs_interpolate_elfield) gives the same cells
    p = g["points"]
    _, cells = r.interpolate_results(p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy(), "elfield")
    assert np.array_equal(cells, r.markers)


@pytest.mark.parametrize("name", MESHES)
def test_particles_golden(name, fb, golden, gpu_interp):
    g = golden("interp", name)
    c, s, it = gpu_interp[name]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    pic = fb.Pic(it)
    pc = pic.update_point_cells(g["points"], g["pic_guess"])
    assert np.array_equal(pc, g["pic_cells"])
    assert np.array_equal(pic.update_point_cells(g["points"], np.maximum(pc, 0)), g["pic_cells2"])
    ok = g["pic_ok"]
    E = pic.fields(g["points"][ok], pc[ok])
    assert np.abs(E - g["pic_field"]).max() <= 1e-12 * np.abs(g["pic_field"]).max()


@pytest.mark.parametrize("rank", [1, 3])
def test_emission_lines_match_oracle(rank, fb, golden, gpu_interp, oracles):
    """batched EmissionReader::emission_line look-ups (32 potential samples along every face normal, each line an
    independent guess chain) == the oracle's chained locate run once per line: cells bit-exact"""
    m = golden("mesh", "mdsmall"); o = oracles["mdsmall"]
    c, s, it = gpu_interp["mdsmall"]
    nod = hash_field(it.n_nodes, 5, 1)
    it.set_solutions(nod); o.set_nodal(nod)
    tri = m["tris"][::7][:120]
    cent = m["nodes"][tri].mean(1); nrm = m["tri_norms"][::7][:120]
    rmax = np.linspace(2.0, 40.0, len(tri))
    t = np.linspace(1e-5, 1.0, 32)                                   # rmin = 1e-5 rmax (EmissionReader.cpp:52-58)
    lines = cent[:, None, :] + nrm[:, None, :] * (rmax[:, None, None] * t[None, :, None])
    r = fb.SolutionReader(it); r.set_preferences(False, 3, rank)
    cells, sol = r.interpolate_lines(lines)
    for k in range(len(lines)):
        oc, os_ = o.locate_interpolate(3, rank, lines[k])
        assert np.array_equal(cells[k], oc), "line %d" % k
        assert np.abs(sol[k] - os_).max() <= 1e-12 * max(1.0, np.abs(os_).max())
    # one chain over all points gives a different (chained) answer path but the batched call must not depend on batch order
    cells2, _ = r.interpolate_lines(lines[::-1].copy())
    assert np.array_equal(cells2[::-1], cells)


@pytest.mark.parametrize("name", ["hemicone", "mdsmall"])
def test_pic_push_golden(name, fb, golden, gpu_interp):
    """Pic::update_positions (+clear_lost) and Pic::update_velocities on the device against vectors produced with the
    reference's own cell search (tests/golden/picpush_*.npz): positions, cells, survivors bit-exact; velocities 1e-12"""
    g = golden("picpush", name)
    c, s, it = gpu_interp[name]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    pic = fb.Pic(it)
    dt = float(g["dt"][0]); qm = float(g["q_over_m"][0])
    for periodic in (1, 0):
        pos, vel, cells = g["pos0"], g["vel0"], g["cells0"]
        for step in range(3):
            tag = "p%d_s%d_" % (periodic, step)
            p1, v1, c1, lost = pic.update_positions(pos, vel, cells, dt, g["box"], bool(periodic))
            assert lost == int(g[tag + "lost"][0]) and len(c1) == len(g[tag + "cells"])
            assert np.array_equal(c1, g[tag + "cells"]) and np.array_equal(p1, g[tag + "pos"])
            v2 = pic.update_velocities(p1, v1, c1, dt, qm)
            assert np.abs(v2 - g[tag + "vel"]).max() <= 1e-12 * np.abs(g[tag + "vel"]).max()
            pos, vel, cells = g[tag + "pos"], g[tag + "vel"], g[tag + "cells"]      # next step starts from the golden state
    # nobody lost / everybody lost / empty set
    p1, v1, c1, lost = pic.update_positions(g["pos0"][:100], np.zeros((100, 3)), g["cells0"][:100], dt, g["box"], True)
    assert lost == 0 and np.array_equal(p1, g["pos0"][:100]) and np.array_equal(c1, g["cells0"][:100])
    up = np.tile(np.array([0.0, 0.0, 1e4]), (100, 1))
    p1, v1, c1, lost = pic.update_positions(g["pos0"][:100], up, g["cells0"][:100], dt, g["box"], True)
    assert lost == 100 and len(c1) == 0
    p1, v1, c1, lost = pic.update_positions(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.int32), dt, g["box"], True)
    assert lost == 0 and len(c1) == 0
    with pytest.raises(fb.FemocsB200Error):
        pic.update_positions(g["pos0"][:4], g["vel0"][:4], g["cells0"][:4], dt, np.array([1.0, 0.0, 0.0, 1.0, 0.0, 1.0]), True)


@pytest.mark.parametrize("name", MESHES)
def test_extract_solution_golden(name, fb, golden, gpu_interp):
    m = golden("mesh", name); g = golden("interp", name)
    c, s, it = gpu_interp[name]
    phi = hash_field(s.n_vertices, 1, 2)[:, 0]
    s.import_solution(phi)
    assert np.array_equal(s.export_solution(), phi)
    vac = m["node_femocs2deal"] >= 0
    it.extract_solution(s, False)
    nod = it.get_solutions()
    assert np.all(nod[~vac] == 0) and np.array_equal(nod[vac, 4], phi) and np.all(nod[:, 3] == 0)
    scale = np.abs(g["extract_E"]).max()
    assert np.abs(nod[vac, :3] - g["extract_E"]).max() <= 1e-12 * scale
    n_voro = len(m["voro_off"]) - 1
    it.extract_solution(s, True)
    nod_s = it.get_solutions()
    assert np.abs(nod_s[:n_voro, :3] - g["extract_E_smooth_tetnodes"]).max() <= 1e-12 * scale
    assert np.array_equal(nod_s[n_voro:], nod[n_voro:])


@pytest.mark.parametrize("name", ["mdsmall"])
def test_space_charge_rhs_matches_oracle(name, fb, golden, gpu_interp, oracles):
    m = golden("mesh", name); g = golden("interp", name); o = oracles[name]
    c, s, it = gpu_interp[name]
    ok = g["pic_ok"]
    pts = g["points"][ok]; cells = g["pic_cells"][ok]
    cf = -180.9512268 * 0.01                                       # q_over_eps0 * Wsp (Pic.h:91, Config.cpp:110)
    s.conf.mode = "transient"
    s.set_particles(pts, cells, cf)
    s.setup(0.5, 0.0); s.assemble(True)
    o.setup(0.5, 0.0, False); o.assemble(True, pts, cells, cf)
    rhs = o.vectors()[0]
    assert np.abs(s.get_system()["rhs"] - rhs).max() <= 1e-11 * np.abs(rhs).max()
    # PIC step: matrix restored, RHS rebuilt, warm-started solve (ProjectRunaway.cpp:492-533)
    s.conf.cg_tolerance = 1e-11
    assert s.solve() > 0
    s.assemble(False)
    assert np.abs(s.get_system()["rhs"] - rhs).max() <= 1e-11 * np.abs(rhs).max()
    assert s.solve(cg_tolerance=1e-8) == 0          # warm start from the converged potential
    o.solve(10000, 1e-11, 1.2, 0)
    assert _rel(s.export_solution(), o.export_solution()) < REL
    s.conf.mode = "laplace"; s.set_particles(None, None, 0)


def test_space_charge_without_interpolator_tables(fb, golden, oracles):
    """solver-only context (no fb_interp_initialize): the hexahedron coefficients are rebuilt from the cell vertices"""
    m = golden("mesh", "mdsmall"); g = golden("interp", "mdsmall"); o = oracles["mdsmall"]
    ok = g["pic_ok"]
    pts = g["points"][ok]; cells = g["pic_cells"][ok]
    cf = -180.9512268 * 0.01
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(mode="transient"))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.set_particles(pts, cells, cf)
    s.setup(0.5, 0.0); s.assemble(True)
    o.setup(0.5, 0.0, False); o.assemble(True, pts, cells, cf)
    rhs = o.vectors()[0]
    assert np.abs(s.get_system()["rhs"] - rhs).max() <= 1e-11 * np.abs(rhs).max()
    c.close()


def test_full_step_matches_oracle(fb, golden, oracles):
    """solve -> extract (smoothed) -> fields on surface atoms: the per-MD-step path of config 2."""
    m = golden("mesh", "mdbig"); o = oracles["mdbig"]
    E0 = -0.5
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(E0=E0, cg_tolerance=1e-11))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(-E0, 0.0); s.assemble(True)
    assert s.solve() > 0
    it = fb.Interpolator(c); it.initialize(m); it.extract_solution(s, True)
    f = fb.FieldReader(it); f.set_preferences(False, 2, 1); f.interpolate(m["surf_atoms"], m["surf_ids"])
    o.setup(-E0, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
    nod = o.extract_solution(True)
    cells, sol = o.locate_interpolate(2, 1, m["surf_atoms"])
    assert np.array_equal(f.markers, cells)
    assert _rel(it.get_solutions(), nod) < REL
    assert _rel(f.interpolation, sol) < REL
    assert abs(f.E_max - np.sqrt((sol[:, :3] ** 2).sum(1)).max()) < REL * f.E_max
    # interpolate_elfield on ALL atoms (dim 3), mostly bulk atoms -> nearest-cell fallback
    f3 = fb.FieldReader(it); f3.set_preferences(False, 3, 1); f3.interpolate(m["atoms"][::7])
    cells3, sol3 = o.locate_interpolate(3, 1, m["atoms"][::7])
    assert np.array_equal(f3.markers, cells3)
    assert _rel(f3.interpolation, sol3) < REL
    c.close()


def test_edge_cases(fb, golden, gpu_interp):
    c, s, it = gpu_interp["hemicone"]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    r = fb.SolutionReader(it); r.set_preferences(False, 3, 3)
    r.interpolate(np.zeros((0, 3)))                                 # empty input
    assert r.size() == 0 and len(r.markers) == 0
    m = golden("mesh", "hemicone")
    far = m["nodes"].max(0) + 1000.0                                # a single point far outside
    r.interpolate(far[None, :])
    assert r.markers[0] <= 0
    with pytest.raises(ValueError):
        r.set_preferences(False, 4, 1)
    with pytest.raises(fb.FemocsB200Error):
        fb.PoissonSolver(fb.Context(0)).setup(1.0)                  # call order error, no crash


# ------------------------------------------------------------------------------------------
# large sizes: properties that do not need the oracle
# ------------------------------------------------------------------------------------------
def test_large_refined_mesh_properties(fb, golden):
    m = golden("mesh", "mdsmall")
    nodes, hexs, mk = synth.refine_vacuum(m["nodes"], m["hexs"], m["hex_markers"], 2)    # ~0.9 M hexes
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-9))
    assert s.import_mesh(nodes, hexs, mk)
    assert s.n_cells == 64 * int((m["hex_markers"] > 0).sum())
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > 0
    assert s.solve_kernel() == 306          # HBM-sized systems take the segmented block-JDS SpMV (evict-first matrix stream)
    phi1 = s.export_solution()
    c.set_option("spmv_kernel", 310)         # the symmetric (lower-triangle) layout gives the same solution
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > 0 and s.solve_kernel() == 310
    assert _rel(s.export_solution(), phi1) < 1e-7        # both stop at |r| <= 1e-9; phi ~ 1e2
    c.set_option("spmv_kernel", -1)
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > 0 and s.solve_kernel() == 306
    g = s.get_system()
    K = sp.csr_matrix((g["val_save"], g["col"], g["rowptr"]))
    assert abs(K - K.T).max() <= 1e-12 * np.abs(g["val_save"]).max()                     # symmetry
    assert np.abs(np.asarray(K.sum(1))).max() <= 1e-10 * np.abs(g["val_save"]).max()     # constants in kernel
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]))
    drift = 1e-13 * abs(A).sum(1).max() * np.linalg.norm(g["sol"])
    assert np.linalg.norm(A @ g["sol"] - g["rhs"]) <= 2e-9 + drift                       # residual
    # linearity in the applied field
    s.setup(1.0, 0.0); s.assemble(True); s.conf.cg_tolerance = 2e-9
    assert s.solve() > 0
    assert _rel(s.export_solution(), 2 * phi1) < 1e-7
    # the refined solution approaches the coarse one on the shared (coarse) vertices
    oc = Oracle(); oc.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    oc.setup(0.5, 0.0, False); oc.assemble(True); oc.solve(10000, 1e-9, 1.2, 0)
    coarse = oc.export_solution()
    n_coarse = oc.n_vertices
    # refine_vacuum keeps the coarse vertices first, in order
    v2n = g["vertex2node"]
    inv = np.full(len(nodes), -1); inv[v2n] = np.arange(len(v2n))
    fine_on_coarse = phi1[inv[:n_coarse]]
    assert _rel(fine_on_coarse, coarse) < 0.05
    c.close()


# ------------------------------------------------------------------------------------------
# round 2: mask-based Dirichlet conditions, Chebyshev x SpMV variants, stale particle cells
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [0, 8, 100, 200, 300, 304, 305, 310, 311])
def test_chebyshev_with_every_spmv_layout(kernel, fb, golden, oracles):
    """FB_PRECOND_CHEBYSHEV multiplies by the full matrix k - 1 times per iteration: with the symmetric (lower
    triangle) layouts 310 / 311 selected the solver must switch to the full block-JDS layout, not multiply by L"""
    m = golden("mesh", "mdsmall"); o = oracles["mdsmall"]
    c = fb.Context(0)
    c.set_option("cg_persistent", 0)
    c.set_option("spmv_kernel", kernel)
    c.set_option("cheb_degree", 3)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11, precond=fb.PRECOND_CHEBYSHEV))
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    s.setup(0.5, 0.0); s.assemble(True)
    it = s.solve()
    assert it > 0
    if kernel >= 310:
        assert s.solve_kernel() == 304
    o.setup(0.5, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
    assert _rel(s.export_solution(), o.export_solution()) < REL
    # then Jacobi on the same context and layout (the symmetric kernel accumulates into h: it must start from zero)
    s.conf.precond = fb.PRECOND_JACOBI
    s.setup(0.5, 0.0); s.assemble(True)
    assert s.solve() > it
    assert _rel(s.export_solution(), o.export_solution()) < REL
    c.close()


def test_pic_step_does_no_matrix_work(fb, golden, oracles):
    """assemble(false) (every PIC step but the first, ProjectRunaway.cpp:497) leaves the stiffness matrix and its
    block-JDS copy alone: the Dirichlet conditions are a mask, not an elimination of a restored copy"""
    m = golden("mesh", "mdsmall"); o = oracles["mdsmall"]
    for persistent in (1, 0):
        c = fb.Context(0)
        c.set_option("cg_persistent", persistent)
        if not persistent:
            c.set_option("spmv_kernel", 304)
        s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11))
        s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        s.setup(0.5, 0.0); s.assemble(True)
        assert s.solve() > 0
        n0 = c.kernel_launches
        s.assemble(False)
        assert c.kernel_launches - n0 == 2              # Neumann faces + constrained values; nothing per non-zero
        n1 = c.kernel_launches
        assert s.solve(cg_tolerance=1e-8) == 0          # warm start: initial residual only
        assert c.kernel_launches - n1 <= 3
        # a different applied field on the same matrix (assemble(false) after setup keeps K: reference restores its copy)
        o.setup(0.5, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
        assert _rel(s.export_solution(), o.export_solution()) < REL
        g = s.get_system()                              # eliminated system materialised on demand
        rp, col, val, save = o.csr()
        assert np.abs(g["val"] - val).max() <= 1e-12 * np.abs(save).max()
        assert np.abs(g["rhs"] - o.vectors()[0]).max() <= 1e-12 * np.abs(o.vectors()[0]).max()
        c.close()


def test_stale_particle_cells_are_not_dereferenced(fb, golden, gpu_interp):
    """cell ids left over from a larger mesh (Pic.cpp:188-191) are treated as 'no guess' by the search and give no
    field / no acceleration -- never an out-of-bounds read of the cell maps"""
    g = golden("interp", "hemicone")
    c, s, it = gpu_interp["hemicone"]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    pic = fb.Pic(it)
    pts = g["points"]; n = len(pts)
    stale = np.full(n, 2 ** 30, np.int32)
    fresh = pic.update_point_cells(pts, np.zeros(n, np.int32))
    assert np.array_equal(pic.update_point_cells(pts, stale), fresh)
    E = pic.fields(pts[:50], stale[:50])
    assert np.all(E == 0)
    v = pic.update_velocities(pts[:50], np.ones((50, 3)), stale[:50], 0.5, -17.5882)
    assert np.all(v == 1)


@pytest.mark.parametrize("name", MESHES)
def test_export_solution_grad_counts_and_mesh_export(name, fb, ctx, golden, oracles):
    """DealSolver::export_solution_grad, the #faces / #edges of operator<< and the mesh behind write('*.vtk'|'*.msh')"""
    m = golden("mesh", name); o = oracles[name]
    s = fb.PoissonSolver(ctx)
    s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    phi = hash_field(s.n_vertices, 1, 3)[:, 0]
    s.import_solution(phi)
    _, _, v2d, v2n = o.vectors()
    sol = np.zeros(o.n_dofs); sol[v2d] = phi
    o.set_solution(sol)
    ref = o.export_solution_grad()
    assert np.abs(s.export_solution_grad() - ref).max() <= 1e-11 * np.abs(ref).max()
    assert s.mesh_counts() == o.mesh_counts()
    nf, ne = o.mesh_counts()
    assert s.to_str() == "#elems=%d, #faces=%d, #edges=%d, #nodes=%d, #dofs=%d" % (o.n_cells, nf, ne, o.n_vertices, o.n_dofs)
    xyz, cells = s.solver_mesh()
    assert np.array_equal(xyz, m["nodes"][v2n])
    vac = m["hex_markers"] > 0
    assert np.array_equal(xyz[cells], m["nodes"][m["hexs"][vac]])
