"""The device-resident entry points (fb_*_dev: inputs and outputs already in HBM, asynchronous on the context's
stream) are what bench.py times; these tests hold every one of them to the same bar as the host-buffer entry
points: golden vectors of the reference's own code (cells bit-exact, 1e-12 weights) and the CPU oracle.
torch is used only to own device memory."""
import ctypes as C

import numpy as np
import pytest

from oracle.fields import hash_field
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

MESHES = ["hemicone", "mdsmall", "mdbig"]


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


@pytest.fixture(scope="module")
def fb():
    import femocs_b200
    return femocs_b200


@pytest.fixture(scope="module")
def gpu(fb, golden):
    out = {}
    for name in MESHES:
        m = golden("mesh", name)
        c = fb.Context(0)
        s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11, mode="transient"))
        assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        it = fb.Interpolator(c); it.initialize(m)
        out[name] = (c, s, it)
    yield out
    for c, _, _ in out.values():
        c.close()


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("rank", [1, 2, 3])
def test_locate_interpolate_dev_golden(name, dim, rank, fb, torch, golden, gpu):
    g = golden("interp", name)
    c, s, it = gpu[name]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    n = len(g["points"])
    d_pts = _dev(torch, g["points"])
    d_cells = torch.full((n,), -12345, dtype=torch.int32, device="cuda")
    d_sol = torch.zeros(n, 5, dtype=torch.float64, device="cuda")
    c.check(c.L.fb_locate_interpolate_dev(c.h, dim, rank, n, d_pts.data_ptr(), d_cells.data_ptr(), d_sol.data_ptr()))
    c.synchronize()
    assert np.array_equal(d_cells.cpu().numpy(), g["cells_d%dr%d" % (dim, rank)])
    ref = g["sol_d%dr%d" % (dim, rank)]
    assert np.abs(d_sol.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    # a second call on the same buffers (the bench repeats the call): same answer, inputs untouched
    c.check(c.L.fb_locate_interpolate_dev(c.h, dim, rank, n, d_pts.data_ptr(), d_cells.data_ptr(), d_sol.data_ptr()))
    c.synchronize()
    assert np.array_equal(d_cells.cpu().numpy(), g["cells_d%dr%d" % (dim, rank)])
    assert np.array_equal(d_pts.cpu().numpy(), g["points"])


@pytest.mark.parametrize("name", MESHES)
def test_particle_cells_and_field_dev_golden(name, fb, torch, golden, gpu):
    g = golden("interp", name)
    c, s, it = gpu[name]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    n = len(g["points"])
    d_pts = _dev(torch, g["points"]); d_cells = _dev(torch, g["pic_guess"])
    c.check(c.L.fb_particle_cells_dev(c.h, n, d_pts.data_ptr(), d_cells.data_ptr()))
    c.synchronize()
    pc = d_cells.cpu().numpy()
    assert np.array_equal(pc, g["pic_cells"])
    d_cells2 = _dev(torch, np.maximum(pc, 0))
    c.check(c.L.fb_particle_cells_dev(c.h, n, d_pts.data_ptr(), d_cells2.data_ptr()))
    c.synchronize()
    assert np.array_equal(d_cells2.cpu().numpy(), g["pic_cells2"])
    ok = g["pic_ok"]
    d_p = _dev(torch, g["points"][ok]); d_c = _dev(torch, pc[ok])
    d_E = torch.zeros(int(ok.sum()), 3, dtype=torch.float64, device="cuda")
    c.check(c.L.fb_particle_field_dev(c.h, int(ok.sum()), d_p.data_ptr(), d_c.data_ptr(), d_E.data_ptr()))
    c.synchronize()
    assert np.abs(d_E.cpu().numpy() - g["pic_field"]).max() <= 1e-12 * np.abs(g["pic_field"]).max()


@pytest.mark.parametrize("name", ["hemicone", "mdsmall"])
def test_pic_push_dev_golden(name, fb, torch, golden, gpu):
    """fb_pic_update_positions_dev / fb_pic_update_velocities_dev with the particles resident in HBM across the
    steps (the bench's PIC leg): positions, cells and survivors bit-exact with the reference-made goldens"""
    g = golden("picpush", name)
    c, s, it = gpu[name]
    it.set_solutions(hash_field(it.n_nodes, 5, 1))
    dt = float(g["dt"][0]); qm = float(g["q_over_m"][0])
    box = np.ascontiguousarray(g["box"])
    for periodic in (1, 0):
        d_pos = _dev(torch, g["pos0"]); d_vel = _dev(torch, g["vel0"]); d_cell = _dev(torch, g["cells0"])
        n = len(g["cells0"])
        for step in range(3):
            tag = "p%d_s%d_" % (periodic, step)
            lost = C.c_long(-1)
            c.check(c.L.fb_pic_update_positions_dev(c.h, n, d_pos.data_ptr(), d_vel.data_ptr(), d_cell.data_ptr(), dt,
                                                    box.ctypes.data, periodic, C.byref(lost)))
            assert lost.value == int(g[tag + "lost"][0])
            n -= lost.value
            assert n == len(g[tag + "cells"])
            assert np.array_equal(d_cell[:n].cpu().numpy(), g[tag + "cells"])
            assert np.array_equal(d_pos[:n].cpu().numpy(), g[tag + "pos"])
            c.check(c.L.fb_pic_update_velocities_dev(c.h, n, d_pos.data_ptr(), d_cell.data_ptr(), d_vel.data_ptr(), dt, qm))
            c.synchronize()
            v = d_vel[:n].cpu().numpy()
            assert np.abs(v - g[tag + "vel"]).max() <= 1e-12 * np.abs(g[tag + "vel"]).max()
            # the next step starts from the golden velocities (1e-12 differences must not move a particle across a face)
            d_vel[:n] = _dev(torch, g[tag + "vel"])
    lost = C.c_long(-1)
    c.check(c.L.fb_pic_update_positions_dev(c.h, 0, 0, 0, 0, dt, box.ctypes.data, 1, C.byref(lost)))      # empty set
    assert lost.value == 0


def test_poisson_assemble_dev_matches_oracle(fb, torch, golden, gpu):
    """fb_poisson_assemble_dev (space-charge RHS from particles in HBM) == host-buffer entry point == oracle, for the
    first step (first_time = 1) and for a PIC step (first_time = 0, warm start)"""
    name = "mdsmall"
    m = golden("mesh", name); g = golden("interp", name)
    c, s, it = gpu[name]
    ok = g["pic_ok"]
    pts = np.ascontiguousarray(g["points"][ok]); cells = np.ascontiguousarray(g["pic_cells"][ok])
    cf = -180.9512268 * 0.01
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    o.setup(0.5, 0.0, False); o.assemble(True, pts, cells, cf)
    rhs = o.vectors()[0]
    d_p = _dev(torch, pts); d_c = _dev(torch, cells)
    s.setup(0.5, 0.0)
    s.assemble_dev(True, d_p.data_ptr(), d_c.data_ptr(), len(cells), cf)
    c.synchronize()
    sys_dev = s.get_system()
    assert np.abs(sys_dev["rhs"] - rhs).max() <= 1e-11 * np.abs(rhs).max()
    _, _, val, save = o.csr()
    assert np.abs(sys_dev["val_save"] - save).max() <= 1e-12 * np.abs(save).max()
    assert np.abs(sys_dev["val"] - val).max() <= 1e-12 * np.abs(save).max()
    assert s.solve() > 0
    o.solve(10000, 1e-11, 1.2, 0)
    ref = o.export_solution()
    assert np.abs(s.export_solution() - ref).max() <= 1e-8 * np.abs(ref).max()
    # PIC step: half of the particles gone, warm start
    k = len(cells) // 2
    s.assemble_dev(False, d_p.data_ptr(), d_c.data_ptr(), k, cf)
    c.synchronize()
    o.assemble(False, pts[:k], cells[:k], cf)
    rhs2 = o.vectors()[0]
    assert np.abs(s.get_system()["rhs"] - rhs2).max() <= 1e-11 * np.abs(rhs2).max()
    assert s.solve() > 0
    o.solve(10000, 1e-11, 1.2, 0)
    ref2 = o.export_solution()
    assert np.abs(s.export_solution() - ref2).max() <= 1e-8 * np.abs(ref2).max()
    # no particles at all through the _dev entry = Laplace
    s.setup(0.5, 0.0); s.assemble_dev(True, 0, 0, 0, 0.0); c.synchronize()
    o.setup(0.5, 0.0, False); o.assemble(True)
    rhs0 = o.vectors()[0]
    assert np.abs(s.get_system()["rhs"] - rhs0).max() <= 1e-12 * np.abs(rhs0).max()


def test_bench_pic_step_sequence_matches_oracle(fb, torch, golden):
    """The exact call sequence of bench.py's PIC leg (all _dev entry points, particles resident, lost electrons
    re-seeded) on a small population, checked step by step against the CPU oracle run on the same state."""
    import bench
    from oracle import pic as opic
    m = golden("mesh", "mdsmall")
    leg = bench.PicLeg(fb, torch, m, n_particles=4000, device=0)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    o.setup(-bench.E0, 0.0, False); o.assemble(True); o.solve(bench.N_CG, 1e-11, 1.2, 0); o.extract_solution(True)
    leg.solver.conf.cg_tolerance = 1e-11
    leg.start()
    for step in range(3):
        pos, vel, cell = leg.state()
        p1, v1, c1, lost = opic.update_positions(o, pos, vel, cell, leg.dt, leg.box, True)
        o.assemble(False, p1, c1, leg.cf); assert o.solve(bench.N_CG, 1e-11, 1.2, 0) >= 0
        o.extract_solution(True)
        v2 = opic.update_velocities(o, p1, v1, c1, leg.dt, leg.q_over_m)
        leg.step(reseed=False)
        gp, gv, gc = leg.state()
        assert leg.n_lost_last == lost and len(gc) == len(c1)
        assert np.array_equal(gc, c1) and np.array_equal(gp, p1)
        ref = o.export_solution()
        assert np.abs(leg.solver.export_solution() - ref).max() <= 1e-8 * np.abs(ref).max()
        assert np.abs(gv - v2).max() <= 1e-8 * np.abs(v2).max()
        leg.set_velocities(v2)           # keep both sides on the same trajectory
    # re-seeding keeps the population at its nominal size
    leg.step(reseed=True)
    assert leg.n == 4000
    leg.close()


def test_charge_density_when_the_host_writes_files(golden):
    """PoissonSolver.cpp:196-207: with write_time() true the solver keeps charge_density = rhs / dof_volume (taken before
    the Dirichlet conditions) and Interpolator::extract_solution stores it as scalar1; zeros otherwise"""
    import femocs_b200 as fb
    from oracle.oracle import Oracle
    m = golden("mesh", "mdsmall"); g = golden("interp", "mdsmall")
    ok = g["pic_ok"]; cf = -180.9512268 * 0.01
    c = fb.Context(0)
    s = fb.PoissonSolver(c, fb.FieldConfig(cg_tolerance=1e-11, mode="transient"))
    assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    it = fb.Interpolator(c); it.initialize(m)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    s.set_particles(g["points"][ok], g["pic_cells"][ok], cf)
    s.setup(0.5, 0.0); s.assemble(True); assert s.solve() > 0
    assert np.all(s.export_charge_dens() == 0)
    c.set_option("charge_density", 1); o.set_write_time(True)
    s.setup(0.5, 0.0); s.assemble(True); assert s.solve() > 0
    o.setup(0.5, 0.0, False); o.assemble(True, g["points"][ok], g["pic_cells"][ok], cf); assert o.solve(10000, 1e-11, 1.2, 0) > 0
    rho, rho_o = s.export_charge_dens(), o.export_charge_dens()
    assert np.abs(rho_o).max() > 0 and np.abs(rho - rho_o).max() <= 1e-11 * np.abs(rho_o).max()
    it.extract_solution(s, True)
    nod, nod_o = it.get_solutions(), o.extract_solution(True)
    assert np.abs(nod[:, 3] - nod_o[:, 3]).max() <= 1e-11 * np.abs(nod_o[:, 3]).max()
    assert np.abs(nod - nod_o).max() <= 1e-8 * np.abs(nod_o).max()
    c.close()
