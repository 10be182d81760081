"""Differential test: the oracle against the reference's own compiled code, on fresh meshes and
fresh seeded inputs (more points than the committed golden vectors).  Needs oracle/_ref and the
reference's input files, i.e. only runs in the build container; skipped elsewhere."""
import numpy as np
import pytest

from oracle import reflib
from oracle.oracle import Oracle

pytestmark = pytest.mark.skipif(not (reflib.available() and reflib.reference_inputs_available()),
                                reason="oracle/_ref or /root/reference/in not present")


@pytest.mark.parametrize("preset", ["hemicone", "mdsmall"])
def test_oracle_matches_reference_code(preset):
    r = reflib.RefLib()
    m = r.generate(preset)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    rng = np.random.default_rng(7)
    sol5 = rng.normal(size=(len(m["nodes"]), 5))
    r.set_nodal(sol5); o.set_nodal(sol5)
    lo = m["nodes"].min(0); hi = m["nodes"].max(0)
    pts = rng.uniform(lo, hi, size=(1500, 3))
    if "surf_atoms" in m:
        pts = np.vstack([m["surf_atoms"], m["atoms"][::9], pts])
    for dim in (2, 3):
        for rank in (1, 2, 3):
            c1, s1 = r.locate_interpolate(dim, rank, pts)
            c2, s2 = o.locate_interpolate(dim, rank, pts)
            assert np.array_equal(c1, c2) and np.array_equal(s1, s2)
    guess = rng.integers(0, o.n_cells, size=len(pts)).astype(np.int32)
    pc = r.particle_cells(pts, guess)
    assert np.array_equal(pc, o.particle_cells(pts, guess))
    ok = pc >= 0
    assert np.array_equal(r.particle_field(pts[ok], pc[ok]), o.particle_field(pts[ok], pc[ok]))
    assert np.array_equal(r.particle_weights(pts[ok], pc[ok]), o.particle_weights(pts[ok], pc[ok]))
    _, _, v2d, _ = o.vectors()
    phi = rng.normal(size=o.n_vertices)
    sd = np.zeros(o.n_dofs); sd[v2d] = phi; o.set_solution(sd)
    for sm in (0, 1):
        assert np.array_equal(r.extract_solution(phi, np.zeros_like(phi), sm, len(m["nodes"])), o.extract_solution(sm))


def test_export_results_mirror_matches_reference_code():
    """SolutionReader::export_results (label-case append / overwrite, id filtering; SolutionReader.cpp:269-398) of the
    compiled reference against the host mirror femocs_b200.SolutionReader.export_results on the same atoms"""
    import types
    import femocs_b200 as fb
    r = reflib.RefLib()
    r.generate("hemicone")
    rng = np.random.default_rng(3)
    n, n_points = 300, 260
    ids = rng.permutation(n).astype(np.int32) - 15              # some ids < 0 and some >= n_points: skipped by both
    ids[5] = ids[6]                                             # a repeated id: appended twice / last one wins
    sol = rng.normal(size=(n, 5))
    mirror = fb.SolutionReader(types.SimpleNamespace(ctx=None))
    mirror.points = np.zeros((n, 3)); mirror.ids = ids; mirror.interpolation = sol; mirror.markers = np.zeros(n, np.int32)
    for label in ("elfield", "ELFIELD", "Elfield", "elfield_norm", "ELFIELD_NORM", "charge_density", "CHARGE_DENSITY",
                  "potential", "POTENTIAL"):
        width = 3 if label.lower() == "elfield" else 1
        a = rng.normal(size=n_points * width); b = a.copy()
        assert r.export_results(ids, sol, n_points, label, a) == 0
        assert mirror.export_results(n_points, label, b) == 0
        assert np.array_equal(a, b), label
    # nothing to export -> 1 (check_return), data untouched
    empty = fb.SolutionReader(types.SimpleNamespace(ctx=None))
    a = np.ones(4)
    assert empty.export_results(4, "elfield_norm", a) == 1 and np.all(a == 1)
    assert r.export_results(np.zeros(0, np.int32), np.zeros((0, 5)), 4, "elfield_norm", a) == 1 and np.all(a == 1)


def test_pic_push_pieces_match_reference_code():
    """the two pieces of the PIC push that compile from the reference without the solver -- periodic_image
    (Macros.cpp:41-48) and ParticleSpecies::clear_lost (ParticleSpecies.cpp:16-31) -- against oracle/pic.py"""
    from oracle import pic
    r = reflib.RefLib()
    rng = np.random.default_rng(11)
    p = rng.uniform(-30.0, 30.0, size=5000)
    p[:6] = [-10.0, 10.0, -10.0 - 1e-13, 10.0 + 1e-13, 0.0, 25.0]       # exactly on / just past the box faces
    assert np.array_equal(r.periodic_image(p, 10.0, -10.0), pic.periodic_image(p, 10.0, -10.0))

    class Keep:                                   # a locator that keeps the incoming cell: only the bookkeeping is tested
        def particle_cells(self, xyz, guess):
            return np.asarray(guess, np.int32).copy()

    n = 4000
    pos = rng.normal(size=(n, 3)); vel = rng.normal(size=(n, 3))
    cells = rng.integers(-1, 50, size=n).astype(np.int32)                # about 2 % carry cell == -1
    p_ref, v_ref, c_ref, lost_ref = r.clear_lost(pos, vel, cells)
    box = (-1e9, 1e9, -1e9, 1e9, -1e9, 1e9)
    p_o, v_o, c_o, lost_o = pic.update_positions(Keep(), pos, np.zeros_like(vel), cells, 1.0, box, True)   # zero velocity: pure clear_lost
    assert lost_ref == lost_o == int((cells == -1).sum()) and lost_ref > 0
    assert np.array_equal(c_ref, c_o) and np.array_equal(p_ref, p_o)
    assert np.array_equal(v_ref, vel[cells != -1])                       # survivors keep their order


def test_known_cell_interpolation_linhex_locate_and_nodal_gradients_match_reference_code():
    """the remaining cell-class entry points of the compiled reference against the oracle, bit for bit:
    interp_solution with given cells (SolutionReader.cpp:167-190), LinearHexahedra::locate_cell (:1536-1562),
    LinearHexahedra::interp_gradient(hex, node) (:425-439, :1462-1505)"""
    r = reflib.RefLib()
    m = r.generate("mdsmall")
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    rng = np.random.default_rng(21)
    sol5 = rng.normal(size=(len(m["nodes"]), 5))
    r.set_nodal(sol5); o.set_nodal(sol5)
    lo = m["nodes"].min(0); hi = m["nodes"].max(0)
    pts = np.vstack([m["surf_atoms"][::3], rng.uniform(lo, hi, size=(800, 3))])
    for dim in (2, 3):
        for rank in (1, 2, 3):
            cells, _ = r.locate_interpolate(dim, rank, pts)
            assert np.array_equal(r.interp_known_cells(dim, rank, pts, cells), o.interpolate(dim, rank, pts, cells))
    guess = rng.integers(0, len(m["hexs"]), size=len(pts)).astype(np.int32)
    assert np.array_equal(r.linhex_locate(pts, guess), o.linhex_locate(pts, guess))
    vac = np.flatnonzero(m["hex_markers"] > 0)
    hexs = rng.choice(vac, size=3000).astype(np.int32); nodes = rng.integers(0, 8, size=3000).astype(np.int32)
    assert np.array_equal(r.nodal_gradient(hexs, nodes), o.nodal_gradient(hexs, nodes))
