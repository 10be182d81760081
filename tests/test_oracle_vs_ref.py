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
