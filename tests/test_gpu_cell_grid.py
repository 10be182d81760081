"""The uniform-grid filter of the tetrahedron scan (locate_cell's linear-scan tail, InterpolatorCells.cpp:288-306) must
return EXACTLY what the brute-force scan returns -- first hit in index order, else minus the first index attaining the
minimal centroid distance -- for points inside, on and far outside the mesh; and through it the oracle's answers."""
import numpy as np
import pytest

from oracle.fields import hash_field
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


def _points(m, seed=1):
    rng = np.random.default_rng(seed)
    nodes = m["nodes"]; lo, hi = nodes.min(0), nodes.max(0)
    cent = nodes[m["tets"]].mean(1)
    pts = [rng.uniform(lo, hi, size=(4000, 3)),                                   # inside the box
           rng.uniform(lo - 0.7 * (hi - lo), hi + 0.7 * (hi - lo), size=(3000, 3)),  # around it: nearest-centroid answers
           nodes[rng.integers(0, len(nodes), 1500)],                              # exactly on vertices: several cells hit
           cent[rng.integers(0, len(cent), 1500)],                                # exactly on centroids
           0.5 * (cent[:-1] + cent[1:])[:1500],                                   # midway between two centroids: ties in distance
           np.array([[1e4, -2e4, 3e4], [lo[0], lo[1], lo[2]], [hi[0], hi[1], hi[2]]])]
    if "atoms" in m:
        pts.append(m["atoms"][::3])
    p = np.ascontiguousarray(np.vstack(pts))
    return p[rng.permutation(len(p))]                                             # incoherent order: the guess chain helps nobody


@pytest.mark.parametrize("name", ["hemicone", "mdsmall", "mdbig"])
def test_grid_scan_equals_brute_force_and_oracle(name, golden):
    import femocs_b200 as fb
    m = golden("mesh", name)
    pts = _points(m)
    sol5 = hash_field(len(m["nodes"]), 5, 1)
    res = {}
    for grid in (1, 0):
        ctx = fb.Context(0)
        ctx.set_option("cell_grid", grid)
        s = fb.PoissonSolver(ctx); assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        it = fb.Interpolator(ctx); it.initialize(m); it.set_solutions(sol5)
        for rank in (1, 2, 3):
            f = fb.FieldReader(it); f.set_preferences(False, 3, rank); f.interpolate(pts)
            res[(grid, rank)] = (f.markers.copy(), f.interpolation.copy())
        ctx.close()
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m); o.set_nodal(sol5)
    for rank in (1, 2, 3):
        cg, sg = res[(1, rank)]; cb, sb = res[(0, rank)]
        assert np.array_equal(cg, cb), (name, rank, int((cg != cb).sum()))
        assert np.array_equal(sg, sb)
        oc, osol = o.locate_interpolate(3, rank, pts)
        assert np.array_equal(cg, oc), (name, rank, int((cg != oc).sum()))
        assert np.abs(sg - osol).max() <= 1e-12 * max(1.0, np.abs(osol).max())
        assert (cg < 0).sum() > 1000 and (cg >= 0).sum() > 1000                   # both branches of the scan were exercised
