"""TEST INFRASTRUCTURE -- regenerates tests/golden/*.npz from the reference's own code.

Run in the build container (needs /root/reference and oracle/_ref/libfemocs_ref.so):

    make -f oracle/Makefile.ref && python oracle/make_golden.py

For each scenario (Main.cpp presets restated in oracle/reflib.py) it stores
  mesh_<name>.npz    the arrays the reference's mesher hands to the hot path (SURVEY 8a'),
                     produced by the verbatim AtomReader -> Surface -> TetGen -> Tethex code;
  interp_<name>.npz  seeded inputs and the outputs of the reference's verbatim
                     Interpolator / InterpolatorCells / SolutionReader code on them:
                     chained locate+interpolate for every (dim, rank), PIC particle
                     locate / weights / field, and Interpolator::extract_solution.
The solver half (deal.II) has no reference-produced golden: parity unpinned (see
oracle/femocs_oracle.cpp header).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.fields import hash_field  # noqa: E402
from oracle.reflib import RefLib  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_SURF, N_ATOM, N_RAND = 700, 700, 600


def main():
    os.makedirs(OUT, exist_ok=True)
    r = RefLib()
    for name in ("hemicone", "mdsmall", "mdbig"):
        m = r.generate(name)
        np.savez_compressed(os.path.join(OUT, "mesh_%s.npz" % name), **m)

        rng = np.random.default_rng(20240 + len(name))
        n_nodes = len(m["nodes"])
        g = {}
        # a seeded nodal Solution field (so that interpolation goldens do not depend on a solve)
        # (hash_field(n_nodes, 5, 1): recomputed by the tests, not stored)
        r.set_nodal(hash_field(n_nodes, 5, 1))
        lo = m["nodes"].min(0); hi = m["nodes"].max(0)
        pts = [rng.uniform(lo, hi, size=(N_RAND, 3))]
        if "surf_atoms" in m:
            # consecutive surface atoms exercise the chained guess; strided atoms the fallback scan
            pts = [m["surf_atoms"][:N_SURF], m["atoms"][:: max(1, len(m["atoms"]) // N_ATOM)]] + pts
        else:
            # points just above the surface triangles (centroid + small normal offset)
            tri = m["tris"][: N_SURF + N_ATOM]
            c = m["nodes"][tri].mean(1) + 0.3 * m["tri_norms"][: len(tri)] * rng.uniform(-1, 1, size=(len(tri), 1))
            pts = [c] + pts
        pts = np.ascontiguousarray(np.vstack(pts))
        g["points"] = pts
        for dim in (2, 3):
            for rank in (1, 2, 3):
                cells, sol = r.locate_interpolate(dim, rank, pts)
                g["cells_d%dr%d" % (dim, rank)] = cells
                g["sol_d%dr%d" % (dim, rank)] = sol
        n_cells = int((m["hex_markers"] > 0).sum())
        guess = rng.integers(0, n_cells, size=len(pts)).astype(np.int32)
        g["pic_guess"] = guess
        g["pic_cells"] = r.particle_cells(pts, guess)
        g["pic_cells2"] = r.particle_cells(pts, np.maximum(g["pic_cells"], 0))   # steady state: good guess
        ok = g["pic_cells"] >= 0
        g["pic_ok"] = ok
        g["pic_field"] = r.particle_field(pts[ok], g["pic_cells"][ok])
        g["pic_weights"] = r.particle_weights(pts[ok], g["pic_cells"][ok])
        # Interpolator::extract_solution on a seeded vertex potential
        n_vert = int((m["node_femocs2deal"] >= 0).sum())
        # (phi = hash_field(n_vert, 1, 2); only E of the vacuum nodes is stored, and for the smoothed
        # variant only the rows that smoothing touches: the tet nodes owning a pseudo-Voronoi cell)
        phi = hash_field(n_vert, 1, 2)[:, 0]
        vac = m["node_femocs2deal"] >= 0
        ne = r.extract_solution(phi, np.zeros(n_vert), 0, n_nodes)
        assert np.all(ne[~vac] == 0) and np.all(ne[vac, 3] == 0) and np.all(ne[vac, 4] == phi)
        g["extract_E"] = ne[vac, :3]
        ns = r.extract_solution(phi, np.zeros(n_vert), 1, n_nodes)
        n_voro = len(m["voro_off"]) - 1
        assert np.all(ns[n_voro:] == ne[n_voro:]) and np.all(ns[:, 3:] == ne[:, 3:])
        g["extract_E_smooth_tetnodes"] = ns[:n_voro, :3]
        np.savez_compressed(os.path.join(OUT, "interp_%s.npz" % name), **g)
        print(name, "nodes", n_nodes, "points", len(pts),
              {k: os.path.getsize(os.path.join(OUT, k + "_%s.npz" % name)) for k in ("mesh", "interp")})


if __name__ == "__main__":
    main()
