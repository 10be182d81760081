"""TEST/BENCH INFRASTRUCTURE -- regenerates tests/golden/bench_x90.npz and mesh_tip110.npz from the reference's own mesher.

Run in the build container (needs /root/reference and oracle/_ref/libfemocs_ref.so):

    make -f oracle/Makefile.ref && python oracle/make_bench_mesh.py

The file holds the base mesh of BASELINE.json's config 4 ("extension_90nm.xyz synthetic mesh at
~2e7 DoF"): the vacuum hexahedra that the reference's verbatim AtomReader -> Surface -> TetGen ->
Tethex pipeline produces for in/apex.ckx extended by in/extension_90nm.xyz (Main.cpp:173-179
"extend" preset): 350 424 hexahedra on 407 018 vertices, node list compacted
to the vertices those hexahedra touch (order preserved).  bench.py refines it twice
(femocs_b200.synth.refine_hexes, 1 -> 8 trilinear split) to the 2.24e7-hexahedron / 2.3e7-DoF
workload; /root/reference is not needed at bench time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reflib import RefLib  # noqa: E402


def main():
    m = RefLib().generate("extend90")
    vac = m["hex_markers"] > 0
    hexs = m["hexs"][vac]
    used = np.unique(hexs)
    remap = np.full(len(m["nodes"]), -1, np.int64)
    remap[used] = np.arange(len(used))
    out = os.path.join(ROOT, "tests", "golden", "bench_x90.npz")
    np.savez_compressed(out, nodes=m["nodes"][used], hexs=remap[hexs].astype(np.int32))
    print(out, "hexes", len(hexs), "vertices", len(used), "%.1f MB" % (os.path.getsize(out) / 1e6))
    # config 5 ("ProjectHeat coupled Poisson + current/heat on tip110.ckx"): the complete mesh of the Main.cpp:216-220
    # preset (vacuum AND bulk hexahedra, tetrahedra, surface triangles / quadrangles, surface atoms); the full atom
    # list is not needed by the bench leg and is left out
    m = RefLib().generate("tip110")
    m.pop("atoms", None)
    out = os.path.join(ROOT, "tests", "golden", "mesh_tip110.npz")
    np.savez_compressed(out, **m)
    print(out, "vacuum hexes", int((m["hex_markers"] > 0).sum()), "bulk hexes", int((m["hex_markers"] < 0).sum()),
          "%.1f MB" % (os.path.getsize(out) / 1e6))


if __name__ == "__main__":
    main()
