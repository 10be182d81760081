"""TEST INFRASTRUCTURE -- regenerates tests/golden/picpush_<name>.npz: three PIC pushes
(Pic::update_positions -> Pic::update_velocities, src/Pic.cpp:137-209) of seeded super-particles on the
fixture meshes, with the cell search and the gradient look-up done by the REFERENCE's own compiled
LinearHexahedra (oracle/_ref) and the push arithmetic by oracle/pic.py.

    make -f oracle/Makefile.ref && python oracle/make_golden_pic.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pic  # noqa: E402
from oracle.fields import hash_field  # noqa: E402
from oracle.reflib import RefLib  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DT = 0.5            # fs; with |v| ~ 1 A/fs a particle crosses a cell every few steps
Q_OVER_M = -17.5882  # Pic.h:90
N_STEPS = 3


def scenario(m, interp_golden, seed):
    """initial particles: the located golden points (inside vacuum cells), seeded velocities; the box is the
    node bounding box shrunk in x/y so that periodic wrapping and (non-periodic) losses both happen"""
    rng = np.random.default_rng(seed)
    ok = interp_golden["pic_ok"]
    pos = interp_golden["points"][ok]; cells = interp_golden["pic_cells"][ok]
    rep = 3
    pos = np.repeat(pos, rep, axis=0); cells = np.repeat(cells, rep)
    vel = rng.normal(0.0, 4.0, size=pos.shape)
    lo = m["nodes"].min(0); hi = m["nodes"].max(0)
    box = np.array([lo[0] + 2.0, hi[0] - 2.0, lo[1] + 2.0, hi[1] - 2.0, lo[2], hi[2] - 3.0])
    return pos, vel, cells.astype(np.int32), box


def main():
    r = RefLib()
    for name in ("hemicone", "mdsmall"):
        m = r.generate(name)
        with np.load(os.path.join(OUT, "mesh_%s.npz" % name)) as z:
            assert np.array_equal(z["nodes"], m["nodes"]) and np.array_equal(z["hexs"], m["hexs"]), "mesh fixture out of date"
        with np.load(os.path.join(OUT, "interp_%s.npz" % name)) as z:
            ig = {k: z[k] for k in z.files}
        r.set_nodal(hash_field(len(m["nodes"]), 5, 1))
        pos0, vel0, cells0, box = scenario(m, ig, 99 + len(name))
        g = {"pos0": pos0, "vel0": vel0, "cells0": cells0, "box": box, "dt": np.array([DT]), "q_over_m": np.array([Q_OVER_M])}
        for periodic in (1, 0):
            pos, vel, cells = pos0, vel0, cells0
            for step in range(N_STEPS):
                pos, vel, cells, lost = pic.update_positions(r, pos, vel, cells, DT, box, bool(periodic))
                vel = pic.update_velocities(r, pos, vel, cells, DT, Q_OVER_M)
                tag = "p%d_s%d_" % (periodic, step)
                g[tag + "pos"] = pos; g[tag + "vel"] = vel; g[tag + "cells"] = cells; g[tag + "lost"] = np.array([lost])
                print(name, "periodic" if periodic else "box", "step", step, "particles", len(cells), "lost", lost)
        np.savez_compressed(os.path.join(OUT, "picpush_%s.npz" % name), **g)
        print(name, os.path.getsize(os.path.join(OUT, "picpush_%s.npz" % name)), "bytes")


if __name__ == "__main__":
    main()
