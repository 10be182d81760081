"""TEST INFRASTRUCTURE -- CPU restatement of the PIC particle push of the reference (numpy).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product (femocs_b200/) never does.

Follows, line by line:
  * Pic<3>::update_positions / update_position      src/Pic.cpp:137-184
  * periodic_image                                   src/Macros.cpp:41-48
  * Pic<3>::update_point_cell                        src/Pic.cpp:186-196
  * ParticleSpecies::clear_lost                      src/ParticleSpecies.cpp:16-31
  * Pic<3>::update_velocities                        src/Pic.cpp:198-209

The cell search and the gradient look-up are delegated to a `locator` with the methods
``particle_cells(xyz, guess)`` and ``particle_field(xyz, cells)``: either the C++ oracle
(oracle.oracle.Oracle) or the reference's own compiled LinearHexahedra (oracle.reflib.RefLib), which is how
tests/test_oracle_pic.py pins this restatement.  All arithmetic is separate IEEE multiply / add, as the
reference's Vec3 operators compile for baseline x86-64 (no FMA contraction).
"""
import numpy as np


def periodic_image(p, pmax, pmin):
    """src/Macros.cpp:41-48, vectorised"""
    p = np.asarray(p, np.float64)
    from_max = p - pmax
    from_min = p - pmin
    out = p.copy()
    hi = from_max > 0
    lo = (~hi) & (from_min < 0)
    out[hi] = pmin + from_max[hi]
    out[lo] = pmax + from_min[lo]
    return out


def update_positions(locator, pos, vel, cells, dt, box, periodic=True):
    """Returns (pos, vel, cells, n_lost) after one Pic::update_positions; box = (xmin, xmax, ymin, ymax, zmin, zmax)."""
    pos = np.array(pos, np.float64, copy=True).reshape(-1, 3)
    vel = np.array(vel, np.float64, copy=True).reshape(-1, 3)
    cells = np.array(cells, np.int32, copy=True)
    xmin, xmax, ymin, ymax, _, zmax = [float(b) for b in box]
    pos = pos + vel * dt                                           # :155  electron.pos += electron.vel * data.dt
    b1 = np.ones(len(pos), bool); b2 = np.ones(len(pos), bool)
    b3 = pos[:, 2] < zmax                                          # :159
    if periodic:                                                   # :161-164
        pos[:, 0] = periodic_image(pos[:, 0], xmax, xmin)
        pos[:, 1] = periodic_image(pos[:, 1], ymax, ymin)
    else:                                                          # :166-168
        b1 = (pos[:, 0] > xmin) & (pos[:, 0] < xmax)
        b2 = (pos[:, 1] > ymin) & (pos[:, 1] < ymax)
    inside = b1 & b2 & b3
    new_cells = np.full(len(pos), -1, np.int32)                    # :177-180
    if inside.any():
        new_cells[inside] = locator.particle_cells(pos[inside], cells[inside])
    keep = new_cells != -1                                         # clear_lost: stable removal
    return pos[keep], vel[keep], new_cells[keep], int((~keep).sum())


def update_velocities(locator, pos, vel, cells, dt, q_over_m):
    """vel += E * (dt * q_over_m), E = linhex.interp_gradient(pos, deal2femocs(cell))   (:198-209)"""
    vel = np.array(vel, np.float64, copy=True).reshape(-1, 3)
    if len(vel) == 0:
        return vel
    E = locator.particle_field(pos, cells)
    return vel + E * (dt * q_over_m)
