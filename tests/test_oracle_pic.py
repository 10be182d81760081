"""CPU tests of the PIC push restatement (oracle/pic.py): the committed golden vectors were produced with the
REFERENCE's own compiled cell search / gradient code (oracle/make_golden_pic.py); here the same push runs with the
C++ oracle as locator and must reproduce them bit for bit."""
import numpy as np
import pytest

from oracle import pic
from oracle.fields import hash_field
from oracle.oracle import Oracle


def test_periodic_image_known_answers():
    # src/Macros.cpp:41-48: one box length is added/subtracted once, points inside are untouched
    p = np.array([-1.5, 0.0, 0.25, 1.0, 1.75, 3.5])
    assert np.array_equal(pic.periodic_image(p, 1.0, 0.0), np.array([-0.5, 0.0, 0.25, 1.0, 0.75, 2.5]))


class _FakeLocator:
    """every particle keeps its cell unless x > 0.5 (then it is 'in the material': -1)"""
    def particle_cells(self, xyz, guess):
        out = np.asarray(guess, np.int32).copy()
        out[np.asarray(xyz)[:, 0] > 0.5] = -1
        return out

    def particle_field(self, xyz, cells):
        return np.tile(np.array([1.0, 2.0, 3.0]), (len(xyz), 1))


def test_update_positions_losses_and_order():
    pos = np.array([[0.1, 0.1, 0.1], [0.45, 0.2, 0.2], [0.2, 0.95, 0.3], [0.3, 0.3, 0.95], [0.05, 0.5, 0.5]])
    vel = np.array([[0.1, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1], [-0.1, 0, 0]])
    cells = np.arange(5, dtype=np.int32)
    box = (0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    # periodic: particle 1 moves into the 'material' (lost), 2 wraps in y, 3 leaves through the top (lost), 4 wraps in x to 0.95 (lost)
    p, v, c, lost = pic.update_positions(_FakeLocator(), pos, vel, cells, 1.0, box, True)
    assert lost == 3 and np.array_equal(c, [0, 2])                       # survivors keep their order (clear_lost)
    assert np.allclose(p, [[0.2, 0.1, 0.1], [0.2, 0.05, 0.3]]) and np.array_equal(v, vel[[0, 2]])
    # non-periodic: 2 and 4 leave the x/y box as well
    p, v, c, lost = pic.update_positions(_FakeLocator(), pos, vel, cells, 1.0, box, False)
    assert lost == 4 and np.array_equal(c, [0])
    # empty set
    p, v, c, lost = pic.update_positions(_FakeLocator(), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.int32), 1.0, box, True)
    assert lost == 0 and len(c) == 0
    assert np.array_equal(pic.update_velocities(_FakeLocator(), pos[:2], vel[:2], cells[:2], 0.5, -2.0), vel[:2] + np.array([1.0, 2.0, 3.0]) * (0.5 * -2.0))


@pytest.mark.parametrize("name", ["hemicone", "mdsmall"])
def test_oracle_push_reproduces_reference_golden(name, golden):
    m = golden("mesh", name); g = golden("picpush", name)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    o.set_nodal(hash_field(len(m["nodes"]), 5, 1))
    dt = float(g["dt"][0]); qm = float(g["q_over_m"][0])
    for periodic in (1, 0):
        pos, vel, cells = g["pos0"], g["vel0"], g["cells0"]
        for step in range(3):
            tag = "p%d_s%d_" % (periodic, step)
            pos, vel, cells, lost = pic.update_positions(o, pos, vel, cells, dt, g["box"], bool(periodic))
            vel = pic.update_velocities(o, pos, vel, cells, dt, qm)
            assert lost == int(g[tag + "lost"][0])
            assert np.array_equal(cells, g[tag + "cells"]) and np.array_equal(pos, g[tag + "pos"]) and np.array_equal(vel, g[tag + "vel"])
