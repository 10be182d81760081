"""The CPU oracle's interpolation half against golden vectors produced by the reference's OWN
code (verbatim InterpolatorCells / Interpolator / SolutionReader compiled from /root/reference,
see oracle/make_golden.py).  Bit-exact: cell indices AND floating point results."""
import numpy as np
import pytest

from oracle.fields import hash_field
from oracle.oracle import Oracle

MESHES = ["hemicone", "mdsmall", "mdbig"]


@pytest.fixture(scope="module")
def oracles(golden):
    out = {}
    for name in MESHES:
        m = golden("mesh", name)
        o = Oracle()
        o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        o.interp_initialize(m)
        out[name] = o
    return out


@pytest.mark.parametrize("name", MESHES)
def test_index_maps(name, golden, oracles):
    m = golden("mesh", name); o = oracles[name]
    _, _, v2d, v2n = o.vectors()
    n2v = np.full(len(m["nodes"]), -1, np.int32); n2v[v2n] = np.arange(len(v2n))
    assert np.array_equal(n2v, m["node_femocs2deal"])            # InterpolatorCells.cpp:38-66
    assert o.n_cells == int((m["hex_markers"] > 0).sum())
    assert sorted(v2d.tolist()) == list(range(o.n_dofs))


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("rank", [1, 2, 3])
def test_locate_interpolate(name, dim, rank, golden, oracles):
    g = golden("interp", name); o = oracles[name]
    o.set_nodal(hash_field(o.n_nodes, 5, 1))
    cells, sol = o.locate_interpolate(dim, rank, g["points"])
    assert np.array_equal(cells, g["cells_d%dr%d" % (dim, rank)])
    assert np.array_equal(sol, g["sol_d%dr%d" % (dim, rank)])
    # re-interpolation with cached cells (SolutionReader.cpp:167-190) gives the same values
    assert np.array_equal(o.interpolate(dim, rank, g["points"], cells), sol)


@pytest.mark.parametrize("name", MESHES)
def test_particles(name, golden, oracles):
    g = golden("interp", name); o = oracles[name]
    o.set_nodal(hash_field(o.n_nodes, 5, 1))
    pc = o.particle_cells(g["points"], g["pic_guess"])
    assert np.array_equal(pc, g["pic_cells"])
    assert np.array_equal(o.particle_cells(g["points"], np.maximum(pc, 0)), g["pic_cells2"])
    ok = g["pic_ok"]
    assert np.array_equal(o.particle_field(g["points"][ok], pc[ok]), g["pic_field"])
    assert np.array_equal(o.particle_weights(g["points"][ok], pc[ok]), g["pic_weights"])


@pytest.mark.parametrize("name", MESHES)
def test_extract_solution(name, golden, oracles):
    m = golden("mesh", name); g = golden("interp", name); o = oracles[name]
    _, _, v2d, _ = o.vectors()
    phi = hash_field(o.n_vertices, 1, 2)[:, 0]
    sol = np.zeros(o.n_dofs); sol[v2d] = phi
    o.set_solution(sol)
    vac = m["node_femocs2deal"] >= 0
    nod = o.extract_solution(False)
    assert np.all(nod[~vac] == 0)
    assert np.array_equal(nod[vac, 4], phi) and np.all(nod[vac, 3] == 0)
    assert np.array_equal(nod[vac, :3], g["extract_E"])
    n_voro = len(m["voro_off"]) - 1
    nod_s = o.extract_solution(True)
    assert np.array_equal(nod_s[:n_voro, :3], g["extract_E_smooth_tetnodes"])
    assert np.array_equal(nod_s[n_voro:], nod[n_voro:])
