"""Host-side mesh import (femocs_b200/csrc/host_setup.cpp through the host-only plan entry points) against the
oracle and against itself on edge cases.  No GPU needed."""
import numpy as np
import pytest

from femocs_b200 import PartitionPlan, synth
from oracle.oracle import Oracle


def _plan(nodes, hexs, mk):
    return PartitionPlan(0, 1).import_whole(nodes, hexs, mk)


def test_all_negative_mesh_is_flipped_like_deal_ii():
    """GridReordering::invert_all_cells_of_negative_grid (DealSolver.cpp:198): every cell of a grid with negative
    measure has its old-style vertices i and i + 4 swapped.  Swapping them in the INPUT of a positive mesh therefore
    yields, after the import, exactly the cells of the positive mesh: same dofs per cell, same sparsity."""
    nodes, hexs, mk = synth.box_mesh(4, 3, 5, 2.0, 1.5, 2.5, jitter=0.15)
    neg = hexs.copy()
    neg[:, :4], neg[:, 4:] = hexs[:, 4:], hexs[:, :4]
    a = _plan(nodes, hexs, mk); b = _plan(nodes, neg, mk)
    assert np.array_equal(a.cells_dof, b.cells_dof)
    assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col)
    assert np.array_equal(a.copper, b.copper) and np.array_equal(a.top, b.top)
    # the oracle restates the same rule
    o1 = Oracle(); o1.import_mesh(nodes, hexs, mk)
    o2 = Oracle(); o2.import_mesh(nodes, neg, mk)
    assert np.array_equal(o1.cells(), o2.cells())
    v2d = o1.vectors()[2]
    assert np.array_equal(v2d[o1.cells()], a.cells_dof)


def test_mixed_orientation_is_refused():
    nodes, hexs, mk = synth.box_mesh(3, 3, 3)
    bad = hexs.copy()
    bad[0, :4], bad[0, 4:] = hexs[0, 4:], hexs[0, :4]
    from femocs_b200 import FemocsB200Error
    with pytest.raises(FemocsB200Error):
        _plan(nodes, bad, mk)
    with pytest.raises(RuntimeError):
        Oracle().import_mesh(nodes, bad, mk)


@pytest.mark.parametrize("name", ["hemicone", "mdsmall"])
def test_bulk_mesh_import_matches_oracle(name, golden):
    """fb_import_bulk_mesh's host half (Hexahedra::export_bulk selection, CurrentHeatSolver::mark_mesh boundary roles,
    numbering, sparsity, face order of export_surface_centroids) against the oracle restatement"""
    m = golden("mesh", name)
    o = Oracle(); o.import_bulk_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    a = PartitionPlan(0, 1).import_whole(m["nodes"], m["hexs"], m["hex_markers"], bulk=True)
    assert (a.n_rows, a.n_cells, a.nnz) == (o.n_dofs, o.n_cells, o.nnz)
    assert a.n_cells == int((m["hex_markers"] < 0).sum())
    rp, col, _, _ = o.csr()
    assert np.array_equal(a.rowptr, rp) and np.array_equal(a.col, col)
    v2d = o.vectors()[2]
    assert np.array_equal(v2d[o.cells()], a.cells_dof)
    # Dirichlet set = dofs of the copper_bottom faces (BoundaryID 7); no "top" role on the bulk mesh
    cell, face, bid = o.bfaces()
    FV = np.array([[0, 2, 4, 6], [1, 3, 5, 7], [0, 1, 4, 5], [2, 3, 6, 7], [0, 1, 2, 3], [4, 5, 6, 7]])
    bottom = np.zeros(o.n_dofs, np.int32)
    sel = bid == 7
    bottom[a.cells_dof[cell[sel]][np.arange(sel.sum())[:, None], FV[face[sel]]].ravel()] = 1
    assert sel.sum() > 0 and np.array_equal(a.copper, bottom) and a.top.sum() == 0
    cen = a.surface_centroids()
    assert np.array_equal(cen, o.surface_centroids()) and len(cen) == int((bid == 2).sum()) == len(m["quads"])
    # the vacuum import of the same arrays is untouched by the bulk option
    b = PartitionPlan(0, 1).import_whole(m["nodes"], m["hexs"], m["hex_markers"])
    assert b.n_cells == int((m["hex_markers"] > 0).sum())


def _interior_jitter(nodes, scale, seed=3):
    """moves every node that does not lie on one of the six bounding planes (the faces there keep their boundary ids)"""
    rng = np.random.default_rng(seed)
    lo, hi = nodes.min(0), nodes.max(0)
    inside = np.all((nodes > lo + 1e-9) & (nodes < hi - 1e-9), axis=1)
    out = nodes.copy()
    out[inside] += scale * rng.standard_normal((int(inside.sum()), 3))
    return out


def test_unchanged_topology_is_reused(golden):
    """a second import with identical connectivity only refreshes the geometry -- and must leave the context
    exactly as a fresh import of the moved mesh would (numbering, sparsity, Dirichlet sets, face order); anything that
    changes the integer side (connectivity, a boundary face leaving its plane) takes the full path"""
    m = golden("mesh", "mdsmall")
    nodes, hexs, mk = m["nodes"], m["hexs"], m["hex_markers"]
    p = PartitionPlan(0, 1)
    p.import_whole(nodes, hexs, mk)
    assert not p.reused
    cen0 = p.surface_centroids()
    moved = _interior_jitter(nodes, 0.02)
    p.import_whole(moved, hexs, mk)
    assert p.reused
    f = PartitionPlan(0, 1).import_whole(moved, hexs, mk)
    for key in ("rowptr", "col", "cells_dof", "copper", "top"):
        assert np.array_equal(getattr(p, key), getattr(f, key)), key
    assert np.array_equal(p.surface_centroids(), f.surface_centroids())          # the refreshed coordinates are in use
    assert not np.array_equal(p.surface_centroids(), cen0)
    # the bulk mesh of the same arrays is another mesh kind: full import
    p.import_whole(moved, hexs, mk, bulk=True)
    assert not p.reused
    # permuted hexahedra: full import
    q = PartitionPlan(0, 1); q.import_whole(nodes, hexs, mk)
    h2 = hexs.copy(); h2[[0, 1]] = h2[[1, 0]]; k2 = mk.copy(); k2[[0, 1]] = k2[[1, 0]]
    q.import_whole(nodes, h2, k2)
    assert not q.reused
    # a node of the top plane pushed outwards: its faces define a new zmax, the others lose the "top" id -> full import,
    # same result as a fresh context
    top = np.flatnonzero(nodes[:, 2] > nodes[:, 2].max() - 1e-9)
    bumped = nodes.copy(); bumped[top[0], 2] += 0.5
    q.import_whole(bumped, h2, k2)
    assert not q.reused
    g = PartitionPlan(0, 1).import_whole(bumped, h2, k2)
    assert np.array_equal(q.top, g.top) and np.array_equal(q.copper, g.copper)
    # option mesh_reuse = 0 is honoured by the full library only (plan contexts have no options): covered on the GPU


def test_parallel_numbering_on_a_large_mesh_matches_oracle():
    """meshes with more than 2e6 cell vertices take the parallel paths of the host import (atomic counting sort of the
    vertex -> cells adjacency, first-touch numbering as a scan over first-occurrence flags): the result must be the serial
    deal.II numbering the oracle restates"""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    nodes, hexs, mk = bench.load_x_mesh(0)                 # 350 424 hexahedra
    p = PartitionPlan(0, 1).import_whole(nodes, hexs, mk)
    o = Oracle(); o.import_mesh(nodes, hexs, mk)
    rp, col, _, _ = o.csr()
    assert np.array_equal(p.rowptr, rp) and np.array_equal(p.col, col)
    assert np.array_equal(o.vectors()[2][o.cells()], p.cells_dof)


def _q2_face_nodes(f):
    axis, fixed = f // 2, (f % 2) * 2
    stride = [1, 3, 9]
    a0 = 1 if axis == 0 else 0; a1 = 1 if axis == 2 else 2
    return [fixed * stride[axis] + a * stride[a0] + b * stride[a1] for b in range(3) for a in range(3)]


@pytest.mark.parametrize("name", ["box", "hemicone"])
def test_q2_import_matches_oracle(name, golden):
    """option fe_degree = 2 (q2.cu, host half): first-touch numbering of vertex / line / quad / hex dofs in deal.II's
    per-cell order, sparsity and the Dirichlet candidates, against the oracle's FE_Q(2) restatement -- index work, bit-exact"""
    if name == "box":
        nodes, hexs, mk = synth.box_mesh(5, 4, 6, 2.0, 1.5, 2.5, jitter=0.15)
    else:
        m = golden("mesh", name); nodes, hexs, mk = m["nodes"], m["hexs"], m["hex_markers"]
    o = Oracle(); o.set_fe_degree(2); o.import_mesh(nodes, hexs, mk)
    a = PartitionPlan(0, 1).import_whole(nodes, hexs, mk, fe_degree=2)
    assert (a.n_rows, a.n_cells, a.nnz) == (o.n_dofs, o.n_cells, o.nnz)
    rp, col, _, _ = o.csr()
    assert np.array_equal(a.rowptr, rp) and np.array_equal(a.col, col)
    cd = o.cell_dofs27()
    assert np.array_equal(a.cells27(), cd)
    v2d = o.vectors()[2]
    assert np.array_equal(v2d[o.cells()], a.cells_dof)                    # the 8 vertex dofs per cell (geometry look-ups)
    cell, face, bid = o.bfaces()
    for flag, want in ((a.copper, 2), (a.top, 8)):
        ref = np.zeros(o.n_dofs, np.int32)
        for c_, f_ in zip(cell[bid == want], face[bid == want]):
            ref[cd[c_, _q2_face_nodes(f_)]] = 1
        assert ref.sum() > 0 and np.array_equal(flag, ref)
    # an FE_Q(1) import on the same plan afterwards is the ordinary one
    b = PartitionPlan(0, 1).import_whole(nodes, hexs, mk)
    o1 = Oracle(); o1.import_mesh(nodes, hexs, mk)
    assert b.n_rows == o1.n_dofs and b.nnz == o1.nnz
