"""FE_Q(2) variant of the oracle's solver half (north star: "Q1/Q2 stiffness assembly", config 2 "Q2 Laplace solve").
The reference fixes shape_degree = 1 at compile time (include/DealSolver.h:130-131), so there is no reference output for
Q2 at all: the oracle restates what the same call sites do with FE_Q(2) / QGauss(3) in deal.II 9.2, and this file pins
that restatement's arithmetic to an INDEPENDENT derivation (numpy / scipy, [-1,1]^3 element in the UCD vertex order,
np.poly1d Lagrange basis, entities keyed by their vertex tuples, direct solve of the reduced system) that shares no code
with it, plus the properties the space must have (constants in the kernel, linear fields exact, closer to the refined
Q1 answer than Q1 is)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from femocs_b200 import synth
from oracle.oracle import Oracle
from test_oracle_independent import SU, SV, SW, FACES, _boundary_faces

L1 = [np.poly1d([0.5, -0.5, 0.0]), np.poly1d([-1.0, 0.0, 1.0]), np.poly1d([0.5, 0.5, 0.0])]      # nodes -1, 0, 1
XI = np.array([-1.0, 0.0, 1.0])


def _entities(hexs):
    """(n, 27, 8) vertex tuples (sorted, padded with -1) of the entity every local node (i, j, k) sits on"""
    n = len(hexs)
    ent = np.full((n, 27, 8), -1, np.int64)
    for i in range(3):
        for j in range(3):
            for k in range(3):
                on = np.ones(8, bool)
                for s, idx in ((SU, i), (SV, j), (SW, k)):
                    if idx != 1:
                        on &= s == XI[idx]
                v = np.sort(hexs[:, on], axis=1)
                ent[:, i + 3 * j + 9 * k, :v.shape[1]] = v
    return ent


def _q2_system(m, field, anode_potential=None):
    nodes = m["nodes"]; hexs = m["hexs"][m["hex_markers"] > 0].astype(np.int64)
    n = len(hexs)
    ent = _entities(hexs)
    uniq, dof = np.unique(ent.reshape(-1, 8), axis=0, return_inverse=True)
    dof = dof.reshape(n, 27); nd = len(uniq)
    X = nodes[hexs]
    gp, gw = np.polynomial.legendre.leggauss(3)
    Ke = np.zeros((n, 27, 27))
    for a, wa in zip(gp, gw):
        for b, wb in zip(gp, gw):
            for c, wc in zip(gp, gw):
                dN = np.stack([SU * (1 + SV * b) * (1 + SW * c), (1 + SU * a) * SV * (1 + SW * c), (1 + SU * a) * (1 + SV * b) * SW], 1) / 8.0
                J = np.einsum("nkd,ke->nde", X, dN)
                det = np.abs(np.linalg.det(J))
                ref = np.zeros((27, 3))
                for i in range(3):
                    for j in range(3):
                        for k in range(3):
                            ref[i + 3 * j + 9 * k] = [L1[i].deriv()(a) * L1[j](b) * L1[k](c), L1[i](a) * L1[j].deriv()(b) * L1[k](c),
                                                      L1[i](a) * L1[j](b) * L1[k].deriv()(c)]
                G = np.einsum("ke,ned->nkd", ref, np.linalg.inv(J))
                Ke += (wa * wb * wc) * det[:, None, None] * np.einsum("nid,njd->nij", G, G)
    rows = np.repeat(dof, 27, axis=1).reshape(-1); cols = np.tile(dof, (1, 27)).reshape(-1)
    K = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(nd, nd)).tocsr()
    # boundary faces and their 9 dofs: the local nodes whose entity lies inside the face's vertex set
    bf = _boundary_faces(hexs)
    ctr = nodes[bf].mean(1)
    mn, mx = ctr.min(0), ctr.max(0); eps = 1e-6
    side = (np.abs(ctr[:, 0] - mn[0]) <= eps) | (np.abs(ctr[:, 0] - mx[0]) <= eps) | (np.abs(ctr[:, 1] - mn[1]) <= eps) | (np.abs(ctr[:, 1] - mx[1]) <= eps)
    top = ~side & (np.abs(ctr[:, 2] - mx[2]) <= eps)
    copper = ~side & ~top
    key_of = {tuple(u): d for d, u in enumerate(uniq)}
    def face_dofs(f4):      # corners cyclic p0 p1 p2 p3 -> 3 x 3 grid (s, t) in {-1, 0, 1}^2
        grid = {(-1, -1): (f4[0],), (1, -1): (f4[1],), (1, 1): (f4[2],), (-1, 1): (f4[3],), (0, -1): (f4[0], f4[1]), (1, 0): (f4[1], f4[2]),
                (0, 1): (f4[2], f4[3]), (-1, 0): (f4[3], f4[0]), (0, 0): tuple(f4)}
        out = {}
        for st, vs in grid.items():
            k = sorted(int(v) for v in vs); k += [-1] * (8 - len(k))
            out[st] = key_of[tuple(k)]
        return out
    b = np.zeros(nd)
    fixed = np.zeros(nd, bool); phi = np.zeros(nd)
    for f4 in bf[copper]:
        fixed[list(face_dofs(f4).values())] = True
    s4 = np.array([-1, 1, 1, -1.0]); t4 = np.array([-1, -1, 1, 1.0])
    for f4 in bf[top]:
        fd = face_dofs(f4)
        if anode_potential is not None:
            for d in fd.values(): fixed[d] = True; phi[d] = anode_potential
            continue
        P = nodes[f4]
        for s, ws in zip(gp, gw):
            for t, wt in zip(gp, gw):
                ds = (s4 * (1 + t4 * t) / 4.0) @ P; dt = ((1 + s4 * s) * t4 / 4.0) @ P
                dA = np.linalg.norm(np.cross(ds, dt))
                for (si, ti), d in fd.items():
                    b[d] += ws * wt * field * dA * L1[si + 1](s) * L1[ti + 1](t)
    free = ~fixed
    phi[free] = spla.spsolve(K[free][:, free].tocsc(), b[free] - K[free][:, fixed] @ phi[fixed])
    vert_dof = {int(u[0]): d for d, u in enumerate(uniq) if u[1] < 0}
    return K, phi, vert_dof, nd


def _bump_mesh(nx=7, ny=7, nz=8, jitter=0.2):
    """jittered box whose bottom centre is NOT vacuum (a 3 x 3 x 4 block of copper): the walls of the block are 'other'
    boundary faces = copper_surface (DealSolver.cpp:460-518), so the field is not uniform"""
    nodes, hexs, mk = synth.box_mesh(nx, ny, nz, 3.5, 3.5, 4.0, jitter=jitter)
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    hole = ((np.abs(I - nx // 2) <= 1) & (np.abs(J - ny // 2) <= 1) & (K < nz // 2)).ravel()
    mk = mk.copy(); mk[hole] = -1
    return dict(nodes=nodes, hexs=hexs, hex_markers=mk)


def _oracle_q2(m):
    o = Oracle(); o.set_fe_degree(2); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    return o


def test_q2_dof_count_and_numbering():
    nodes, hexs, mk = synth.box_mesh(4, 3, 5, 3.0, 2.5, 4.0, jitter=0.2)
    o = _oracle_q2(dict(nodes=nodes, hexs=hexs, hex_markers=mk))
    assert o.n_dofs == 9 * 7 * 11                              # (2 nx + 1)(2 ny + 1)(2 nz + 1) support points
    cd = o.cell_dofs27()
    assert cd.min() == 0 and cd.max() == o.n_dofs - 1 and len(np.unique(cd)) == o.n_dofs
    # deal.II first touch: cell 0 numbers its 8 vertices 0..7, 12 lines 8..19, 6 quads 20..25, the interior 26
    corner = [0, 2, 6, 8, 18, 20, 24, 26]
    assert cd[0, corner].tolist() == list(range(8)) and cd[0, 13] == 26
    assert sorted(cd[0, [12, 14, 10, 16, 4, 22]].tolist()) == list(range(20, 26))
    _, _, v2d, _ = o.vectors()
    assert np.array_equal(np.sort(v2d), np.sort(np.unique(cd[:, corner])))


@pytest.mark.parametrize("jitter", [0.0, 0.25])
def test_q2_invariants_and_linear_exactness(jitter):
    nodes, hexs, mk = synth.box_mesh(5, 4, 6, 3.0, 2.5, 4.0, jitter=jitter)
    o = _oracle_q2(dict(nodes=nodes, hexs=hexs, hex_markers=mk))
    F = 0.37
    o.setup(F, 0.0, False); o.assemble(True)
    rp, col, val, save = o.csr()
    K = sp.csr_matrix((save, col, rp))
    assert abs(K - K.T).max() < 1e-13 and np.abs(np.asarray(K.sum(1))).max() < 1e-12
    assert o.solve(10000, 1e-12, 1.2, 0) > 0
    _, _, _, v2n = o.vectors()
    assert np.abs(o.export_solution() - F * nodes[v2n, 2]).max() < 1e-10
    o.setup(0.0, 5.0, True); o.assemble(True)
    assert o.solve(10000, 1e-12, 1.2, 0) > 0
    assert np.abs(o.export_solution() - 5.0 * nodes[v2n, 2] / 4.0).max() < 1e-10


def test_q2_oracle_matches_an_independent_derivation():
    m = _bump_mesh()
    field = 0.5
    K, phi, vert_dof, nd = _q2_system(m, field)
    o = _oracle_q2(m)
    assert o.n_dofs == nd
    o.setup(field, 0.0, False); o.assemble(True)
    assert o.solve(20000, 1e-12, 1.2, 0) > 0
    _, _, v2d, v2n = o.vectors()
    ref = o.export_solution()
    mine = np.array([phi[vert_dof[int(nn)]] for nn in v2n])
    assert np.abs(ref - mine).max() <= 1e-9 * np.abs(ref).max()
    # the matrices agree entry by entry once both are brought to a common numbering (vertex dofs suffice to anchor it:
    # compare the vertex-vertex block)
    rp, col, val, save = o.csr()
    Ko = sp.csr_matrix((save, col, rp))[v2d][:, v2d]
    idx = np.array([vert_dof[int(nn)] for nn in v2n])
    assert abs(Ko - K[idx][:, idx]).max() <= 1e-12 * abs(K).max()
    _, phi_d, _, _ = _q2_system(m, 0.0, anode_potential=7.5)
    o.setup(0.0, 7.5, True); o.assemble(True)
    assert o.solve(20000, 1e-12, 1.2, 0) > 0
    ref = o.export_solution()
    assert np.abs(ref - np.array([phi_d[vert_dof[int(nn)]] for nn in v2n])).max() <= 1e-9 * np.abs(ref).max()


def test_q2_is_closer_to_the_refined_answer_than_q1():
    """on the bump mesh (re-entrant copper edges: a real field enhancement) the Q2 potential at the vertices lies closer to the Q1 potential of the once-refined mesh
    than the Q1 potential of the same mesh does (the point of asking for Q2)"""
    m = _bump_mesh(jitter=0.0)
    nodes0, hexs0, mk0 = synth.refine_vacuum(m["nodes"], m["hexs"], m["hex_markers"], 0)
    nodes1, hexs1 = synth.refine_hexes(nodes0, hexs0)
    mk1 = np.ones(len(hexs1), np.int32)
    sols = []
    for deg, nn, hh, mm in ((1, nodes0, hexs0, mk0), (2, nodes0, hexs0, mk0), (1, nodes1, hexs1, mk1)):
        o = Oracle(); o.set_fe_degree(deg); o.import_mesh(nn, hh, mm)
        o.setup(0.5, 0.0, False); o.assemble(True)
        assert o.solve(20000, 1e-11, 1.2, 0) > 0
        _, _, _, v2n = o.vectors()
        full = np.full(len(nn), np.nan); full[v2n] = o.export_solution()
        sols.append(full)
    used = ~np.isnan(sols[0])
    fine = sols[2][:len(nodes0)][used]                         # refinement keeps the coarse nodes first
    e1 = np.abs(sols[0][used] - fine).max(); e2 = np.abs(sols[1][used] - fine).max()
    assert e2 < e1


def test_q2_space_charge_weights_by_construction():
    """PoissonSolver.cpp:276-296 with FE_Q(2): particles are PLACED at known unit-cell coordinates of known cells (trilinear
    image), so the right-hand side they must add is the scatter of the 27 tensor Lagrange values at those coordinates --
    no inverse map involved on this side.  Checks the oracle's inverse map, its axis conventions and the lattice order."""
    m = _bump_mesh(jitter=0.2)
    o = _oracle_q2(m)
    rng = np.random.default_rng(9)
    n = 400
    cells = rng.integers(0, o.n_cells, n).astype(np.int32)
    xi = rng.uniform(0.02, 0.98, (n, 3))
    v2n = o.vectors()[3]
    X = m["nodes"][v2n[o.cells()[cells]]]                                   # (n, 8, 3) lexicographic vertices
    w8 = np.stack([np.where((v >> 0) & 1, xi[:, 0], 1 - xi[:, 0]) * np.where((v >> 1) & 1, xi[:, 1], 1 - xi[:, 1])
                   * np.where((v >> 2) & 1, xi[:, 2], 1 - xi[:, 2]) for v in range(8)], 1)
    pts = np.einsum("nv,nvd->nd", w8, X)
    cf = -1.7
    o.setup(0.4, 0.0, False); o.assemble(True)
    rhs0 = o.vectors()[0].copy()
    o.setup(0.4, 0.0, False); o.assemble(True, pts, cells, cf)
    rhs1 = o.vectors()[0]
    lag = lambda x: np.stack([2 * (x - 0.5) * (x - 1), 4 * x * (1 - x), 2 * x * (x - 0.5)], 1)      # (n, 3)
    Lx, Ly, Lz = lag(xi[:, 0]), lag(xi[:, 1]), lag(xi[:, 2])
    w27 = np.einsum("ni,nj,nk->nkji", Lx, Ly, Lz).reshape(n, 27)             # index i + 3 j + 9 k
    assert np.abs(w27.sum(1) - 1).max() < 1e-13
    expect = np.zeros(o.n_dofs)
    np.add.at(expect, o.cell_dofs27()[cells].reshape(-1), (cf * w27).reshape(-1))
    rp, col, val, _ = o.csr()
    constrained = np.array([np.count_nonzero(val[rp[r]:rp[r + 1]]) == 1 for r in range(o.n_dofs)])
    free = ~constrained
    assert np.abs((rhs1 - rhs0)[free] - expect[free]).max() <= 1e-10 * np.abs(expect).max()
    assert np.array_equal(rhs1[constrained], rhs0[constrained])
