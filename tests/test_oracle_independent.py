"""An INDEPENDENT second derivation of the solver half, written from the published finite-element definitions only
(numpy / scipy; reference cube [-1,1]^3 with the FEMOCS/UCD vertex signs, Gauss-Legendre points from numpy, node-index
numbering, a direct sparse solve of the Dirichlet-reduced system), checked against the oracle's restatement of the
deal.II path (oracle/femocs_oracle.cpp: [0,1]^3 lexicographic element, first-touch DoF numbering, symmetric row/column
elimination, SSOR-CG).  The reference holds no golden vector at the deal.II boundary (parity unpinned);
this test pins the oracle's ARITHMETIC -- stiffness, Neumann load, elimination, CG -- to a formulation that shares no
code with it.  What stays unpinned is only the reading of deal.II's semantics (boundary ids, which faces carry which
condition), restated here from src/DealSolver.cpp:460-518 and src/PoissonSolver.cpp:52-55."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle.oracle import Oracle

SU = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0]); SV = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0]); SW = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]      # UCD hexahedron


def _stiffness(nodes, hexs):
    """K_e(i,j) = sum_q |J| (J^-T grad N_i).(J^-T grad N_j), 2x2x2 Gauss-Legendre on [-1,1]^3 (weights 1)"""
    gp, gw = np.polynomial.legendre.leggauss(2)
    X = nodes[hexs]                                           # (n, 8, 3)
    Ke = np.zeros((len(hexs), 8, 8)); vol = np.zeros(len(hexs))
    for a, wa in zip(gp, gw):
        for b, wb in zip(gp, gw):
            for c, wc in zip(gp, gw):
                dN = np.stack([SU * (1 + SV * b) * (1 + SW * c), (1 + SU * a) * SV * (1 + SW * c), (1 + SU * a) * (1 + SV * b) * SW], 1) / 8.0
                J = np.einsum("nkd,ke->nde", X, dN)           # J[d][e] = d x_d / d xi_e
                det = np.linalg.det(J)
                # the (u, v, w) axes of the UCD vertex signs are a reflection of deal.II's lexicographic (xi, eta, zeta)
                # (f0->f3 and f0->f4 swap roles, src/TetgenCells.cpp:673-686): the integrand only sees |det J|
                assert np.all(det < 0) or np.all(det > 0)
                det = np.abs(det)
                G = np.einsum("ke,ned->nkd", dN, np.linalg.inv(J))          # grad_x N_k = J^-T grad_xi N_k
                Ke += (wa * wb * wc) * det[:, None, None] * np.einsum("nid,njd->nij", G, G)
                vol += (wa * wb * wc) * det
    return Ke, vol


def _boundary_faces(hexs):
    """faces of the cell complex that belong to exactly one hexahedron"""
    fv = np.stack([hexs[:, f] for f in FACES], 1).reshape(-1, 4)          # (6n, 4)
    key = np.sort(fv, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    return fv[cnt[inv.reshape(-1)] == 1]


def _solve(m, field, anode_potential=None):
    nodes = m["nodes"]; hexs = m["hexs"][m["hex_markers"] > 0].astype(np.int64)
    n = len(nodes)
    Ke, vol = _stiffness(nodes, hexs)
    assert vol.min() > 0
    rows = np.repeat(hexs, 8, axis=1).reshape(-1); cols = np.tile(hexs, (1, 8)).reshape(-1)
    K = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    bf = _boundary_faces(hexs)
    ctr = nodes[bf].mean(1)
    mn, mx = ctr.min(0), ctr.max(0)
    eps = 1e-6
    side = (np.abs(ctr[:, 0] - mn[0]) <= eps) | (np.abs(ctr[:, 0] - mx[0]) <= eps) | (np.abs(ctr[:, 1] - mn[1]) <= eps) | (np.abs(ctr[:, 1] - mx[1]) <= eps)
    top = ~side & (np.abs(ctr[:, 2] - mx[2]) <= eps)
    copper = ~side & ~top
    # Neumann load on the top faces: b_i += sum_q N_i(q) * field * |dx/ds x dx/dt| on [-1,1]^2
    b = np.zeros(n)
    gp, gw = np.polynomial.legendre.leggauss(2)
    P = nodes[bf[top]]                                        # corners in cyclic order
    s4 = np.array([-1, 1, 1, -1.0]); t4 = np.array([-1, -1, 1, 1.0])
    for s, ws in zip(gp, gw):
        for t, wt in zip(gp, gw):
            N = (1 + s4 * s) * (1 + t4 * t) / 4.0
            ds = np.einsum("k,nkd->nd", s4 * (1 + t4 * t) / 4.0, P); dt = np.einsum("k,nkd->nd", (1 + s4 * s) * t4 / 4.0, P)
            dA = np.linalg.norm(np.cross(ds, dt), axis=1)
            np.add.at(b, bf[top].reshape(-1), (ws * wt * field * dA[:, None] * N[None, :]).reshape(-1))
    used = np.zeros(n, bool); used[hexs.reshape(-1)] = True
    fixed = np.zeros(n, bool); fixed[bf[copper].reshape(-1)] = True
    phi = np.zeros(n)
    if anode_potential is not None:          # Dirichlet anode (append_dirichlet on vacuum_top, later call wins on shared dofs)
        b[:] = 0.0
        fixed[bf[top].reshape(-1)] = True
        phi[bf[top].reshape(-1)] = anode_potential
    free = used & ~fixed
    phi[free] = spla.spsolve(K[free][:, free].tocsc(), b[free] - (K[free][:, fixed] @ phi[fixed]))
    return K, b, phi, used, vol


@pytest.mark.parametrize("name", ["hemicone", "mdsmall"])
def test_oracle_arithmetic_matches_an_independent_derivation(name, golden):
    m = golden("mesh", name)
    field = 0.5
    K, b, phi, used, vol = _solve(m, field)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    o.setup(field, 0.0, False); o.assemble(True)
    assert o.solve(10000, 1e-12, 1.2, 0) > 0
    rhs, sol, v2d, v2n = o.vectors()
    assert np.array_equal(v2n, np.flatnonzero(used))                       # vertex = rank among the nodes of vacuum hexes
    # stiffness matrix before boundary conditions, brought to node numbering
    rp, col, val, save = o.csr()
    node_of_dof = np.zeros(o.n_dofs, np.int64); node_of_dof[v2d] = v2n
    Ko = sp.csr_matrix((save, col, rp))
    Ko = sp.coo_matrix((Ko.data, (node_of_dof[Ko.tocoo().row], node_of_dof[Ko.tocoo().col])), shape=K.shape).tocsr()
    assert abs(Ko - K).max() <= 1e-12 * abs(K).max()
    assert abs(sum(o.cell_vol(k) for k in range(0, o.n_cells, 97)) - vol[::97].sum()) <= 1e-10 * vol[::97].sum()
    # potential: direct solve of the reduced system vs the oracle's eliminate-and-CG
    ref = o.export_solution()
    assert np.abs(ref - phi[v2n]).max() <= 1e-9 * np.abs(ref).max()
    # Dirichlet anode (anode_BC = dirichlet): V0 on the top plane, 0 on copper, no Neumann load
    _, _, phi_d, _, _ = _solve(m, 0.0, anode_potential=7.5)
    o.setup(0.0, 7.5, True); o.assemble(True)
    assert o.solve(10000, 1e-12, 1.2, 0) > 0
    ref = o.export_solution()
    assert np.abs(ref - phi_d[v2n]).max() <= 1e-9 * np.abs(ref).max()


def test_space_charge_rhs_from_reference_pinned_weights(golden):
    """PoissonSolver::assemble_space_charge_fast (PoissonSolver.cpp:299-319) re-derived: the right-hand side gained by
    adding particles = scatter of (shape function x charge factor) over the 8 vertices of each particle's cell, with the
    shape functions taken from the oracle entry point that tests/test_oracle_vs_ref.py pins to the compiled reference;
    constrained (Dirichlet) rows keep their boundary value"""
    m = golden("mesh", "mdsmall"); g = golden("interp", "mdsmall")
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    ok = g["pic_ok"]
    pts = g["points"][ok]; cells = g["pic_cells"][ok]
    cf = -180.9512268 * 0.01
    o.setup(0.5, 0.0, False); o.assemble(True)
    rhs0 = o.vectors()[0].copy()
    o.setup(0.5, 0.0, False); o.assemble(True, pts, cells, cf)
    rhs1, _, v2d, _ = o.vectors()
    w = o.particle_weights(pts, cells)              # shape_funs_dealii (InterpolatorCells.cpp:1355-1358): deal.II vertex order
    dofs = v2d[o.cells()[cells]]                    # (n, 8) DoFs of each particle's cell, same (lexicographic) order
    expect = np.zeros(o.n_dofs)
    np.add.at(expect, dofs.reshape(-1), (w * cf).reshape(-1))
    rp, col, val, save = o.csr()
    constrained = np.array([rp[r + 1] - rp[r] >= 1 and np.count_nonzero(val[rp[r]:rp[r + 1]]) == 1 for r in range(o.n_dofs)])
    free = ~constrained
    assert free.sum() > 0.5 * o.n_dofs
    scale = np.abs(expect).max()
    assert np.abs((rhs1 - rhs0)[free] - expect[free]).max() <= 1e-12 * scale
    assert np.array_equal(rhs1[constrained], rhs0[constrained])
