"""Synthetic hexahedral meshes for tests and benchmarks (host side, numpy).

FEMOCS hexahedra are stored in the "old-style / UCD" vertex order that
``Hexahedra::export_vacuum`` hands to deal.II (reference src/TetgenCells.cpp:673-686,
orientation fixed in src/Tethex.cpp:1565-1592): the solver's lexicographic cell is
``[f0, f1, f4, f5, f3, f2, f7, f6]``.  The helpers here produce meshes in that same
convention so that they can be fed to ``PoissonSolver.import_mesh`` unchanged:

* ``box_mesh``      -- a structured box, used for known-answer tests (uniform field);
* ``refine_hexes``  -- uniform 1 -> 8 refinement by trilinear midpoint insertion, used to
  build the HBM-sized "config 4" style workloads from a native tet-split mesh
  ("uniformly refined ... each hex -> 8 by trilinear midpoint insertion").
"""
import numpy as np

# lexicographic (x fastest) position d = a + 2b + 4c  ->  FEMOCS/UCD local vertex index
DEAL2UCD = np.array([0, 1, 4, 5, 3, 2, 7, 6], dtype=np.int64)


def box_mesh(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, origin=(0.0, 0.0, 0.0), jitter=0.0, seed=0):
    """Structured box of nx*ny*nz hexahedra; returns (nodes[n,3], hexs[n,8] int32, markers[n] int32)."""
    xs = np.linspace(0, lx, nx + 1); ys = np.linspace(0, ly, ny + 1); zs = np.linspace(0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1) + np.asarray(origin)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        interior = ((X > 0) & (X < lx) & (Y > 0) & (Y < ly) & (Z > 0) & (Z < lz)).ravel()
        h = min(lx / nx, ly / ny, lz / nz)
        nodes[interior] += rng.uniform(-jitter * h, jitter * h, size=(interior.sum(), 3))

    def nid(i, j, k):
        return (k * (ny + 1) + j) * (nx + 1) + i

    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I = I.ravel(); J = J.ravel(); K = K.ravel()
    lex = np.stack([nid(I + (d & 1), J + ((d >> 1) & 1), K + ((d >> 2) & 1)) for d in range(8)], axis=1)
    hexs = np.empty_like(lex)
    hexs[:, DEAL2UCD] = lex
    return nodes, hexs.astype(np.int32), np.ones(len(hexs), np.int32)


def _unique_rows_sorted(keys):
    """keys: (n, k) int64 rows already sorted within the row; returns (inverse, n_unique, first_index)."""
    order = np.lexsort(tuple(keys[:, c] for c in range(keys.shape[1] - 1, -1, -1)))
    sk = keys[order]
    new = np.ones(len(sk), bool)
    new[1:] = np.any(sk[1:] != sk[:-1], axis=1)
    uid = np.cumsum(new) - 1
    inverse = np.empty(len(sk), np.int64)
    inverse[order] = uid
    first = order[new]
    return inverse, int(uid[-1]) + 1 if len(uid) else 0, first


def refine_hexes(nodes, hexs):
    """Split every hexahedron into 8 (edge mid-points, face centres, cell centre).

    Shared edges/faces get one new node each, so the refined mesh is conforming.  Input and
    output hexes are in FEMOCS/UCD order.  Returns (nodes, hexs int32).
    """
    nodes = np.asarray(nodes, np.float64)
    lex = np.asarray(hexs, np.int64)[:, DEAL2UCD]          # (n, 8) lexicographic
    n_hex = len(lex); n0 = len(nodes)

    # 12 edges, 6 faces of the lexicographic cube: local (a,b,c) lattice coordinates in {0,1,2}
    def corner(d):
        return (d & 1) * 2, ((d >> 1) & 1) * 2, ((d >> 2) & 1) * 2

    edges = [(d0, d1) for d0 in range(8) for d1 in range(d0 + 1, 8) if bin(d0 ^ d1).count("1") == 1]
    faces = [[d for d in range(8) if ((d >> ax) & 1) == side] for ax in range(3) for side in range(2)]

    lattice = np.full((n_hex, 3, 3, 3), -1, np.int64)
    for d in range(8):
        a, b, c = corner(d)
        lattice[:, a, b, c] = lex[:, d]

    # edge midpoints
    ek = np.stack([np.stack([lex[:, d0], lex[:, d1]], axis=1) for d0, d1 in edges], axis=1)   # (n,12,2)
    ek = np.sort(ek.reshape(-1, 2), axis=1)
    key = ek[:, 0] * np.int64(n0) + ek[:, 1]
    uk, first, inv = np.unique(key, return_index=True, return_inverse=True)
    e_nodes = 0.5 * (nodes[ek[first, 0]] + nodes[ek[first, 1]])
    e_ids = (n0 + inv).reshape(n_hex, 12)
    for e, (d0, d1) in enumerate(edges):
        a0, b0, c0 = corner(d0); a1, b1, c1 = corner(d1)
        lattice[:, (a0 + a1) // 2, (b0 + b1) // 2, (c0 + c1) // 2] = e_ids[:, e]
    n1 = n0 + len(uk)

    # face centres
    fk = np.stack([lex[:, f] for f in faces], axis=1).reshape(-1, 4)                         # (n*6,4)
    fks = np.sort(fk, axis=1)
    inv, nf, first = _unique_rows_sorted(fks[:, :3])
    f_nodes = 0.25 * (nodes[fk[first, 0]] + nodes[fk[first, 1]] + nodes[fk[first, 2]] + nodes[fk[first, 3]])
    f_ids = (n1 + inv).reshape(n_hex, 6)
    for f, fv in enumerate(faces):
        cs = np.array([corner(d) for d in fv]).mean(axis=0).astype(int)
        lattice[:, cs[0], cs[1], cs[2]] = f_ids[:, f]
    n2 = n1 + nf

    # cell centres
    c_nodes = nodes[lex].mean(axis=1)
    lattice[:, 1, 1, 1] = n2 + np.arange(n_hex)

    new_nodes = np.vstack([nodes, e_nodes, f_nodes, c_nodes])
    out = np.empty((n_hex, 8, 8), np.int64)
    for s in range(8):
        sa, sb, sc = s & 1, (s >> 1) & 1, (s >> 2) & 1
        for d in range(8):
            out[:, s, DEAL2UCD[d]] = lattice[:, sa + (d & 1), sb + ((d >> 1) & 1), sc + ((d >> 2) & 1)]
    return new_nodes, out.reshape(-1, 8).astype(np.int32)


def refine_vacuum(nodes, hexs, hex_markers, levels):
    """Keep the vacuum hexes (marker > 0), compact the node list and refine ``levels`` times."""
    hexs = np.asarray(hexs)[np.asarray(hex_markers) > 0]
    used = np.unique(hexs)
    remap = np.full(len(nodes), -1, np.int64); remap[used] = np.arange(len(used))
    nodes = np.asarray(nodes)[used]; hexs = remap[hexs].astype(np.int32)
    for _ in range(levels):
        nodes, hexs = refine_hexes(nodes, hexs)
    return nodes, hexs, np.ones(len(hexs), np.int32)
