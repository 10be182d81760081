// Host-side mesh partitioning for the multi-GPU CG (the reference has no distributed
// path -- DealSolver uses the serial Triangulation / SparseMatrix, include/DealSolver.h:135-143 -- so this is
// new functionality behind the same PoissonSolver interface).
//
// One process per GPU; EVERY rank is handed the same full mesh (the host mesher is serial, as in the
// reference) and deterministically derives its own share:
//   * solver vertices are split by recursive coordinate bisection into `world` equal parts (owner rank);
//   * a rank keeps every vacuum hexahedron that touches one of its vertices, so the matrix rows of its
//     owned vertices are complete without any exchange of element contributions;
//   * local numbering: owned vertices first (first touch in cell order), then the ghosts grouped by owner
//     rank and sorted by global id -- the ghost segment of peer p is contiguous and both sides of a halo
//     exchange derive the same ordering without talking to each other;
//   * halo plan: send[p] = owned vertices that share a cell with a vertex of p (= p's ghosts owned here).
// No data-path collective is needed to build any of this; the only setup exchanges are the extremes of
// the boundary-face centres (mark_boundary, src/DealSolver.cpp:460-518) and the Dirichlet flags of ghosts.
#include <omp.h>

#include <algorithm>
#include <numeric>

#include "ctx.h"

namespace {

// recursive coordinate bisection of idx[lo, hi) into parts [p0, p0 + np): split along the longest axis so
// that the two halves hold shares proportional to their number of parts; ties broken by vertex id
void rcb(const double* vx, std::vector<int>& idx, int lo, int hi, int p0, int np, std::vector<int>& owner) {
    if (np == 1) { for (int i = lo; i < hi; ++i) owner[idx[i]] = p0; return; }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = lo; i < hi; ++i)
        for (int d = 0; d < 3; ++d) { const double x = vx[3 * (size_t) idx[i] + d]; mn[d] = std::min(mn[d], x); mx[d] = std::max(mx[d], x); }
    int ax = 0;
    for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
    const int npl = np / 2;
    const int mid = lo + (int) ((long) (hi - lo) * npl / np);
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
        const double xa = vx[3 * (size_t) a + ax], xb = vx[3 * (size_t) b + ax];
        return xa < xb || (xa == xb && a < b);
    });
    rcb(vx, idx, lo, mid, p0, npl, owner);
    rcb(vx, idx, mid, hi, p0 + npl, np - npl, owner);
}

}  // namespace

int fb_host_partition_phase1(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    const int rank = c->rank, world = c->world;
    for (size_t i = 0; i < 8 * (size_t) n_hex; ++i)
        if (hex8[i] < 0 || hex8[i] >= n_nodes) return c->fail(FB_ERR_MESH, "hexahedron %zu references node %d", i / 8, hex8[i]);
    // ---- global solver vertices (order-preserving compaction, as on one GPU) and cells ----
    std::vector<int> node2vert(n_nodes, -1), cells_g;
    for (int h = 0; h < n_hex; ++h)
        if (hex_marker[h] > 0) { cells_g.push_back(h); for (int k = 0; k < 8; ++k) node2vert[hex8[8 * (size_t) h + k]] = 0; }
    if (cells_g.empty()) return c->fail(FB_ERR_MESH, "no vacuum hexahedra (marker > 0) in the mesh");
    std::vector<int> vert2node;
    for (int i = 0; i < n_nodes; ++i) if (node2vert[i] == 0) { node2vert[i] = (int) vert2node.size(); vert2node.push_back(i); }
    const int nv = (int) vert2node.size(), ncg = (int) cells_g.size();
    c->n_vert_global = nv; c->n_cells_global = ncg; c->n_dofs_global = nv;
    // ---- owner of every vertex ----
    std::vector<double> vx(3 * (size_t) nv);
    for (int v = 0; v < nv; ++v) for (int d = 0; d < 3; ++d) vx[3 * (size_t) v + d] = xyz[3 * (size_t) vert2node[v] + d];
    std::vector<int> idx(nv), owner(nv, 0);
    std::iota(idx.begin(), idx.end(), 0);
    rcb(vx.data(), idx, 0, nv, 0, world, owner);
    // ---- local cells: every cell touching an owned vertex ----
    c->part_cell_g.clear();
    for (int ce = 0; ce < ncg; ++ce) {
        const int* h = &hex8[8 * (size_t) cells_g[ce]];
        bool mine = false;
        for (int k = 0; k < 8 && !mine; ++k) mine = owner[node2vert[h[k]]] == rank;
        if (mine) c->part_cell_g.push_back(ce);
    }
    const int ncl = (int) c->part_cell_g.size();
    if (ncl == 0) return c->fail(FB_ERR_MESH, "rank %d of %d owns no vertices", rank, world);
    // ---- local vertices: owned in first-touch order, ghosts by (owner, global id) ----
    std::vector<int> g2l(nv, -1), owned, ghosts;
    for (int lc = 0; lc < ncl; ++lc) {
        const int* h = &hex8[8 * (size_t) cells_g[c->part_cell_g[lc]]];
        static const int LEX2UCD[8] = {0, 1, 4, 5, 3, 2, 7, 6};          // lexicographic position d holds UCD vertex LEX2UCD[d]
        for (int d = 0; d < 8; ++d) {
            const int v = node2vert[h[LEX2UCD[d]]];
            if (g2l[v] != -1) continue;               // already numbered (owned) or already listed (ghost, -2)
            if (owner[v] == rank) { g2l[v] = (int) owned.size(); owned.push_back(v); }
            else { g2l[v] = -2; ghosts.push_back(v); }
        }
    }
    std::sort(ghosts.begin(), ghosts.end(), [&](int a, int b) { return owner[a] < owner[b] || (owner[a] == owner[b] && a < b); });
    const int n_own = (int) owned.size(), n_loc = n_own + (int) ghosts.size();
    for (size_t i = 0; i < ghosts.size(); ++i) g2l[ghosts[i]] = n_own + (int) i;
    c->part_l2g = owned; c->part_l2g.insert(c->part_l2g.end(), ghosts.begin(), ghosts.end());
    c->part_owner.resize(n_loc);
    for (int l = 0; l < n_loc; ++l) c->part_owner[l] = owner[c->part_l2g[l]];
    // ghost segments per peer
    c->recv_off.assign(world + 1, 0);
    for (int v : ghosts) c->recv_off[owner[v] + 1]++;
    for (int p = 0; p < world; ++p) c->recv_off[p + 1] += c->recv_off[p];
    // ---- halo send lists: owned vertices sharing a cell with a vertex of p, sorted by global id ----
    std::vector<std::vector<int>> send(world);
    for (int lc = 0; lc < ncl; ++lc) {
        const int* h = &hex8[8 * (size_t) cells_g[c->part_cell_g[lc]]];
        int own8[8]; bool any_other = false;
        for (int k = 0; k < 8; ++k) { own8[k] = owner[node2vert[h[k]]]; any_other |= own8[k] != rank; }
        if (!any_other) continue;
        for (int k = 0; k < 8; ++k) {
            if (own8[k] != rank) continue;
            for (int q = 0; q < 8; ++q) if (own8[q] != rank) send[own8[q]].push_back(node2vert[h[k]]);
        }
    }
    c->send_off.assign(world + 1, 0); c->send_idx.clear();
    for (int p = 0; p < world; ++p) {
        std::sort(send[p].begin(), send[p].end());
        send[p].erase(std::unique(send[p].begin(), send[p].end()), send[p].end());
        for (int v : send[p]) c->send_idx.push_back(g2l[v]);
        c->send_off[p + 1] = (int) c->send_idx.size();
    }
    // ---- local sub-mesh in local vertex order, then the ordinary import on it ----
    std::vector<double> xyz_l(3 * (size_t) n_loc);
    for (int l = 0; l < n_loc; ++l) for (int d = 0; d < 3; ++d) xyz_l[3 * (size_t) l + d] = vx[3 * (size_t) c->part_l2g[l] + d];
    std::vector<int> hex_l(8 * (size_t) ncl), mark_l(ncl, 1);
#pragma omp parallel for schedule(static)
    for (int lc = 0; lc < ncl; ++lc) {
        const int* h = &hex8[8 * (size_t) cells_g[c->part_cell_g[lc]]];
        for (int k = 0; k < 8; ++k) hex_l[8 * (size_t) lc + k] = g2l[node2vert[h[k]]];
    }
    c->part_n_owned = n_own;
    c->part_gxyz.swap(vx);              // global vertex coordinates: the two-level preconditioner aggregates along the GLOBAL Morton curve
    return fb_host_import_phase1(c, xyz_l.data(), n_loc, hex_l.data(), mark_l.data(), ncl);
}
