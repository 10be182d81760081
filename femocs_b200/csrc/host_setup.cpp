// Host-side setup of libfemocs_b200: everything the north star keeps on the host
// ("Mesh generation and DoF numbering stay on the host in C++"): vertex compaction, cell
// orientation, boundary marking, DoF numbering, CSR sparsity, and the interpolator's
// per-cell precompute tables.  OpenMP-parallel, no CUDA calls here.
//
// Reference behaviour followed (paths relative to the reference root):
//   src/TetgenCells.cpp:673-686   Hexahedra::export_vacuum
//   src/DealSolver.cpp:191-209    import_mesh (delete_unused_vertices, invert_all_cells_of_negative_grid,
//                                 create_triangulation_compatibility)
//   src/DealSolver.cpp:460-518    mark_boundary, src/PoissonSolver.cpp:52-55 mark_mesh
//   src/DealSolver.cpp:368-387    setup_system (distribute_dofs, make_sparsity_pattern)
//   src/InterpolatorCells.cpp:523-629,1205-1267,1585-1637,1151-1173,1873-1895  precompute()
//   src/Interpolator.cpp:60-76    node2cells
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>

#include "ctx.h"

namespace {

// lexicographic (deal.II) local vertex d  <-  femocs/UCD local vertex: lex[ucd_to_lex[k]] = ucd[k]
const int UCD_TO_LEX[8] = {0, 1, 5, 4, 2, 3, 7, 6};
// vertices of the six faces of the lexicographic cube (deal.II GeometryInfo<3>)
const int FACE_VERTS[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};

struct P3 { double x, y, z; };

// signed volume of a trilinear hexahedron by 2x2x2 Gauss quadrature of det(J) (exact)
double hex_volume(const P3 v[8]) {
    const double a = 0.5 * (1.0 - 1.0 / std::sqrt(3.0)), b = 0.5 * (1.0 + 1.0 / std::sqrt(3.0));
    const double g[2] = {a, b};
    double vol = 0;
    for (int q = 0; q < 8; ++q) {
        const double xi[3] = {g[q & 1], g[(q >> 1) & 1], g[(q >> 2) & 1]};
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < 8; ++i) {
            double f[3], df[3];
            for (int d = 0; d < 3; ++d) {
                const int bit = (i >> d) & 1;
                f[d] = bit ? xi[d] : 1.0 - xi[d];
                df[d] = bit ? 1.0 : -1.0;
            }
            const double dn[3] = {df[0] * f[1] * f[2], f[0] * df[1] * f[2], f[0] * f[1] * df[2]};
            for (int e = 0; e < 3; ++e) {
                J[0][e] += v[i].x * dn[e]; J[1][e] += v[i].y * dn[e]; J[2][e] += v[i].z * dn[e];
            }
        }
        vol += 0.125 * (J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                      - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                      + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]));
    }
    return vol;
}

inline uint64_t spread21(uint64_t v) {          // interleave helper for 63-bit Morton keys
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

// FB_VERBOSE=1: wall-clock laps of the host set-up on stderr
struct Laps {
    const char* what; double t0, t; bool on; std::string out;
    explicit Laps(const char* w) : what(w), t0(omp_get_wtime()), t(t0), on(getenv("FB_VERBOSE") != nullptr) {}
    void lap(const char* name) { if (!on) return; const double n = omp_get_wtime(); char b[64]; snprintf(b, sizeof b, " %s %.2f", name, 1e3 * (n - t)); out += b; t = n; }
    ~Laps() { if (on) fprintf(stderr, "[fb] %s (ms):%s | total %.2f\n", what, out.c_str(), 1e3 * (omp_get_wtime() - t0)); }
};

}  // namespace

// Phase 1: vertex compaction, orientation, boundary faces and the extremes of their centres (c->bb_mn/bb_mx).
// On one GPU phase 2 follows directly; with a partitioned mesh the extremes are first reduced over the ranks
// (mark_boundary of the reference works on the global mesh).  c->part_n_owned >= 0 marks a partition: the
// local vertices [0, part_n_owned) are owned, the rest are ghosts, and only faces touching an owned vertex count.
int fb_host_import_phase1(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    Laps laps("import phase 1");
    c->mesh_ok = false;
    c->n_nodes = n_nodes; c->n_hex = n_hex;
    // (resize + parallel copy: vector::assign is one serial memcpy, 1.3 GB on the 2.3e7-DoF mesh)
    c->xyz.resize(3 * (size_t) n_nodes); c->hex8.resize(8 * (size_t) n_hex); c->hex_marker.resize(n_hex);
    long bad_entry = -1;
#pragma omp parallel
    {
#pragma omp for schedule(static) nowait
        for (long i = 0; i < 3L * n_nodes; ++i) c->xyz[i] = xyz[i];
#pragma omp for schedule(static) nowait
        for (long h = 0; h < n_hex; ++h) c->hex_marker[h] = hex_marker[h];
#pragma omp for schedule(static)
        for (long i = 0; i < 8L * n_hex; ++i) {
            const int v = hex8[i];
            c->hex8[i] = v;
            if (v < 0 || v >= n_nodes) {
#pragma omp critical
                if (bad_entry < 0 || i < bad_entry) bad_entry = i;
            }
        }
    }
    if (bad_entry >= 0) return c->fail(FB_ERR_MESH, "hexahedron %ld references node %d", bad_entry / 8, hex8[bad_entry]);

    laps.lap("copy+check");
    // ---- solver cells and order-preserving vertex compaction ----
    c->hex2cell.assign(n_hex, -1); c->cell2hex.clear();
    c->node2vert.assign(n_nodes, -1);
    // mesh_kind 0: vacuum hexahedra (Hexahedra::export_vacuum, marker > 0); 1: bulk (export_bulk, marker < 0; TetgenCells.cpp:688-701)
    for (int h = 0; h < n_hex; ++h)
        if (c->mesh_kind ? hex_marker[h] < 0 : hex_marker[h] > 0) {
            c->hex2cell[h] = (int) c->cell2hex.size();
            c->cell2hex.push_back(h);
        }
    {   // nodes touched by a solver cell (every writer stores the same 0: benign)
        const long ncell = (long) c->cell2hex.size();
#pragma omp parallel for schedule(static)
        for (long ce = 0; ce < ncell; ++ce) {
            const int* h8 = &hex8[8 * (size_t) c->cell2hex[ce]];
            for (int k = 0; k < 8; ++k) if (c->node2vert[h8[k]] != 0) c->node2vert[h8[k]] = 0;
        }
    }
    const int n_cells = c->n_cells = (int) c->cell2hex.size();
    if (n_cells == 0) return c->fail(FB_ERR_MESH, c->mesh_kind ? "no bulk hexahedra (marker < 0) in the mesh" : "no vacuum hexahedra (marker > 0) in the mesh");
    c->vert2node.clear();
    for (int i = 0; i < n_nodes; ++i)
        if (c->node2vert[i] == 0) { c->node2vert[i] = (int) c->vert2node.size(); c->vert2node.push_back(i); }
    const int n_vert = c->n_vert = (int) c->vert2node.size();

    std::vector<int>& cv = c->h_cv;                  // lexicographic vertex ids
    cv.assign(8 * (size_t) n_cells, 0);
#pragma omp parallel for schedule(static)
    for (int ce = 0; ce < n_cells; ++ce) {
        const int* h = &hex8[8 * (size_t) c->cell2hex[ce]];
        for (int k = 0; k < 8; ++k) cv[8 * (size_t) ce + UCD_TO_LEX[k]] = c->node2vert[h[k]];
    }

    laps.lap("compaction");
    // ---- orientation (invert_all_cells_of_negative_grid) ----
    long n_neg = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_neg)
    for (int ce = 0; ce < n_cells; ++ce) {
        P3 v[8];
        for (int k = 0; k < 8; ++k) {
            const double* p = &xyz[3 * (size_t) c->vert2node[cv[8 * (size_t) ce + k]]];
            v[k] = {p[0], p[1], p[2]};
        }
        if (hex_volume(v) < 0) ++n_neg;
    }
    c->mesh_flipped = (n_neg == n_cells);
    c->imported_kind = c->mesh_kind;
    if (n_neg == n_cells) {
        // deal.II 9.2 swaps vertices i <-> i + 4 of the OLD-STYLE (UCD) numbering the cells arrive in; through
        // UCD_TO_LEX these are the lexicographic pairs (0,2) (1,3) (4,6) (5,7)
#pragma omp parallel for schedule(static)
        for (int ce = 0; ce < n_cells; ++ce)
            for (int k : {0, 1, 4, 5}) std::swap(cv[8 * (size_t) ce + k], cv[8 * (size_t) ce + k + 2]);
    } else if (n_neg > 0) {
        return c->fail(FB_ERR_MESH, "%ld of %d hexahedra have negative volume", n_neg, n_cells);
    }

    laps.lap("orientation");
    // ---- vertex -> cells adjacency (CSR) ----
    std::vector<int>& v2c_off = c->h_v2c_off; std::vector<int>& v2c = c->h_v2c;
    // counting sort with atomic counters (all threads), then every vertex's short list is put in ascending cell order
    v2c_off.assign(n_vert + 1, 0);
    const long n_cv = (long) cv.size();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_cv; ++i) __atomic_fetch_add(&v2c_off[cv[i] + 1], 1, __ATOMIC_RELAXED);
    for (int v = 0; v < n_vert; ++v) v2c_off[v + 1] += v2c_off[v];
    v2c.resize(cv.size());
    std::vector<int> fill(v2c_off.begin(), v2c_off.end() - 1);
#pragma omp parallel for schedule(static)
    for (int ce = 0; ce < n_cells; ++ce)
        for (int k = 0; k < 8; ++k) v2c[__atomic_fetch_add(&fill[cv[8 * (size_t) ce + k]], 1, __ATOMIC_RELAXED)] = ce;
#pragma omp parallel for schedule(static, 4096)
    for (int v = 0; v < n_vert; ++v) std::sort(v2c.begin() + v2c_off[v], v2c.begin() + v2c_off[v + 1]);     // ascending cell ids per vertex

    laps.lap("v2c");
    // ---- boundary faces: a face is interior iff another cell holds all 4 of its vertices ----
    std::vector<unsigned char>& is_b = c->h_isb;
    is_b.assign(6 * (size_t) n_cells, 0);
    const int n_owned_v = c->part_n_owned >= 0 ? c->part_n_owned : n_vert;
#pragma omp parallel for schedule(dynamic, 512)
    for (int ce = 0; ce < n_cells; ++ce)
        for (int f = 0; f < 6; ++f) {
            int fv[4];
            for (int k = 0; k < 4; ++k) fv[k] = cv[8 * (size_t) ce + FACE_VERTS[f][k]];
            int best = 0;
            for (int k = 1; k < 4; ++k)
                if (v2c_off[fv[k] + 1] - v2c_off[fv[k]] < v2c_off[fv[best] + 1] - v2c_off[fv[best]]) best = k;
            bool interior = false;
            for (int q = v2c_off[fv[best]]; q < v2c_off[fv[best] + 1] && !interior; ++q) {
                const int other = v2c[q];
                if (other == ce) continue;
                int hit = 0;
                for (int k = 0; k < 8; ++k) {
                    const int ov = cv[8 * (size_t) other + k];
                    hit += (ov == fv[0]) + (ov == fv[1]) + (ov == fv[2]) + (ov == fv[3]);
                }
                interior = (hit == 4);
            }
            // partition: a face without an owned vertex may border a cell this rank does not hold -> not ours to judge
            const bool ours = fv[0] < n_owned_v || fv[1] < n_owned_v || fv[2] < n_owned_v || fv[3] < n_owned_v;
            is_b[6 * (size_t) ce + f] = !interior && ours;
        }
    auto face_centre = [&](int ce, int f, double ctr[3]) {      // TriaAccessor::center: vertex mean
        double s[3] = {0, 0, 0};
        for (int k = 0; k < 4; ++k) {
            const double* p = &xyz[3 * (size_t) c->vert2node[cv[8 * (size_t) ce + FACE_VERTS[f][k]]]];
            s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
        }
        ctr[0] = s[0] / 4.0; ctr[1] = s[1] / 4.0; ctr[2] = s[2] / 4.0;
    };
    laps.lap("faces");
    c->bfaces.clear();
    double mx[3] = {-1e16, -1e16, -1e16}, mn[3] = {1e16, 1e16, 1e16};
    for (int ce = 0; ce < n_cells; ++ce)
        for (int f = 0; f < 6; ++f)
            if (is_b[6 * (size_t) ce + f]) {
                c->bfaces.push_back({ce, f, 0});
                double p[3]; face_centre(ce, f, p);
                for (int d = 0; d < 3; ++d) { mx[d] = std::max(mx[d], p[d]); mn[d] = std::min(mn[d], p[d]); }
            }
    for (int d = 0; d < 3; ++d) { c->bb_mn[d] = mn[d]; c->bb_mx[d] = mx[d]; }
    return FB_OK;
}

// Phase 2: boundary ids from the (global) extremes, DoF numbering, sparsity of the owned rows, Dirichlet candidates.
int fb_host_import_phase2(fb_ctx* c) {
    const double* xyz = c->xyz.data();
    const int n_vert = c->n_vert;
    std::vector<int>& cv = c->h_cv; std::vector<int>& v2c_off = c->h_v2c_off; std::vector<int>& v2c = c->h_v2c;
    const double* mn = c->bb_mn; const double* mx = c->bb_mx;
    Laps laps("import phase 2");
    auto face_centre = [&](int ce, int f, double ctr[3]) {      // TriaAccessor::center: vertex mean
        double s[3] = {0, 0, 0};
        for (int k = 0; k < 4; ++k) {
            const double* p = &xyz[3 * (size_t) c->vert2node[cv[8 * (size_t) ce + FACE_VERTS[f][k]]]];
            s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
        }
        ctr[0] = s[0] / 4.0; ctr[1] = s[1] / 4.0; ctr[2] = s[2] / 4.0;
    };
    const double eps = 1e-6;
    c->n_top_faces = 0;
    for (auto& bf : c->bfaces) {
        double p[3]; face_centre(bf.cell, bf.face, p);
        auto on = [&](double v, double b) { return std::fabs(v - b) <= eps; };
        if (c->mesh_kind == 0) {
            // PoissonSolver::mark_mesh (PoissonSolver.cpp:52-55): top = vacuum_top, bottom = other = copper_surface
            if (on(p[0], mn[0]) || on(p[0], mx[0]) || on(p[1], mn[1]) || on(p[1], mx[1])) bf.id = 4;   // vacuum_sides
            else if (on(p[2], mx[2])) { bf.id = 8; c->n_top_faces++; }                               // vacuum_top
            else bf.id = 2;                                                                           // copper_surface (bottom & other)
        } else {
            // CurrentHeatSolver::mark_mesh (CurrentHeatSolver.cpp:526-530): top = other = copper_surface (Neumann faces with
            // per-face emission data), bottom = copper_bottom (Dirichlet), sides = copper_sides
            if (on(p[0], mn[0]) || on(p[0], mx[0]) || on(p[1], mn[1]) || on(p[1], mx[1])) bf.id = 4;
            else if (on(p[2], mx[2])) { bf.id = 2; c->n_top_faces++; }
            else if (on(p[2], mn[2])) bf.id = 7;                                                      // copper_bottom
            else { bf.id = 2; c->n_top_faces++; }
        }
    }

    laps.lap("face ids");
    c->imported_degree = c->fe_degree;
    if (c->fe_degree == 2) return fb_host_q2_phase2(c);          // FE_Q(2): numbering, sparsity, boundary sets in q2.cu
    // ---- DoF numbering ----
    c->vertex2dof.assign(n_vert, -1);
    int n_dofs = 0;
    if (c->part_n_owned >= 0) {
        // partition: the local vertex order (owned first, ghosts grouped by owner) IS the numbering
        for (int v = 0; v < n_vert; ++v) c->vertex2dof[v] = v;
        n_dofs = n_vert;
    } else if (c->dof_order == 1) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int v = 0; v < n_vert; ++v)
            for (int d = 0; d < 3; ++d) {
                const double x = xyz[3 * (size_t) c->vert2node[v] + d];
                lo[d] = std::min(lo[d], x); hi[d] = std::max(hi[d], x);
            }
        std::vector<std::pair<uint64_t, int>> keys(n_vert);
#pragma omp parallel for schedule(static)
        for (int v = 0; v < n_vert; ++v) {
            uint64_t q[3];
            for (int d = 0; d < 3; ++d) {
                const double t = (xyz[3 * (size_t) c->vert2node[v] + d] - lo[d]) / std::max(hi[d] - lo[d], 1e-300);
                q[d] = (uint64_t) std::min(2097151.0, std::max(0.0, t * 2097151.0));
            }
            keys[v] = {spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2), v};
        }
        std::sort(keys.begin(), keys.end());
        for (int i = 0; i < n_vert; ++i) c->vertex2dof[keys[i].second] = i;
        n_dofs = n_vert;
    } else {
        // FE_Q(1) first-touch numbering: cells in order, local vertices 0..7 (deal.II distribute_dofs).  In parallel: the
        // dof of a vertex is the number of vertices whose FIRST occurrence in the cell list precedes its own, i.e. an
        // exclusive scan over the "first occurrence" flags of the list.
        const long n_cv = (long) cv.size();
        if (n_cv < 2000000 || n_cv > 2147483000L) {
            for (size_t i = 0; i < cv.size(); ++i)
                if (c->vertex2dof[cv[i]] < 0) c->vertex2dof[cv[i]] = n_dofs++;
        } else {
            std::vector<int> first(n_vert, 0x7fffffff);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < n_cv; ++i) {
                int* f = &first[cv[i]];
                int cur = __atomic_load_n(f, __ATOMIC_RELAXED);
                while ((int) i < cur && !__atomic_compare_exchange_n(f, &cur, (int) i, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
            }
            const int nt = omp_get_max_threads();
            std::vector<long> part(nt + 1, 0);
#pragma omp parallel num_threads(nt)
            {
                const int t = omp_get_thread_num();
                const long a = n_cv * t / nt, b = n_cv * (t + 1) / nt;
                long cnt = 0;
                for (long i = a; i < b; ++i) cnt += (first[cv[i]] == (int) i);
                part[t + 1] = cnt;
#pragma omp barrier
#pragma omp single
                for (int q = 0; q < nt; ++q) part[q + 1] += part[q];
                long run = part[t];
                for (long i = a; i < b; ++i)
                    if (first[cv[i]] == (int) i) c->vertex2dof[cv[i]] = (int) run++;
            }
            n_dofs = (int) part[nt];
        }
    }
    c->n_cols = n_dofs;
    const int n_all = n_dofs;
    if (c->part_n_owned >= 0) n_dofs = c->part_n_owned;       // rows = owned dofs only
    c->n_dofs = n_dofs;
    c->dof2vertex.assign(n_all, -1);
    for (int v = 0; v < n_vert; ++v) c->dof2vertex[c->vertex2dof[v]] = v;
    c->cells_dof.resize(cv.size());
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long) cv.size(); ++i) c->cells_dof[i] = c->vertex2dof[cv[i]];

    laps.lap("numbering");
    // ---- CSR sparsity: row r couples to every dof sharing a cell with it ----
    // One pass: every thread builds the sorted column lists of its contiguous range of rows (static schedule) into
    // its own arena; after the prefix sum of the row lengths each arena is one memcpy into the column array.
    c->rowptr.assign(n_dofs + 1, 0);
    {
        const int nt = omp_get_max_threads();
        std::vector<std::vector<int>> arena(nt);
        std::vector<int> first_row(nt, -1), cnt(n_dofs, 0);
#pragma omp parallel num_threads(nt)
        {
            const int t = omp_get_thread_num();
            std::vector<int>& mine = arena[t];
            mine.reserve((size_t) n_dofs / nt * 30 + 64);
            std::vector<int> buf, stamp(n_all, -1);          // stamp[d] == r: dof d is already in the list of row r
#pragma omp for schedule(static)
            for (int r = 0; r < n_dofs; ++r) {
                if (first_row[t] < 0) first_row[t] = r;
                const int v = c->dof2vertex[r];
                buf.clear();
                for (int q = v2c_off[v]; q < v2c_off[v + 1]; ++q) {
                    const int* cd = &c->cells_dof[8 * (size_t) v2c[q]];
                    for (int k = 0; k < 8; ++k)
                        if (stamp[cd[k]] != r) { stamp[cd[k]] = r; buf.push_back(cd[k]); }
                }
                std::sort(buf.begin(), buf.end());
                cnt[r] = (int) buf.size();
                mine.insert(mine.end(), buf.begin(), buf.end());
            }
        }
        long tot = 0;
        for (int r = 0; r < n_dofs; ++r) { tot += cnt[r]; if (tot > 2147483647L) return c->fail(FB_ERR_MESH, "nnz exceeds 32-bit index range"); c->rowptr[r + 1] = (int) tot; }
        c->nnz = tot;
        c->col.resize(tot);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int t = 0; t < nt; ++t)
            if (first_row[t] >= 0) std::copy(arena[t].begin(), arena[t].end(), c->col.begin() + c->rowptr[first_row[t]]);
    }

    laps.lap("sparsity");
    // ---- Dirichlet candidate dofs per boundary id (interpolate_boundary_values) ----
    std::vector<unsigned char> on_cu(n_all, 0), on_top(n_all, 0);
    for (const auto& bf : c->bfaces)
        for (int k = 0; k < 4; ++k) {
            const int d = c->cells_dof[8 * (size_t) bf.cell + FACE_VERTS[bf.face][k]];
            // "copper" = the Dirichlet set of the mesh kind: copper_surface for the vacuum mesh, copper_bottom for the bulk
            if (bf.id == (c->mesh_kind ? 7 : 2)) on_cu[d] = 1;
            if (bf.id == 8) on_top[d] = 1;
        }
    c->copper_dofs.clear(); c->top_dofs.clear();
    for (int d = 0; d < n_all; ++d) { if (on_cu[d]) c->copper_dofs.push_back(d); if (on_top[d]) c->top_dofs.push_back(d); }
    std::vector<int>().swap(c->h_cv); std::vector<int>().swap(c->h_v2c_off); std::vector<int>().swap(c->h_v2c);
    std::vector<unsigned char>().swap(c->h_isb);
    c->mesh_ok = true;
    return FB_OK;
}

// Mesh hand-off with UNCHANGED topology (the reference decides between "skip" and "re-mesh" on the rmsd of
// the atoms, ProjectRunaway.cpp:55-67 / GeneralProject.cpp:19-55 -- a host code that moves nodes without re-meshing hands
// over the same connectivity with new coordinates): when node count, hexahedra and markers are identical to the mesh
// already held, everything integer -- vertex compaction, cell orientation, boundary faces, DoF numbering, sparsity,
// Dirichlet sets, and with them the block-JDS tables, persistent-CG slices and device copies -- stays valid.  Only the
// geometry is refreshed and re-validated: no cell may have changed orientation and every boundary face must keep its id
// (mark_boundary works on face centres, DealSolver.cpp:460-518).  Returns false (nothing modified but c->xyz untouched)
// when the full import has to run.
bool fb_host_try_reuse(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex, int mesh_kind) {
    if (!c->mesh_ok || c->mesh_reuse == 0 || c->part_n_owned >= 0 || c->mesh_flipped) return false;
    if (c->fe_degree != 1 || c->imported_degree != 1) return false;      // FE_Q(2) systems are always rebuilt
    if (mesh_kind != c->imported_kind || n_nodes != c->n_nodes || n_hex != c->n_hex) return false;
    if (memcmp(hex_marker, c->hex_marker.data(), sizeof(int) * (size_t) n_hex) != 0) return false;
    if (memcmp(hex8, c->hex8.data(), sizeof(int) * 8 * (size_t) n_hex) != 0) return false;
    Laps laps("import (topology unchanged)");
    const int n_cells = c->n_cells;
    auto vertex = [&](int ce, int k) { return &xyz[3 * (size_t) c->vert2node[c->dof2vertex[c->cells_dof[8 * (size_t) ce + k]]]]; };
    long n_neg = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_neg)
    for (int ce = 0; ce < n_cells; ++ce) {
        P3 v[8];
        for (int k = 0; k < 8; ++k) { const double* p = vertex(ce, k); v[k] = {p[0], p[1], p[2]}; }
        if (hex_volume(v) < 0) ++n_neg;
    }
    if (n_neg > 0) return false;
    laps.lap("orientation");
    double mx[3] = {-1e16, -1e16, -1e16}, mn[3] = {1e16, 1e16, 1e16};
    std::vector<double> ctr(3 * c->bfaces.size());
    for (size_t i = 0; i < c->bfaces.size(); ++i) {
        double s[3] = {0, 0, 0};
        for (int k = 0; k < 4; ++k) { const double* p = vertex(c->bfaces[i].cell, FACE_VERTS[c->bfaces[i].face][k]); s[0] += p[0]; s[1] += p[1]; s[2] += p[2]; }
        for (int d = 0; d < 3; ++d) { ctr[3 * i + d] = s[d] / 4.0; mx[d] = std::max(mx[d], ctr[3 * i + d]); mn[d] = std::min(mn[d], ctr[3 * i + d]); }
    }
    const double eps = 1e-6;
    auto on = [&](double v, double b) { return std::fabs(v - b) <= eps; };
    for (size_t i = 0; i < c->bfaces.size(); ++i) {
        const double* p = &ctr[3 * i];
        int id;
        const bool side = on(p[0], mn[0]) || on(p[0], mx[0]) || on(p[1], mn[1]) || on(p[1], mx[1]);
        if (c->mesh_kind == 0) id = side ? 4 : (on(p[2], mx[2]) ? 8 : 2);
        else id = side ? 4 : (on(p[2], mx[2]) ? 2 : (on(p[2], mn[2]) ? 7 : 2));
        if (id != c->bfaces[i].id) return false;
    }
    laps.lap("face ids");
    c->xyz.assign(xyz, xyz + 3 * (size_t) n_nodes);
    for (int d = 0; d < 3; ++d) { c->bb_mn[d] = mn[d]; c->bb_mx[d] = mx[d]; }
    return true;
}

int fb_host_import_mesh(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    c->part_n_owned = -1;
    const int rc = fb_host_import_phase1(c, xyz, n_nodes, hex8, hex_marker, n_hex);
    return rc ? rc : fb_host_import_phase2(c);
}

// tria->n_active_lines(): distinct vertex pairs along the 12 edges of every cell (to_str(), DealSolver.h:107-117)
long fb_host_count_edges(const fb_ctx* c) {
    static const int E[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    const size_t nc = (size_t) c->n_cells;
    std::vector<uint64_t> keys(12 * nc);
#pragma omp parallel for schedule(static)
    for (long ce = 0; ce < (long) nc; ++ce)
        for (int e = 0; e < 12; ++e) {
            const uint64_t a = (uint64_t) c->cells_dof[8 * (size_t) ce + E[e][0]], b = (uint64_t) c->cells_dof[8 * (size_t) ce + E[e][1]];
            keys[12 * (size_t) ce + e] = a < b ? (a << 32 | b) : (b << 32 | a);
        }
    std::sort(keys.begin(), keys.end());
    return (long) (std::unique(keys.begin(), keys.end()) - keys.begin());
}

// vertex2cell / vertex2node of DealSolver::calc_vertex2dof (DealSolver.cpp:317-341): cells in order, the last one wins
void fb_host_vertex_lastcell(const fb_ctx* c, std::vector<int>& out) {
    out.assign(c->n_vert, 0);
    for (int ce = 0; ce < c->n_cells; ++ce)
        for (int k = 0; k < 8; ++k) out[c->dof2vertex[c->cells_dof[8 * (size_t) ce + k]]] = 8 * ce + k;
}

// Row blocks of the streaming SpMV: whole rows, <= chunk non-zeros and <= maxrows rows per block.
// Returns false when a single row is longer than the chunk (the per-row kernel must be used).
bool fb_host_row_blocks(fb_ctx* c, int chunk, int maxrows) {
    c->rowblk.clear(); c->rowblk.push_back(0);
    int start = 0;
    for (int r = 0; r < c->n_dofs; ++r) {
        if (c->rowptr[r + 1] - c->rowptr[r] > chunk) { c->n_rowblk = 0; return false; }
        if (c->rowptr[r + 1] - c->rowptr[start] > chunk || r - start >= maxrows) { c->rowblk.push_back(r); start = r; }
    }
    c->rowblk.push_back(c->n_dofs);
    c->n_rowblk = (int) c->rowblk.size() - 1;
    c->rowblk_chunk = chunk; c->rowblk_maxrows = maxrows;
    return true;
}

// Column windows of the windowed streaming SpMV: for every row block the sorted list of distinct
// columns it references (its "window" of the input vector) and, per non-zero, the 16-bit position
// of its column inside that window.  Returns false if a window exceeds max_window entries.
bool fb_host_col_windows(fb_ctx* c, int max_window) {
    const int nb = c->n_rowblk;
    if (nb <= 0) return false;
    std::vector<std::vector<int>> win(nb);
    c->col16.resize(c->nnz);
    bool ok = true;
#pragma omp parallel
    {
        std::vector<int> buf;
#pragma omp for schedule(dynamic, 64)
        for (int b = 0; b < nb; ++b) {
            const int k0 = c->rowptr[c->rowblk[b]], k1 = c->rowptr[c->rowblk[b + 1]];
            buf.assign(c->col.begin() + k0, c->col.begin() + k1);
            std::sort(buf.begin(), buf.end());
            buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
            if ((int) buf.size() > max_window) {
#pragma omp atomic write
                ok = false;
                continue;
            }
            for (int k = k0; k < k1; ++k)
                c->col16[k] = (unsigned short) (std::lower_bound(buf.begin(), buf.end(), c->col[k]) - buf.begin());
            win[b] = buf;
        }
    }
    if (!ok) { c->col16.clear(); c->col16.shrink_to_fit(); return false; }
    c->win_off.assign(nb + 1, 0);
    int wmax = 0;
    for (int b = 0; b < nb; ++b) {
        c->win_off[b + 1] = c->win_off[b] + (int) win[b].size();
        wmax = std::max(wmax, (int) win[b].size());
    }
    c->win_list.resize(c->win_off[nb]);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) std::copy(win[b].begin(), win[b].end(), c->win_list.begin() + c->win_off[b]);
    c->win_max = wmax;
    return true;
}

// Block-JDS tables of the HBM-roofline SpMV (k_spmv_jds).  Rows are cut into blocks of R consecutive
// rows; inside a block the rows are (stably) sorted by decreasing length and stored as jagged
// diagonals: diagonal j holds the j-th entry of every row longer than j, so that thread t of the
// CTA walks "its" row with perfectly coalesced loads and NO padding.  Columns are replaced by 16-bit
// positions inside the block's window (the sorted distinct columns the block touches).
bool fb_host_jds_build(fb_ctx* c, int R, int max_window, bool sym, int split, int pad) {
    // sym: only the strictly lower triangle (columns < row: a prefix of every sorted CSR row) is stored; the window
    // list then holds the distinct columns BELOW the block (sorted), and the block's own rows follow implicitly:
    // window position of column j is  rank(j) for j < r0,  n_ext + (j - r0) for r0 <= j < row.
    // split > 0 (full layout only): a row longer than `split` entries is stored as ceil(len / split) SEGMENTS of (nearly)
    // equal length, each in a slot of its own, chained by jds_link; the kernel adds the segment sums in chain order.  A
    // block then holds R slots instead of R rows (jds_rowbeg gives its first row).  Without it the one warp that owns the
    // longest rows of a block walks 2-4x more diagonals than the other seven, which wait for it at the block's barrier.
    const int n = c->n_dofs;
    if (sym) split = 0;
    if (split == 0 || pad != 8) pad = 2;       // entries every diagonal is padded to (8: 16-byte aligned slices of the column stream, spmv_kernel 308)
    c->jds_pad = pad;
    auto rowlen = [&](int r) {
        const int* lo = c->col.data() + c->rowptr[r]; const int* hi = c->col.data() + c->rowptr[r + 1];
        return sym ? (int) (std::lower_bound(lo, hi, r) - lo) : (int) (hi - lo);
    };
    auto nseg_of = [&](int len) { return (split > 0 && len > split) ? (len + split - 1) / split : 1; };
    std::vector<int>& rb = c->jds_rowbeg;
    rb.clear(); rb.push_back(0);
    if (split > 0) {
        int used = 0;
        for (int r = 0; r < n; ++r) {
            const int k = nseg_of(c->rowptr[r + 1] - c->rowptr[r]);
            if (k > R) return false;
            if (used + k > R) { rb.push_back(r); used = 0; }
            used += k;
        }
    } else {
        for (long r = R; r < n; r += R) rb.push_back((int) r);
    }
    rb.push_back(n);
    const int nb = (int) rb.size() - 1;
    c->jds_R = R; c->jds_nb = nb; c->jds_sym = sym; c->jds_split = split;
    c->jds_perm.assign((size_t) nb * R, 0xFFFF); c->jds_len.assign((size_t) nb * R, 0); c->jds_slot.assign(n, 0);
    c->jds_link.assign(split > 0 ? (size_t) nb * R : 0, 0xFFFF);
    std::vector<unsigned short> seg_start(split > 0 ? (size_t) nb * R : 0, 0);      // first entry (within its row) of the slot's segment
    c->jds_jdp.assign(nb + 1, 0); c->jds_base.assign(nb + 1, 0);
    std::vector<int> maxlen(nb, 0), nslot(nb, 0);
    // pass 1: per-block slot order (longest first) and jagged-diagonal counts
#pragma omp parallel
    {
        std::vector<int> ord, lens, loc, off, first;
#pragma omp for schedule(static)
        for (int b = 0; b < nb; ++b) {
            const int r0 = rb[b], nr = rb[b + 1] - r0;
            lens.clear(); loc.clear(); off.clear(); first.assign(nr + 1, 0);
            for (int a = 0; a < nr; ++a) {
                const int len = rowlen(r0 + a), k = nseg_of(len), sl = (len + k - 1) / k;
                first[a] = (int) lens.size();
                for (int q = 0; q < k; ++q) { lens.push_back(std::max(0, std::min(sl, len - q * sl))); loc.push_back(a); off.push_back(q * sl); }
            }
            first[nr] = (int) lens.size();
            const int nv = (int) lens.size();
            ord.resize(nv);
            std::iota(ord.begin(), ord.end(), 0);
            std::stable_sort(ord.begin(), ord.end(), [&](int a, int q) { return lens[a] > lens[q]; });
            std::vector<int> slot_of(nv);
            for (int t = 0; t < nv; ++t) slot_of[ord[t]] = t;
            for (int t = 0; t < nv; ++t) {
                const int v = ord[t], a = loc[v];
                const bool head = (v == first[a]);
                c->jds_perm[(size_t) b * R + t] = (unsigned short) (a | (head ? 0 : 0x8000));
                c->jds_len[(size_t) b * R + t] = (unsigned short) std::min(lens[v], 65535);
                if (head) c->jds_slot[r0 + a] = (unsigned short) t;
                if (split > 0) {
                    seg_start[(size_t) b * R + t] = (unsigned short) off[v];
                    if (v + 1 < first[a + 1]) c->jds_link[(size_t) b * R + t] = (unsigned short) slot_of[v + 1];
                }
            }
            maxlen[b] = nv ? lens[ord[0]] : 0;
            nslot[b] = nv;
        }
    }
    for (int b = 0; b < nb; ++b) {
        if (maxlen[b] > 60000) return false;
        c->jds_jdp[b + 1] = c->jds_jdp[b] + maxlen[b] + 1;
    }
    c->jds_jd.assign(c->jds_jdp[nb], 0);
    // every diagonal is padded to an even number of entries (zero value, window position 0) so that a
    // thread can fetch the entries of two neighbouring slots with one 16-byte load
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
        int* jd = &c->jds_jd[c->jds_jdp[b]];
        int t_active = nslot[b];
        jd[0] = 0;
        for (int j = 0; j < maxlen[b]; ++j) {
            while (t_active > 0 && (int) c->jds_len[(size_t) b * R + t_active - 1] <= j) --t_active;
            // (segmented layout: multiples of 8, so that any run of diagonals is a 16-byte aligned slice of the 2-byte column stream)
            jd[j + 1] = jd[j] + ((t_active + pad - 1) & ~(pad - 1));
        }
    }
    for (int b = 0; b < nb; ++b) {
        long next = (long) c->jds_base[b] + c->jds_jd[c->jds_jdp[b + 1] - 1];
        if (next > 2147483000L) return false;
        c->jds_base[b + 1] = (int) next;
    }
    c->jds_size = c->jds_base[nb];
    std::vector<std::vector<int>> win(nb);
    c->col16.assign((size_t) c->jds_size + 8, 0);
    bool ok = true;
#pragma omp parallel
    {
        std::vector<int> buf;
#pragma omp for schedule(dynamic, 16)
        for (int b = 0; b < nb; ++b) {
            const int r0 = rb[b], nr = rb[b + 1] - r0;
            const int* jd = &c->jds_jd[c->jds_jdp[b]];      // jd[j] = (padded) entries stored before diagonal j
            const int k0 = c->rowptr[r0], k1 = c->rowptr[r0 + nr];
            const size_t base = (size_t) c->jds_base[b];
            if (sym) {
                buf.clear();
                for (int k = k0; k < k1; ++k) if (c->col[k] < r0) buf.push_back(c->col[k]);
            } else {
                buf.assign(c->col.begin() + k0, c->col.begin() + k1);
            }
            std::sort(buf.begin(), buf.end());
            buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
            if ((int) buf.size() + (sym ? R : 0) > max_window) {
#pragma omp atomic write
                ok = false;
                continue;
            }
            const int n_ext = (int) buf.size();
            for (int t = 0; t < nslot[b]; ++t) {
                const int r = r0 + (c->jds_perm[(size_t) b * R + t] & 0x7FFF);
                const int len = (int) c->jds_len[(size_t) b * R + t];
                const int kfirst = c->rowptr[r] + (split > 0 ? (int) seg_start[(size_t) b * R + t] : 0);
                // the columns of a row ascend, so do their window positions: gallop from the previous position instead of
                // a binary search of the whole window per non-zero (6e8 x 10 steps on the 2.3e7-DoF mesh)
                const int* wb = buf.data(); int pos = 0;
                for (int k = kfirst; k < kfirst + len; ++k) {
                    const int cj = c->col[k];
                    int w;
                    if (sym && cj >= r0) w = n_ext + (cj - r0);
                    else {
                        int step = 1, hi = pos;
                        while (hi < n_ext && wb[hi] < cj) { pos = hi + 1; hi += step; step <<= 1; }
                        hi = std::min(hi, n_ext);
                        pos = (int) (std::lower_bound(wb + pos, wb + hi, cj) - wb);
                        w = pos;
                    }
                    c->col16[base + jd[k - kfirst] + t] = (unsigned short) w;
                }
            }
            win[b] = buf;
        }
    }
    if (!ok) return false;
    c->win_off.assign(nb + 1, 0);
    int wmax = 0;
    for (int b = 0; b < nb; ++b) {
        c->win_off[b + 1] = c->win_off[b] + (int) win[b].size();
        wmax = std::max(wmax, (int) win[b].size());
    }
    c->win_list.resize(c->win_off[nb]);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) std::copy(win[b].begin(), win[b].end(), c->win_list.begin() + c->win_off[b]);
    c->win_max = wmax;
    c->jds_maxlen = *std::max_element(maxlen.begin(), maxlen.end());
    if (getenv("FB_VERBOSE"))
        fprintf(stderr, "[fb] block-JDS%s: R=%d, %d blocks, %ld stored entries (%.2f per row), window avg %.0f max %d, longest %s %d\n",
                sym ? " (symmetric, lower triangle)" : (split > 0 ? " (rows split into segments)" : ""), R, nb, (long) c->jds_size,
                (double) c->jds_size / std::max(1, n), (double) c->win_off[nb] / std::max(1, nb), wmax, split > 0 ? "segment" : "row", c->jds_maxlen);
    return true;
}

// =======================================================================================
//  Interpolator precompute tables.  The arithmetic below must reproduce the reference's
//  tables bit for bit (cell location is compared bit-exactly), hence the expression order
//  of src/InterpolatorCells.cpp is kept and this file is compiled with -ffp-contract=off.
// =======================================================================================
namespace {

struct V { double x, y, z; };
inline V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V operator*(V a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V operator/(V a, double s) { return {a.x / s, a.y / s, a.z / s}; }

// 3x3 with a unit last column / plain 3x3 (InterpolatorCells.cpp:470-480)
inline double d2(V a, V b) { return a.x * (b.y - b.z) - a.y * (b.x - b.z) + a.z * (b.x - b.y); }
inline double d3(V a, V b, V c) {
    return a.x * (b.y * c.z - c.y * b.z) - b.x * (a.y * c.z - c.y * a.z) + c.x * (a.y * b.z - b.y * a.z);
}

}  // namespace

void fb_host_interp_tables(const fb_ctx* c, const int* node_marker, const int* tet4, const int* tet_nbr4,
                           const int* tet_marker, int n_tet, const int* tri3, const double* tri_norm3, int n_tri,
                           const int* quad4, int n_quad, fb_interp_tables& T) {
    const int n_nodes = c->n_nodes, n_hex = c->n_hex;
    auto node = [&](int i) { return V{c->xyz[3 * (size_t) i], c->xyz[3 * (size_t) i + 1], c->xyz[3 * (size_t) i + 2]}; };
    Laps laps("interp tables");

    // ---------------- tetrahedra (:523-629) ----------------
    T.tet.resize(n_tet); T.tet_cent.resize(3 * (size_t) n_tet); T.tet_mark.resize(n_tet);
    std::vector<int> n2t_off(n_nodes + 1, 0);
    for (long i = 0; i < 4L * n_tet; ++i) n2t_off[tet4[i] + 1]++;
    for (int i = 0; i < n_nodes; ++i) n2t_off[i + 1] += n2t_off[i];
    std::vector<int> n2t(4 * (size_t) n_tet), pos(n2t_off.begin(), n2t_off.end() - 1);
    for (int t = 0; t < n_tet; ++t) for (int k = 0; k < 4; ++k) n2t[pos[tet4[4 * t + k]]++] = t;

    std::vector<std::vector<int>> nbr(n_tet);
#pragma omp parallel for schedule(dynamic, 256)
    for (int t = 0; t < n_tet; ++t) {
        const int* nn = &tet_nbr4[4 * (size_t) t];
        std::vector<int>& L = nbr[t];
        auto push_unique = [&](int v) { if (std::find(L.begin(), L.end(), v) == L.end()) L.push_back(v); };
        // face neighbours in TetGen order, then vertex-sharing tets; duplicates dropped (first
        // occurrence kept) -- the reference list (:549-560) holds repeats, which never change the
        // first hit of locate_cell
        for (int k = 0; k < 4; ++k) if (nn[k] >= 0) push_unique(nn[k]);
        for (int k = 0; k < 4; ++k) {
            const int nd = tet4[4 * t + k];
            for (int q = n2t_off[nd]; q < n2t_off[nd + 1]; ++q) {
                const int o = n2t[q];
                if (o != t && o != nn[0] && o != nn[1] && o != nn[2] && o != nn[3]) push_unique(o);
            }
        }
        const V v1 = node(tet4[4 * t]), v2 = node(tet4[4 * t + 1]), v3 = node(tet4[4 * t + 2]), v4 = node(tet4[4 * t + 3]);
        V ctr = {0, 0, 0};
        ctr = ctr + v1; ctr = ctr + v2; ctr = ctr + v3; ctr = ctr + v4;
        ctr = ctr * (1.0 / 4);
        T.tet_cent[3 * (size_t) t] = ctr.x; T.tet_cent[3 * (size_t) t + 1] = ctr.y; T.tet_cent[3 * (size_t) t + 2] = ctr.z;
        fb::TetRec& R = T.tet[t];
        {   // 4x4 determinant with a unit last column (:483-489)
            const double e1 = d3(v2, v3, v4), e2 = d3(v1, v3, v4), e3 = d3(v1, v2, v4), e4 = d3(v1, v2, v3);
            R.det0 = 1.0 / (e4 - e3 + e2 - e1);
        }
        double a, b, cc, d;
        a = d2({v2.y, v3.y, v4.y}, {v2.z, v3.z, v4.z}); b = d2({v2.x, v3.x, v4.x}, {v2.z, v3.z, v4.z});
        cc = d2({v2.x, v3.x, v4.x}, {v2.y, v3.y, v4.y}); d = d3({v2.x, v3.x, v4.x}, {v2.y, v3.y, v4.y}, {v2.z, v3.z, v4.z});
        R.d[0][0] = a; R.d[0][1] = -b; R.d[0][2] = cc; R.d[0][3] = -d;
        a = d2({v1.y, v3.y, v4.y}, {v1.z, v3.z, v4.z}); b = d2({v1.x, v3.x, v4.x}, {v1.z, v3.z, v4.z});
        cc = d2({v1.x, v3.x, v4.x}, {v1.y, v3.y, v4.y}); d = d3({v1.x, v3.x, v4.x}, {v1.y, v3.y, v4.y}, {v1.z, v3.z, v4.z});
        R.d[1][0] = -a; R.d[1][1] = b; R.d[1][2] = -cc; R.d[1][3] = d;
        a = d2({v1.y, v2.y, v4.y}, {v1.z, v2.z, v4.z}); b = d2({v1.x, v2.x, v4.x}, {v1.z, v2.z, v4.z});
        cc = d2({v1.x, v2.x, v4.x}, {v1.y, v2.y, v4.y}); d = d3({v1.x, v2.x, v4.x}, {v1.y, v2.y, v4.y}, {v1.z, v2.z, v4.z});
        R.d[2][0] = a; R.d[2][1] = -b; R.d[2][2] = cc; R.d[2][3] = -d;
        a = d2({v1.y, v2.y, v3.y}, {v1.z, v2.z, v3.z}); b = d2({v1.x, v2.x, v3.x}, {v1.z, v2.z, v3.z});
        cc = d2({v1.x, v2.x, v3.x}, {v1.y, v2.y, v3.y}); d = d3(v1, v2, v3);
        R.d[3][0] = -a; R.d[3][1] = b; R.d[3][2] = -cc; R.d[3][3] = d;
        T.tet_mark[t] = tet_marker[t] != 3;       // narrow_search_to(TYPES.VACUUM) (:730-732)
    }
    T.tet_nbr_off.assign(n_tet + 1, 0);
    for (int t = 0; t < n_tet; ++t) T.tet_nbr_off[t + 1] = T.tet_nbr_off[t] + (int) nbr[t].size();
    T.tet_nbr.resize(T.tet_nbr_off[n_tet]);
    for (int t = 0; t < n_tet; ++t) std::copy(nbr[t].begin(), nbr[t].end(), T.tet_nbr.begin() + T.tet_nbr_off[t]);

    laps.lap("tets");
    // ---------------- hexahedra (:1205-1267) ----------------
    T.hex.resize(n_hex);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < n_hex; ++h) {
        V x[8];
        for (int k = 0; k < 8; ++k) x[k] = node(c->hex8[8 * (size_t) h + k]);
        const V x1 = x[0], x2 = x[1], x3 = x[2], x4 = x[3], x5 = x[4], x6 = x[5], x7 = x[6], x8 = x[7];
        V f[8];
        f[0] = (x1 + x2 + x3 + x4 + x5 + x6 + x7 + x8) / 8.0;
        f[1] = ((x1 * -1) + x2 + x3 - x4 - x5 + x6 + x7 - x8) / 8.0;
        f[2] = ((x1 * -1) - x2 + x3 + x4 - x5 - x6 + x7 + x8) / 8.0;
        f[3] = ((x1 * -1) - x2 - x3 - x4 + x5 + x6 + x7 + x8) / 8.0;
        f[4] = (x1 - x2 + x3 - x4 + x5 - x6 + x7 - x8) / 8.0;
        f[5] = (x1 - x2 - x3 + x4 - x5 + x6 + x7 - x8) / 8.0;
        f[6] = (x1 + x2 - x3 - x4 - x5 - x6 + x7 + x8) / 8.0;
        f[7] = ((x1 * -1) + x2 - x3 + x4 + x5 - x6 + x7 - x8) / 8.0;
        for (int k = 0; k < 8; ++k) { T.hex[h].f[k][0] = f[k].x; T.hex[h].f[k][1] = f[k].y; T.hex[h].f[k][2] = f[k].z; }
    }

    laps.lap("hexs");
    // ---------------- triangles (:1585-1637) ----------------
    T.tri.resize(n_tri); T.tri_cent.resize(3 * (size_t) n_tri);
    std::vector<std::vector<int>> n2r(n_nodes);
    for (int t = 0; t < n_tri; ++t) for (int k = 0; k < 3; ++k) n2r[tri3[3 * t + k]].push_back(t);
    std::vector<std::vector<int>> rnbr(n_tri);
    for (int t = 0; t < n_tri; ++t) {
        std::vector<int>& L = rnbr[t];
        for (int k = 0; k < 3; ++k)
            for (int o : n2r[tri3[3 * t + k]])
                if (o != t && std::find(L.begin(), L.end(), o) == L.end()) L.push_back(o);
        const V v0 = node(tri3[3 * t]), v1 = node(tri3[3 * t + 1]), v2 = node(tri3[3 * t + 2]);
        const V nrm = {tri_norm3[3 * t], tri_norm3[3 * t + 1], tri_norm3[3 * t + 2]};
        const V e1 = v1 - v0, e2 = v2 - v0;
        const V pv = {nrm.y * e2.z - nrm.z * e2.y, nrm.z * e2.x - nrm.x * e2.z, nrm.x * e2.y - nrm.y * e2.x};
        const double i_det = 1.0 / (e1.x * pv.x + e1.y * pv.y + e1.z * pv.z);
        fb::TriRec& R = T.tri[t];
        const V e1s = e1 * i_det, pvs = pv * i_det;
        R.vert0[0] = v0.x; R.vert0[1] = v0.y; R.vert0[2] = v0.z;
        R.edge1[0] = e1s.x; R.edge1[1] = e1s.y; R.edge1[2] = e1s.z;
        R.edge2[0] = e2.x; R.edge2[1] = e2.y; R.edge2[2] = e2.z;
        R.pvec[0] = pvs.x; R.pvec[1] = pvs.y; R.pvec[2] = pvs.z;
        R.norm[0] = nrm.x; R.norm[1] = nrm.y; R.norm[2] = nrm.z;
        R.maxd = std::sqrt(e2.x * e2.x + e2.y * e2.y + e2.z * e2.z);
        const V ctr = (v0 + v1 + v2) / 3.0;           // TetgenFaces::calc_appendices, TetgenCells.cpp:402
        T.tri_cent[3 * (size_t) t] = ctr.x; T.tri_cent[3 * (size_t) t + 1] = ctr.y; T.tri_cent[3 * (size_t) t + 2] = ctr.z;
    }
    T.tri_nbr_off.assign(n_tri + 1, 0);
    for (int t = 0; t < n_tri; ++t) T.tri_nbr_off[t + 1] = T.tri_nbr_off[t] + (int) rnbr[t].size();
    T.tri_nbr.resize(T.tri_nbr_off[n_tri]);
    for (int t = 0; t < n_tri; ++t) std::copy(rnbr[t].begin(), rnbr[t].end(), T.tri_nbr.begin() + T.tri_nbr_off[t]);

    laps.lap("tris");
    // ---------------- quadratic cells (:1151-1173, :1873-1895) ----------------
    auto common = [](const int* a, int na, const int* b, int nb) {
        for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) if (a[i] == b[j]) return a[i];
        return -1;
    };
    T.qtet.assign(10 * (size_t) n_tet, 0);
    for (int t = 0; t < n_tet; ++t) {
        if (n_hex <= t) continue;
        int en[4][8], ne[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 8; ++k) {
                const int hn = c->hex8[8 * (size_t) (4 * t + i) + k];
                if (node_marker[hn] == 2) en[i][ne[i]++] = hn;       // TYPES.EDGECENTROID
            }
        int* q = &T.qtet[10 * (size_t) t];
        for (int k = 0; k < 4; ++k) q[k] = tet4[4 * t + k];
        q[4] = common(en[0], ne[0], en[1], ne[1]); q[5] = common(en[1], ne[1], en[2], ne[2]);
        q[6] = common(en[2], ne[2], en[0], ne[0]); q[7] = common(en[0], ne[0], en[3], ne[3]);
        q[8] = common(en[1], ne[1], en[3], ne[3]); q[9] = common(en[2], ne[2], en[3], ne[3]);
    }
    T.qtri.assign(6 * (size_t) n_tri, 0);
    for (int f = 0; f < n_tri; ++f) {
        if (n_quad == 0) continue;
        int en[3][4], ne[3] = {0, 0, 0};
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 4; ++k) {
                const int qn = quad4[4 * (size_t) (3 * f + i) + k];
                if (node_marker[qn] == 2) en[i][ne[i]++] = qn;
            }
        int* q = &T.qtri[6 * (size_t) f];
        for (int k = 0; k < 3; ++k) q[k] = tri3[3 * f + k];
        q[3] = common(en[0], ne[0], en[1], ne[1]); q[4] = common(en[1], ne[1], en[2], ne[2]); q[5] = common(en[2], ne[2], en[0], ne[0]);
    }

    laps.lap("quadratic");
    // ---------------- node -> (vacuum hex, local node) CSR (Interpolator.cpp:60-76) ----------------
    T.n2c_off.assign(n_nodes + 1, 0);
    for (int h = 0; h < n_hex; ++h)
        if (c->hex_marker[h] > 0) for (int k = 0; k < 8; ++k) T.n2c_off[c->hex8[8 * (size_t) h + k] + 1]++;
    for (int i = 0; i < n_nodes; ++i) T.n2c_off[i + 1] += T.n2c_off[i];
    T.n2c_list.resize(T.n2c_off[n_nodes]);
    std::vector<int> p2(T.n2c_off.begin(), T.n2c_off.end() - 1);
    for (int h = 0; h < n_hex; ++h)
        if (c->hex_marker[h] > 0) for (int k = 0; k < 8; ++k) T.n2c_list[p2[c->hex8[8 * (size_t) h + k]]++] = 8 * h + k;
}
