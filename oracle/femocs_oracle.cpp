// ============================================================================
//  TEST INFRASTRUCTURE -- CPU ORACLE.  NOT PART OF THE PRODUCT.
//
//  A plain, single-threaded C++ restatement of the reference's per-step
//  electrostatic hot path, used ONLY as the checker in
//  tests/, __graft_entry__.smoke() and as the timed CPU baseline in bench.py.
//  femocs_b200/ never includes, links or calls anything in this file.
//
//  PARITY STATUS
//   * Interpolation half (cell location, shape functions, nodal-field
//     extraction, smoothing, particle weights/gradients): PINNED.  Checked
//     bit-for-bit / to round-off against the reference's own code compiled from
//     /root/reference (oracle/_ref/libfemocs_ref.so, tests/test_oracle_vs_ref.py)
//     and against the committed fixtures in tests/golden/ produced by that code.
//   * Solver half (DoF numbering, Q1 assembly, Neumann RHS, Dirichlet
//     elimination, SSOR-preconditioned CG): **PARITY UNPINNED**.  The arithmetic
//     lives in deal.II 9.2.0 (reference build/makefile.defs:6,
//     build/CMakeLists.txt:89), which is neither vendored in /root/reference nor
//     installed in this image, and the reference has no test or golden vector
//     for it.  What is restated below is deal.II 9.2's published algorithm
//     (FE_Q(1), QGauss(2), MappingQ1, MatrixTools::apply_boundary_values,
//     SolverCG, SparseMatrix::precondition_SSOR) anchored on the reference's call
//     sites; it is validated mathematically in tests/test_oracle_solver.py
//     (uniform-field exactness, symmetry, zero row sums, manufactured solution).
//
//  Every function cites the reference file:line it follows (paths relative to
//  /root/reference).
// ============================================================================

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <omp.h>
#include <set>
#include <string>
#include <vector>

namespace {

struct V3 { double x = 0, y = 0, z = 0; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          // Primitives.h:307
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }  // Primitives.h:310-312
inline double comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
inline double& compref(V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
inline double dist2(V3 a, V3 b) {                                                     // Primitives.h:247-252
    double xx = a.x - b.x, yy = a.y - b.y, zz = a.z - b.z;
    return xx * xx + yy * yy + zz * zz;
}
struct V4 { double x = 0, y = 0, z = 0, w = 0; };
inline double dot4(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }  // Primitives.h:347

struct Sol { V3 v; double s1 = 0, s2 = 0; };   // Primitives.h:507-522 (vector, scalar1, scalar2)

constexpr double ZERO = 1e-15;                 // InterpolatorCells.h:241

// InterpolatorCells.cpp:470-495 -- the determinant helpers, same expression order
double det2(V3 v1, V3 v2) { return v1.x * (v2.y - v2.z) - v1.y * (v2.x - v2.z) + v1.z * (v2.x - v2.y); }
double det3(V3 v1, V3 v2, V3 v3) {
    return v1.x * (v2.y * v3.z - v3.y * v2.z) - v2.x * (v1.y * v3.z - v3.y * v1.z)
         + v3.x * (v1.y * v2.z - v2.y * v1.z);
}
double det4(V3 v1, V3 v2, V3 v3, V3 v4) {
    const double d1 = det3(v2, v3, v4);
    const double d2 = det3(v1, v3, v4);
    const double d3 = det3(v1, v2, v4);
    const double d4 = det3(v1, v2, v3);
    return d4 - d3 + d2 - d1;
}

enum { BID_COPPER = 2, BID_SIDES = 4, BID_TOP = 8, BID_BOTTOM = 7 };   // Globals.h:60-68 (copper_surface, sides, vacuum_top, copper_bottom)
enum { NODE_TET = 1, NODE_EDGE = 2, NODE_FACE = 3, NODE_TETCENTROID = 4 };  // Globals.h:53-56
constexpr int TYPE_VACUUM = 3;  // Globals.h TYPES.VACUUM (tet marker of vacuum tets)

struct Oracle {
    // ---------------- mesh arrays as TetgenMesh hands them out (SURVEY 8a') ----------------
    int n_nodes = 0, n_hex = 0, n_tet = 0, n_tri = 0, n_quad = 0;
    std::vector<V3> xyz;
    std::vector<int> node_marker, hex8, hex_marker, tet4, tet_nbr, tet_marker, tri3, tri2tet, quad4, quad2hex;
    std::vector<V3> tri_norm;
    double tet_edgemax = 1, tri_edgemax = 1;
    std::vector<int> voro_off, voro_list;

    // ---------------- solver (DealSolver / PoissonSolver) ----------------
    std::vector<int> node2vert, vert2node;       // InterpolatorCells.cpp:38-66 / delete_unused_vertices
    std::vector<int> hex2cell, cell2hex;         // InterpolatorCells.cpp:1247-1266
    std::vector<std::array<int, 8>> cells;       // deal.II lexicographic vertex order, compact vertex ids
    std::vector<int> vertex2dof, dof2vertex;     // DealSolver.cpp:317-341
    int n_dofs = 0;
    struct BFace { int cell, face, id; };
    std::vector<BFace> bfaces;                   // boundary faces in deal.II cell/face iteration order
    std::vector<int> rowptr, col;
    std::vector<double> val, val_save, rhs, sol;
    std::map<int, double> boundary_values;       // DealSolver.h:147
    double applied_field = 0, applied_potential = 0;
    int anode_dirichlet = 0;
    double last_res = 0;
    int write_time = 0; std::vector<double> charge_density;   // PoissonSolver.cpp:196-207
    // FE_Q(2) variant (the reference built with DealSolver.h:130 shape_degree = 2, hence QGauss(3): north star config 2)
    int fe_degree = 1;
    std::vector<std::array<int, 27>> cdofs;      // dofs of a cell, local node (i, j, k) in {0, 1, 2}^3 at index i + 3 j + 9 k
    // ---------------- CurrentHeatSolver (bulk mesh: mesh_kind == 1) ----------------
    int mesh_kind = 0;                           // 0 = vacuum hexes (PoissonSolver), 1 = bulk hexes (CurrentHeatSolver)
    std::vector<double> ch_current, ch_heat;     // CurrentSolver::solution, HeatSolver::solution (dof order)
    std::vector<double> res_T, res_rho;          // PhysicalQuantities::resistivity_data
    double lorentz = 2.44e-8, ch_T_ambient = 300.0;

    // ---------------- interpolator (Interpolator / InterpolatorCells) ----------------
    std::vector<Sol> nodal;                      // InterpolatorNodes::solutions
    std::vector<std::vector<std::pair<int, int>>> node2cells;   // Interpolator.cpp:60-76
    // tets (LinearTetrahedra)
    std::vector<double> t_det0; std::vector<V4> t_det[4];
    std::vector<V3> t_cent; std::vector<int> t_mark; std::vector<std::vector<int>> t_nbr;
    // hexes (LinearHexahedra)
    std::vector<V3> h_f[8], h_cent;
    // tris (LinearTriangles)
    std::vector<V3> r_vert0, r_edge1, r_edge2, r_pvec, r_norm, r_cent; std::vector<double> r_maxd;
    std::vector<std::vector<int>> r_nbr;
    // quads
    std::vector<V3> q_cent;
    // quadratic cells
    std::vector<std::array<int, 10>> qtet; std::vector<std::array<int, 6>> qtri;
    double decay_factor = -1;
};

// ============================================================================
//  Solver half  (deal.II semantics restated -- see header: parity unpinned)
// ============================================================================

// Reference cube quantities for FE_Q(1) + QGauss(2), lexicographic (deal.II 9.2 FE_Q / QGauss;
// reference: DealSolver.h:130-131 shape_degree=1, quadrature_degree=2)
struct RefCube {
    double N[8][8];      // N[q][i]
    double dN[8][8][3];  // dN[q][i][d]  on [0,1]^3
    double w[8];
    RefCube() {
        const double g[2] = {0.5 * (1.0 - 1.0 / std::sqrt(3.0)), 0.5 * (1.0 + 1.0 / std::sqrt(3.0))};
        for (int q = 0; q < 8; ++q) {
            const double xi[3] = {g[q & 1], g[(q >> 1) & 1], g[(q >> 2) & 1]};
            w[q] = 0.125;
            for (int i = 0; i < 8; ++i) {
                double f[3], df[3];
                for (int d = 0; d < 3; ++d) {
                    const int bit = (i >> d) & 1;
                    f[d] = bit ? xi[d] : 1.0 - xi[d];
                    df[d] = bit ? 1.0 : -1.0;
                }
                N[q][i] = f[0] * f[1] * f[2];
                dN[q][i][0] = df[0] * f[1] * f[2];
                dN[q][i][1] = f[0] * df[1] * f[2];
                dN[q][i][2] = f[0] * f[1] * df[2];
            }
        }
    }
};
const RefCube REFCUBE;

// deal.II GeometryInfo<3>: vertices of face f in lexicographic order
const int FACE_VERTS[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};

V3 cell_vertex(const Oracle& o, int cell, int v) { return o.xyz[o.vert2node[o.cells[cell][v]]]; }

// MappingQ1 at one quadrature point: JxW and physical gradients (deal.II MappingQGeneric, FEValues)
void cell_geometry(const Oracle& o, int cell, int q, double& JxW, V3 grad[8]) {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // J[d][e] = dx_d / dxi_e
    for (int i = 0; i < 8; ++i) {
        const V3 p = cell_vertex(o, cell, i);
        for (int e = 0; e < 3; ++e) {
            J[0][e] += p.x * REFCUBE.dN[q][i][e];
            J[1][e] += p.y * REFCUBE.dN[q][i][e];
            J[2][e] += p.z * REFCUBE.dN[q][i][e];
        }
    }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                     - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                     + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    double inv[3][3];   // inv = J^{-1}: inv[e][d] = dxi_e / dx_d
    inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    JxW = det * REFCUBE.w[q];
    for (int i = 0; i < 8; ++i) {
        const double* g = REFCUBE.dN[q][i];
        grad[i].x = g[0] * inv[0][0] + g[1] * inv[1][0] + g[2] * inv[2][0];
        grad[i].y = g[0] * inv[0][1] + g[1] * inv[1][1] + g[2] * inv[2][1];
        grad[i].z = g[0] * inv[0][2] + g[1] * inv[1][2] + g[2] * inv[2][2];
    }
}

// TetgenCells.cpp:673-686 (export_vacuum), DealSolver.cpp:191-209 (import_mesh),
// :460-518 (mark_boundary), PoissonSolver.cpp:52-55 (mark_mesh), DealSolver.cpp:368-387 (setup_system)
// kind 1: TetgenCells.cpp:688-701 (export_bulk: marker < 0) and CurrentHeatSolver.cpp:526-530 mark_mesh
// (top -> copper_surface, bottom -> copper_bottom, sides -> copper_sides, other -> copper_surface)
void q2_distribute(Oracle& o);
int import_mesh(Oracle& o, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker_in, int n_hex, int kind = 0) {
    o.mesh_kind = kind;
    std::vector<int> sel(n_hex);                 // > 0: the hexahedron belongs to this solver's mesh
    for (int h = 0; h < n_hex; ++h) sel[h] = kind ? (hex_marker_in[h] < 0) : (hex_marker_in[h] > 0);
    const int* hex_marker = sel.data();
    o.n_nodes = n_nodes; o.n_hex = n_hex;
    o.xyz.resize(n_nodes);
    for (int i = 0; i < n_nodes; ++i) o.xyz[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    o.hex8.assign(hex8, hex8 + 8 * (size_t) n_hex);
    o.hex_marker.assign(hex_marker_in, hex_marker_in + n_hex);

    // GridTools::delete_unused_vertices: order-preserving compaction (InterpolatorCells.cpp:50-65)
    o.node2vert.assign(n_nodes, -1);
    for (int h = 0; h < n_hex; ++h)
        if (hex_marker[h] > 0)
            for (int k = 0; k < 8; ++k) o.node2vert[hex8[8 * h + k]] = -2;
    o.vert2node.clear();
    for (int i = 0; i < n_nodes; ++i)
        if (o.node2vert[i] == -2) { o.node2vert[i] = (int) o.vert2node.size(); o.vert2node.push_back(i); }

    // cells: vacuum hexes in order; UCD -> lexicographic (create_triangulation_compatibility,
    // deal.II GeometryInfo<3>::ucd_to_deal = {0,1,5,4,2,3,7,6}; confirmed by
    // InterpolatorCells.cpp:1355-1358 shape_funs_dealii)
    static const int ucd_to_deal[8] = {0, 1, 5, 4, 2, 3, 7, 6};
    o.hex2cell.assign(n_hex, -1); o.cell2hex.clear(); o.cells.clear();
    for (int h = 0; h < n_hex; ++h)
        if (hex_marker[h] > 0) {
            std::array<int, 8> c;
            for (int k = 0; k < 8; ++k) c[ucd_to_deal[k]] = o.node2vert[hex8[8 * h + k]];
            o.hex2cell[h] = (int) o.cells.size();
            o.cell2hex.push_back(h);
            o.cells.push_back(c);
        }
    const int n_cells = (int) o.cells.size();
    if (n_cells == 0) return 1;

    // GridReordering::invert_all_cells_of_negative_grid (DealSolver.cpp:198): if the cells have
    // negative measure, swap vertex i <-> i+4 of the old-style numbering.  Tethex guarantees positive measure
    // (Tethex.cpp:1565-1592), so this is expected to be a no-op; it is decided per grid.
    {
        int n_neg = 0;
        for (int c = 0; c < n_cells; ++c) {
            double JxW; V3 g[8]; double vol = 0;
            for (int q = 0; q < 8; ++q) { cell_geometry(o, c, q, JxW, g); vol += JxW; }
            if (vol < 0) ++n_neg;
        }
        if (n_neg == n_cells)       // swap of UCD vertices i <-> i + 4 (deal.II 9.2 grid_reordering.cc) = lexicographic (0,2) (1,3) (4,6) (5,7)
            for (auto& c : o.cells) for (int k : {0, 1, 4, 5}) std::swap(c[k], c[k + 2]);
        else if (n_neg > 0) return 2;   // deal.II would throw -> import_mesh returns false
    }

    // boundary faces: faces owned by exactly one cell (sorted key list instead of a std::map: the X meshes of the
    // benchmark have 1e7 faces; same result)
    auto face_key = [&](int c, int f) {
        std::array<int, 4> k;
        for (int v = 0; v < 4; ++v) k[v] = o.cells[c][FACE_VERTS[f][v]];
        std::sort(k.begin(), k.end());
        return k;
    };
    std::vector<unsigned char> is_boundary(6 * (size_t) n_cells, 0);
    {
        struct FK { std::array<int, 4> k; int cf; };
        std::vector<FK> fk(6 * (size_t) n_cells);
#pragma omp parallel for schedule(static)
        for (int c = 0; c < n_cells; ++c)
            for (int f = 0; f < 6; ++f) fk[6 * (size_t) c + f] = {face_key(c, f), 6 * c + f};
        std::sort(fk.begin(), fk.end(), [](const FK& a, const FK& b) { return a.k < b.k; });
        for (size_t i = 0; i < fk.size();) {
            size_t j = i + 1;
            while (j < fk.size() && fk[j].k == fk[i].k) ++j;
            if (j - i == 1) is_boundary[fk[i].cf] = 1;
            i = j;
        }
    }

    // DealSolver.cpp:460-518 mark_boundary: face centre = mean of the 4 vertices (TriaAccessor::center)
    auto face_center = [&](int c, int f) {
        V3 s;
        for (int v = 0; v < 4; ++v) s = s + cell_vertex(o, c, FACE_VERTS[f][v]);
        return s / 4.0;
    };
    const double eps = 1e-6;
    double xmax = -1e16, ymax = -1e16, zmax = -1e16, xmin = 1e16, ymin = 1e16, zmin = 1e16;
    o.bfaces.clear();
    for (int c = 0; c < n_cells; ++c)
        for (int f = 0; f < 6; ++f)
            if (is_boundary[6 * (size_t) c + f]) {
                o.bfaces.push_back({c, f, 0});
                const V3 p = face_center(c, f);
                xmax = std::max(xmax, p.x); xmin = std::min(xmin, p.x);
                ymax = std::max(ymax, p.y); ymin = std::min(ymin, p.y);
                zmax = std::max(zmax, p.z); zmin = std::min(zmin, p.z);
            }
    auto on_b = [&](double v, double b) { return std::fabs(v - b) <= eps; };   // Macros.cpp:168-174
    for (auto& bf : o.bfaces) {
        const V3 p = face_center(bf.cell, bf.face);
        if (on_b(p.x, xmin) || on_b(p.x, xmax) || on_b(p.y, ymin) || on_b(p.y, ymax)) bf.id = BID_SIDES;
        else if (on_b(p.z, zmax)) bf.id = kind ? BID_COPPER : BID_TOP;
        else if (on_b(p.z, zmin)) bf.id = kind ? BID_BOTTOM : BID_COPPER;   // "bottom" = copper_surface, PoissonSolver.cpp:52-55
        else bf.id = BID_COPPER;                       // "other"  = copper_surface
    }

    // DoFHandler::distribute_dofs for FE_Q(1): first touch, cells in order, local vertices 0..7
    const int n_vert = (int) o.vert2node.size();
    o.vertex2dof.assign(n_vert, -1);
    o.n_dofs = 0;
    for (int c = 0; c < n_cells; ++c)
        for (int v = 0; v < 8; ++v)
            if (o.vertex2dof[o.cells[c][v]] < 0) o.vertex2dof[o.cells[c][v]] = o.n_dofs++;
    o.dof2vertex.assign(o.n_dofs, -1);
    for (int v = 0; v < n_vert; ++v) o.dof2vertex[o.vertex2dof[v]] = v;

    // DoFTools::make_sparsity_pattern: all dof pairs sharing a cell (columns kept sorted).  Built row by row from
    // the dof -> cells adjacency (gather, sort, unique) instead of one std::set per row: same pattern, minutes faster
    // on the benchmark meshes.
    {
        std::vector<int> d2c_off(o.n_dofs + 1, 0);
        for (int c = 0; c < n_cells; ++c) for (int i = 0; i < 8; ++i) ++d2c_off[o.vertex2dof[o.cells[c][i]] + 1];
        for (int r = 0; r < o.n_dofs; ++r) d2c_off[r + 1] += d2c_off[r];
        std::vector<int> d2c(d2c_off[o.n_dofs]), fill(d2c_off.begin(), d2c_off.end() - 1);
        for (int c = 0; c < n_cells; ++c) for (int i = 0; i < 8; ++i) d2c[fill[o.vertex2dof[o.cells[c][i]]]++] = c;
        o.rowptr.assign(o.n_dofs + 1, 0);
        std::vector<std::vector<int>> rows(o.n_dofs);
#pragma omp parallel for schedule(dynamic, 4096)
        for (int r = 0; r < o.n_dofs; ++r) {
            std::vector<int>& row = rows[r];
            row.reserve(8 * (size_t) (d2c_off[r + 1] - d2c_off[r]));
            for (int k = d2c_off[r]; k < d2c_off[r + 1]; ++k)
                for (int j = 0; j < 8; ++j) row.push_back(o.vertex2dof[o.cells[d2c[k]][j]]);
            std::sort(row.begin(), row.end());
            row.erase(std::unique(row.begin(), row.end()), row.end());
            row.shrink_to_fit();
        }
        for (int r = 0; r < o.n_dofs; ++r) o.rowptr[r + 1] = o.rowptr[r] + (int) rows[r].size();
        o.col.resize(o.rowptr[o.n_dofs]);
#pragma omp parallel for schedule(static)
        for (int r = 0; r < o.n_dofs; ++r) std::copy(rows[r].begin(), rows[r].end(), o.col.begin() + o.rowptr[r]);
    }
    o.val.assign(o.col.size(), 0.0);
    o.val_save.assign(o.col.size(), 0.0);
    o.rhs.assign(o.n_dofs, 0.0);
    o.sol.assign(o.n_dofs, 0.0);
    o.boundary_values.clear();
    if (o.fe_degree == 2) q2_distribute(o);
    return 0;
}

inline int csr_pos(const Oracle& o, int r, int c) {
    auto b = o.col.begin() + o.rowptr[r], e = o.col.begin() + o.rowptr[r + 1];
    auto it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? (int) (it - o.col.begin()) : -1;
}


// ============================================================================
//  FE_Q(2) variant: what the reference's solver does when DealSolver.h:130 reads shape_degree = 2 (quadrature_degree
//  = shape_degree + 1 = 3 follows, :131).  Same call sites as above (setup_system, assemble_parallel, assemble_rhs,
//  append_dirichlet, calc_vertex2dof); deal.II 9.2 FE_Q(2) (tensor-product Lagrange basis on the support points
//  0, 1/2, 1), QGauss(3), MappingQ1 (trilinear geometry from the 8 vertices).  Parity unpinned like the rest of the
//  solver half; validated in tests/test_oracle_q2.py (independent numpy derivation, linear exactness, order of
//  convergence above FE_Q(1)).
// ============================================================================
const double Q2_GP[3] = {0.5 - 0.5 * std::sqrt(0.6), 0.5, 0.5 + 0.5 * std::sqrt(0.6)};   // QGauss<1>(3) on [0, 1]
const double Q2_GW[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
inline void lagrange2(double x, double L[3], double dL[3]) {
    L[0] = 2.0 * (x - 0.5) * (x - 1.0); L[1] = 4.0 * x * (1.0 - x); L[2] = 2.0 * x * (x - 0.5);
    dL[0] = 4.0 * x - 3.0; dL[1] = 4.0 - 8.0 * x; dL[2] = 4.0 * x - 1.0;
}
// local nodes in the order deal.II numbers the dofs of a cell (dof_handler_policy.cc distribute_dofs_on_cell: vertices,
// lines, quads, hex; GeometryInfo<3> line / face numbering), each as (i, j, k) with 0 / 2 = the end points, 1 = the middle
const int Q2_ORDER[27][3] = {
    {0, 0, 0}, {2, 0, 0}, {0, 2, 0}, {2, 2, 0}, {0, 0, 2}, {2, 0, 2}, {0, 2, 2}, {2, 2, 2},                    // vertices 0..7
    {0, 1, 0}, {2, 1, 0}, {1, 0, 0}, {1, 2, 0}, {0, 1, 2}, {2, 1, 2}, {1, 0, 2}, {1, 2, 2},                    // lines 0..7
    {0, 0, 1}, {2, 0, 1}, {0, 2, 1}, {2, 2, 1},                                                                // lines 8..11
    {0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}, {1, 1, 0}, {1, 1, 2},                                          // quads 0..5
    {1, 1, 1}};                                                                                                // hex
// the 9 local nodes of face f, (a, b) lexicographic in the two free directions (the order of FACE_VERTS at the corners)
inline void q2_face_nodes(int f, int nodes[9]) {
    const int fixed = f / 2, val = (f % 2) * 2;
    const int d0 = fixed == 0 ? 1 : 0, d1 = fixed == 2 ? 1 : 2;
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) {
            int ijk[3]; ijk[fixed] = val; ijk[d0] = a; ijk[d1] = b;
            nodes[a + 3 * b] = ijk[0] + 3 * ijk[1] + 9 * ijk[2];
        }
}

// DoFHandler::distribute_dofs for FE_Q(2) + DoFTools::make_sparsity_pattern (DealSolver.cpp:368-387) and
// calc_vertex2dof (:317-341: vertex_dof_index -> only the vertex dofs are exported)
void q2_distribute(Oracle& o) {
    const int n_cells = (int) o.cells.size(), n_vert = (int) o.vert2node.size();
    o.vertex2dof.assign(n_vert, -1);
    std::map<std::pair<int, int>, int> line_dof;
    std::map<std::array<int, 4>, int> quad_dof;
    o.cdofs.assign(n_cells, {});
    o.n_dofs = 0;
    for (int c = 0; c < n_cells; ++c)
        for (int l = 0; l < 27; ++l) {
            const int* ijk = Q2_ORDER[l];
            // the vertices of the entity (vertex / line / quad / hex) this node sits on
            int ent[8], ne = 0;
            for (int v = 0; v < 8; ++v) {
                const int b[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
                bool on = true;
                for (int d = 0; d < 3; ++d) if (ijk[d] != 1 && ijk[d] != 2 * b[d]) on = false;
                if (on) ent[ne++] = o.cells[c][v];
            }
            std::sort(ent, ent + ne);
            int* slot = nullptr; int hexdof = -1;
            if (ne == 1) slot = &o.vertex2dof[ent[0]];
            else if (ne == 2) slot = &line_dof.emplace(std::make_pair(ent[0], ent[1]), -1).first->second;
            else if (ne == 4) slot = &quad_dof.emplace(std::array<int, 4>{ent[0], ent[1], ent[2], ent[3]}, -1).first->second;
            else slot = &hexdof;
            if (*slot < 0) *slot = o.n_dofs++;
            o.cdofs[c][ijk[0] + 3 * ijk[1] + 9 * ijk[2]] = *slot;
        }
    o.dof2vertex.assign(o.n_dofs, -1);
    for (int v = 0; v < n_vert; ++v) o.dof2vertex[o.vertex2dof[v]] = v;
    std::vector<int> d2c_off(o.n_dofs + 1, 0);
    for (int c = 0; c < n_cells; ++c) for (int i = 0; i < 27; ++i) ++d2c_off[o.cdofs[c][i] + 1];
    for (int r = 0; r < o.n_dofs; ++r) d2c_off[r + 1] += d2c_off[r];
    std::vector<int> d2c(d2c_off[o.n_dofs]), fill(d2c_off.begin(), d2c_off.end() - 1);
    for (int c = 0; c < n_cells; ++c) for (int i = 0; i < 27; ++i) d2c[fill[o.cdofs[c][i]]++] = c;
    o.rowptr.assign(o.n_dofs + 1, 0);
    std::vector<std::vector<int>> rows(o.n_dofs);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int r = 0; r < o.n_dofs; ++r) {
        std::vector<int>& row = rows[r];
        for (int k = d2c_off[r]; k < d2c_off[r + 1]; ++k)
            for (int j = 0; j < 27; ++j) row.push_back(o.cdofs[d2c[k]][j]);
        std::sort(row.begin(), row.end());
        row.erase(std::unique(row.begin(), row.end()), row.end());
    }
    for (int r = 0; r < o.n_dofs; ++r) o.rowptr[r + 1] = o.rowptr[r] + (int) rows[r].size();
    o.col.resize(o.rowptr[o.n_dofs]);
    for (int r = 0; r < o.n_dofs; ++r) std::copy(rows[r].begin(), rows[r].end(), o.col.begin() + o.rowptr[r]);
    o.val.assign(o.col.size(), 0.0); o.val_save.assign(o.col.size(), 0.0);
    o.rhs.assign(o.n_dofs, 0.0); o.sol.assign(o.n_dofs, 0.0);
}

// PoissonSolver.cpp:213-263 with FE_Q(2) / QGauss(3): K_ab = sum_q JxW_q grad phi_a . grad phi_b
void q2_assemble_matrix(Oracle& o) {
    const int n_cells = (int) o.cells.size();
    std::vector<double> Ke(27 * 27);
    for (int c = 0; c < n_cells; ++c) {
        std::fill(Ke.begin(), Ke.end(), 0.0);
        V3 xv[8];
        for (int v = 0; v < 8; ++v) xv[v] = cell_vertex(o, c, v);
        for (int q = 0; q < 27; ++q) {
            const double xi[3] = {Q2_GP[q % 3], Q2_GP[(q / 3) % 3], Q2_GP[q / 9]};
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};      // MappingQ1: J[d][e] = dx_d / dxi_e
            for (int v = 0; v < 8; ++v) {
                double f[3], df[3];
                for (int d = 0; d < 3; ++d) { const int bit = (v >> d) & 1; f[d] = bit ? xi[d] : 1.0 - xi[d]; df[d] = bit ? 1.0 : -1.0; }
                const double dn[3] = {df[0] * f[1] * f[2], f[0] * df[1] * f[2], f[0] * f[1] * df[2]};
                for (int e = 0; e < 3; ++e) { J[0][e] += xv[v].x * dn[e]; J[1][e] += xv[v].y * dn[e]; J[2][e] += xv[v].z * dn[e]; }
            }
            const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                             + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            double inv[3][3];
            inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
            inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
            inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
            inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
            inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
            const double JxW = det * Q2_GW[q % 3] * Q2_GW[(q / 3) % 3] * Q2_GW[q / 9];
            double L[3][3], dL[3][3];
            for (int d = 0; d < 3; ++d) lagrange2(xi[d], L[d], dL[d]);
            V3 g[27];
            for (int a = 0; a < 27; ++a) {
                const int i = a % 3, j = (a / 3) % 3, k = a / 9;
                const double r[3] = {dL[0][i] * L[1][j] * L[2][k], L[0][i] * dL[1][j] * L[2][k], L[0][i] * L[1][j] * dL[2][k]};
                g[a].x = r[0] * inv[0][0] + r[1] * inv[1][0] + r[2] * inv[2][0];
                g[a].y = r[0] * inv[0][1] + r[1] * inv[1][1] + r[2] * inv[2][1];
                g[a].z = r[0] * inv[0][2] + r[1] * inv[1][2] + r[2] * inv[2][2];
            }
            for (int a = 0; a < 27; ++a) for (int b = 0; b < 27; ++b) Ke[27 * a + b] += JxW * dot(g[a], g[b]);
        }
        for (int a = 0; a < 27; ++a)
            for (int b = 0; b < 27; ++b) o.val[csr_pos(o, o.cdofs[c][a], o.cdofs[c][b])] += Ke[27 * a + b];
    }
    o.val_save = o.val;
}

// DealSolver.cpp:389-430 assemble_rhs(bid) with FE_Q(2) face shape functions and QGauss<2>(3)
void q2_assemble_rhs_faces(Oracle& o, int bid) {
    for (const auto& bf : o.bfaces) {
        if (bf.id != bid) continue;
        V3 p[4];
        for (int v = 0; v < 4; ++v) p[v] = cell_vertex(o, bf.cell, FACE_VERTS[bf.face][v]);
        int nodes[9]; q2_face_nodes(bf.face, nodes);
        double cell_rhs[9] = {};
        for (int q = 0; q < 9; ++q) {
            const double s = Q2_GP[q % 3], t = Q2_GP[q / 3];
            const V3 ds = (p[1] - p[0]) * (1 - t) + (p[3] - p[2]) * t;
            const V3 dt = (p[2] - p[0]) * (1 - s) + (p[3] - p[1]) * s;
            const V3 n = cross(ds, dt);
            const double JxW = std::sqrt(dot(n, n)) * Q2_GW[q % 3] * Q2_GW[q / 3];
            double Ls[3], Lt[3], dummy[3];
            lagrange2(s, Ls, dummy); lagrange2(t, Lt, dummy);
            for (int i = 0; i < 9; ++i) cell_rhs[i] += Ls[i % 3] * Lt[i / 3] * o.applied_field * JxW;
        }
        for (int i = 0; i < 9; ++i) o.rhs[o.cdofs[bf.cell][nodes[i]]] += cell_rhs[i];
    }
}

// DealSolver.cpp:432-435 append_dirichlet: interpolate_boundary_values touches every dof on the face (9 for FE_Q(2))
void q2_append_dirichlet(Oracle& o, int bid, double value) {
    for (const auto& bf : o.bfaces)
        if (bf.id == bid) {
            int nodes[9]; q2_face_nodes(bf.face, nodes);
            for (int i = 0; i < 9; ++i) o.boundary_values[o.cdofs[bf.cell][nodes[i]]] = value;
        }
}

// PoissonSolver.cpp:162-167 setup + DealSolver.cpp:368-387 setup_system
void setup(Oracle& o, double field, double potential, int anode_dirichlet) {
    o.boundary_values.clear();
    std::fill(o.val.begin(), o.val.end(), 0.0);
    std::fill(o.val_save.begin(), o.val_save.end(), 0.0);
    std::fill(o.rhs.begin(), o.rhs.end(), 0.0);
    std::fill(o.sol.begin(), o.sol.end(), 0.0);      // solution = dirichlet_bc_value (0)
    o.applied_field = field; o.applied_potential = potential; o.anode_dirichlet = anode_dirichlet;
}

// PoissonSolver.cpp:213-263 assemble_parallel / assemble_local_cell, DealSolver.cpp:64-72 copy_global_cell
void assemble_matrix(Oracle& o) {
    if (o.fe_degree == 2) { q2_assemble_matrix(o); return; }
    const int n_cells = (int) o.cells.size();
    for (int c = 0; c < n_cells; ++c) {
        double Ke[8][8] = {};
        for (int q = 0; q < 8; ++q) {
            double JxW; V3 g[8];
            cell_geometry(o, c, q, JxW, g);
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j)
                    Ke[i][j] += JxW * dot(g[i], g[j]);      // loop order q -> i -> j, PoissonSolver.cpp:249-255
        }
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j)
                o.val[csr_pos(o, o.vertex2dof[o.cells[c][i]], o.vertex2dof[o.cells[c][j]])] += Ke[i][j];
    }
    o.val_save = o.val;                                      // PoissonSolver.cpp:231-232
}

// DealSolver.cpp:389-430 assemble_rhs(bid) with get_face_bc = applied_field (PoissonSolver.cpp:152-154)
// face_bc != nullptr: EmissionSolver::get_face_bc = (*bc_values)[boundary_face_index++] (CurrentHeatSolver.h:53-56)
void assemble_rhs_faces(Oracle& o, int bid, const double* face_bc = nullptr) {
    if (o.fe_degree == 2) { q2_assemble_rhs_faces(o, bid); return; }
    const double g[2] = {0.5 * (1.0 - 1.0 / std::sqrt(3.0)), 0.5 * (1.0 + 1.0 / std::sqrt(3.0))};
    int boundary_face_index = 0;
    for (const auto& bf : o.bfaces) {
        if (bf.id != bid) continue;
        const double bc_value = face_bc ? face_bc[boundary_face_index++] : o.applied_field;
        V3 p[4];
        for (int v = 0; v < 4; ++v) p[v] = cell_vertex(o, bf.cell, FACE_VERTS[bf.face][v]);
        double cell_rhs[4] = {0, 0, 0, 0};
        for (int q = 0; q < 4; ++q) {     // QGauss<2>(2), weights 1/4
            const double s = g[q & 1], t = g[q >> 1];
            const double N[4] = {(1 - s) * (1 - t), s * (1 - t), (1 - s) * t, s * t};
            const V3 ds = (p[1] - p[0]) * (1 - t) + (p[3] - p[2]) * t;
            const V3 dt = (p[2] - p[0]) * (1 - s) + (p[3] - p[1]) * s;
            const V3 n = cross(ds, dt);
            const double JxW = std::sqrt(dot(n, n)) * 0.25;
            for (int i = 0; i < 4; ++i) cell_rhs[i] += N[i] * bc_value * JxW;
        }
        for (int i = 0; i < 4; ++i)
            o.rhs[o.vertex2dof[o.cells[bf.cell][FACE_VERTS[bf.face][i]]]] += cell_rhs[i];
    }
}

// DealSolver.cpp:432-435 append_dirichlet (VectorTools::interpolate_boundary_values, ConstantFunction)
void append_dirichlet(Oracle& o, int bid, double value) {
    if (o.fe_degree == 2) { q2_append_dirichlet(o, bid, value); return; }
    for (const auto& bf : o.bfaces)
        if (bf.id == bid)
            for (int v = 0; v < 4; ++v)
                o.boundary_values[o.vertex2dof[o.cells[bf.cell][FACE_VERTS[bf.face][v]]]] = value;
}

// DealSolver.cpp:437-440 apply_dirichlet -> MatrixTools::apply_boundary_values(bv, A, x, b, true)
void apply_dirichlet(Oracle& o) {
    double first_nonzero_diag = 1;
    for (int r = 0; r < o.n_dofs; ++r) {
        const double d = o.val[csr_pos(o, r, r)];
        if (d != 0) { first_nonzero_diag = d; break; }
    }
    for (const auto& bv : o.boundary_values) {
        const int dof = bv.first;
        const int pd = csr_pos(o, dof, dof);
        for (int k = o.rowptr[dof]; k < o.rowptr[dof + 1]; ++k)
            if (k != pd) o.val[k] = 0;
        double new_rhs;
        if (o.val[pd] != 0) new_rhs = bv.second * o.val[pd];
        else { o.val[pd] = first_nonzero_diag; new_rhs = bv.second * first_nonzero_diag; }
        o.rhs[dof] = new_rhs;
        const double diag = o.val[pd];
        for (int k = o.rowptr[dof]; k < o.rowptr[dof + 1]; ++k) {
            if (k == pd) continue;
            const int row = o.col[k];                 // symmetric sparsity
            const int p = csr_pos(o, row, dof);
            o.rhs[row] -= o.val[p] / diag * new_rhs;
            o.val[p] = 0;
        }
        o.sol[dof] = bv.second;
    }
}

// ---- Newton map to natural coordinates & trilinear weights (used by the solver's charge RHS) ----
// InterpolatorCells.cpp:1272-1317 project_to_nat_coords(point, hex)
void hex_nat_coords(const Oracle& o, V3 point, int hex, double& u, double& v, double& w) {
    const V3 f0 = point - o.h_f[0][hex];
    const V3 f1 = o.h_f[1][hex], f2 = o.h_f[2][hex], f3 = o.h_f[3][hex], f4 = o.h_f[4][hex],
             f5 = o.h_f[5][hex], f6 = o.h_f[6][hex], f7 = o.h_f[7][hex];
    u = 0; v = 0; w = 0;
    for (int i = 0; i < 20; ++i) {    // n_newton_iterations = 20, InterpolatorCells.h
        const V3 f = (f0 - f1 * u - f2 * v - f3 * w - f4 * (u * v) - f5 * (u * w) - f6 * (v * w) - f7 * (u * v * w));
        const V3 fu = f1 + f4 * v + f5 * w + f7 * (v * w);
        const V3 fv = f2 + f4 * u + f6 * w + f7 * (u * w);
        const V3 fw = f3 + f5 * u + f6 * v + f7 * (u * v);
        double D = det3(fu, fv, fw);
        D = 1.0 / D;
        const double du = det3(f, fv, fw) * D;
        const double dv = det3(fu, f, fw) * D;
        const double dw = det3(fu, fv, f) * D;
        u += du; v += dv; w += dw;
        if (du * du + dv * dv + dw * dw < ZERO) return;
    }
}

// InterpolatorCells.cpp:1334-1353 shape_functions(point, hex)
void hex_shape_functions(const Oracle& o, V3 point, int hex, double sf[8]) {
    double u, v, w;
    hex_nat_coords(o, point, hex, u, v, w);
    sf[0] = (1 - u) * (1 - v) * (1 - w) / 8.0;
    sf[1] = (1 + u) * (1 - v) * (1 - w) / 8.0;
    sf[2] = (1 + u) * (1 + v) * (1 - w) / 8.0;
    sf[3] = (1 - u) * (1 + v) * (1 - w) / 8.0;
    sf[4] = (1 - u) * (1 - v) * (1 + w) / 8.0;
    sf[5] = (1 + u) * (1 - v) * (1 + w) / 8.0;
    sf[6] = (1 + u) * (1 + v) * (1 + w) / 8.0;
    sf[7] = (1 - u) * (1 + v) * (1 + w) / 8.0;
}

// InterpolatorCells.cpp:1355-1358 shape_funs_dealii
void hex_shape_functions_dealii(const Oracle& o, V3 point, int hex, double sf[8]) {
    double s[8];
    hex_shape_functions(o, point, hex, s);
    const int perm[8] = {0, 1, 4, 5, 3, 2, 7, 6};
    for (int i = 0; i < 8; ++i) sf[i] = s[perm[i]];
}

// PoissonSolver.cpp:299-319 assemble_space_charge_fast; particle.cell is a solver (deal) cell index
void precompute_hexs(Oracle& o);
void assemble_space_charge(Oracle& o, const double* pxyz, const int* pcell, long n, double charge_factor) {
    // the reference's PoissonSolver holds a pointer to the interpolator's LinearHexahedra (PoissonSolver.cpp:44-49); a
    // solver-only oracle (benchmark meshes without tetrahedra) builds the same f0..f7 coefficients from the hexahedra
    if ((int) o.h_f[0].size() != o.n_hex) precompute_hexs(o);
    for (long p = 0; p < n; ++p) {
        const int cell = pcell[p];
        if (cell < 0 || cell >= (int) o.cells.size()) continue;   // lost particles are cleared before (Pic.cpp:146)
        double sf[8];
        hex_shape_functions_dealii(o, {pxyz[3 * p], pxyz[3 * p + 1], pxyz[3 * p + 2]}, o.cell2hex[cell], sf);
        for (int i = 0; i < 8; ++i) o.rhs[o.vertex2dof[o.cells[cell][i]]] += sf[i] * charge_factor;
    }
}

// PoissonSolver.cpp:276-296, the general path taken when shape_degree != 1: DealSolver::shape_funs (DealSolver.cpp:75-110)
// = the 27 FE_Q(2) shape values at the particle's unit-cell coordinates (MappingQ1 inverse, clamped to the unit cell by
// GeometryInfo::project_to_unit_cell).  The unit-cell point is recovered from the trilinear weights of the pinned
// shape_funs_dealii restatement: xi_d = sum of the weights of the vertices with bit d set.
void precompute_hexs(Oracle& o);
void q2_assemble_space_charge(Oracle& o, const double* pxyz, const int* pcell, long n, double charge_factor) {
    if ((int) o.h_f[0].size() != o.n_hex) precompute_hexs(o);
    for (long p = 0; p < n; ++p) {
        const int cell = pcell[p];
        if (cell < 0 || cell >= (int) o.cells.size()) continue;
        double sf[8];
        hex_shape_functions_dealii(o, {pxyz[3 * p], pxyz[3 * p + 1], pxyz[3 * p + 2]}, o.cell2hex[cell], sf);
        double xi[3] = {0, 0, 0};
        for (int v = 0; v < 8; ++v) for (int d = 0; d < 3; ++d) if ((v >> d) & 1) xi[d] += sf[v];
        double L[3][3], dL[3];
        for (int d = 0; d < 3; ++d) { xi[d] = std::min(1.0, std::max(0.0, xi[d])); lagrange2(xi[d], L[d], dL); }
        for (int a = 0; a < 27; ++a) o.rhs[o.cdofs[cell][a]] += L[0][a % 3] * L[1][(a / 3) % 3] * L[2][a / 9] * charge_factor;
    }
}

// PoissonSolver.cpp:170-210 assemble(first_time)
void assemble(Oracle& o, int first_time, const double* pxyz, const int* pcell, long n_parts, double charge_factor) {
    if (first_time) std::fill(o.val.begin(), o.val.end(), 0.0);
    std::fill(o.rhs.begin(), o.rhs.end(), 0.0);
    if (first_time) assemble_matrix(o); else o.val = o.val_save;
    append_dirichlet(o, BID_COPPER, 0.0);
    if (!o.anode_dirichlet) assemble_rhs_faces(o, BID_TOP);
    else append_dirichlet(o, BID_TOP, o.applied_potential);
    if (pxyz && n_parts > 0) {
        if (o.fe_degree == 2) q2_assemble_space_charge(o, pxyz, pcell, n_parts, charge_factor);
        else assemble_space_charge(o, pxyz, pcell, n_parts, charge_factor);
    }
    // PoissonSolver.cpp:196-207: charge density for the files, before the Dirichlet conditions; DealSolver.cpp:344-366 calc_dof_volumes
    o.charge_density.assign(o.n_dofs, 0.0);
    if (o.write_time && o.fe_degree == 1) {
        std::vector<double> dof_volume(o.n_dofs, 0.0);
        for (int c = 0; c < (int) o.cells.size(); ++c)
            for (int q = 0; q < 8; ++q) {
                double JxW; V3 g[8];
                cell_geometry(o, c, q, JxW, g);
                for (int i = 0; i < 8; ++i) dof_volume[o.vertex2dof[o.cells[c][i]]] += JxW;
            }
        for (int d = 0; d < o.n_dofs; ++d) o.charge_density[d] = o.rhs[d] / dof_volume[d];
    }
    apply_dirichlet(o);
}

// SparseMatrix::vmult -- deal.II runs it TBB-parallel over row ranges; here OpenMP over rows
// (row sums keep their serial order, so the result does not depend on the thread count)
void spmv(const Oracle& o, const std::vector<double>& x, std::vector<double>& y) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < o.n_dofs; ++r) {
        double s = 0;
        for (int k = o.rowptr[r]; k < o.rowptr[r + 1]; ++k) s += o.val[k] * x[o.col[k]];
        y[r] = s;
    }
}

// SparseMatrix::precondition_SSOR (deal.II 9.2 sparse_matrix.templates.h), called through
// PreconditionSSOR<>::vmult from DealSolver.cpp:447-449
void precondition_ssor(const Oracle& o, const std::vector<double>& diag, double om,
                       const std::vector<double>& src, std::vector<double>& dst) {
    const int n = o.n_dofs;
    for (int r = 0; r < n; ++r) {
        double s = 0;
        for (int k = o.rowptr[r]; k < o.rowptr[r + 1] && o.col[k] < r; ++k) s += o.val[k] * dst[o.col[k]];
        dst[r] = (src[r] - s * om) / diag[r];
    }
    for (int r = 0; r < n; ++r) dst[r] *= om * (2.0 - om) * diag[r];
    for (int r = n - 1; r >= 0; --r) {
        double s = 0;
        for (int k = o.rowptr[r + 1] - 1; k >= o.rowptr[r] && o.col[k] > r; --k) s += o.val[k] * dst[o.col[k]];
        dst[r] = (dst[r] - s * om) / diag[r];
    }
}

// DealSolver.cpp:442-458 solve_cg: SolverControl(max_iter, tol) + SolverCG<> (deal.II 9.2 solver_cg.h)
// precond: 0 = reference behaviour (SSOR if ssor_param > 0 else identity), 1 = Jacobi (comparison only)
int solve_cg(Oracle& o, int max_iter, double tol, double ssor_param, int precond) {
    const int n = o.n_dofs;
    std::vector<double> g(n), h(n), d(n), diag(n);
    for (int r = 0; r < n; ++r) diag[r] = o.val[csr_pos(o, r, r)];
    auto dotv = [&](const std::vector<double>& a, const std::vector<double>& b) {
        double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s;
    };
    const bool identity = (precond == 0 && !(ssor_param > 0.0));
    auto apply_prec = [&](const std::vector<double>& src, std::vector<double>& dst) {
        if (precond == 1) for (int i = 0; i < n; ++i) dst[i] = src[i] / diag[i];
        else precondition_ssor(o, diag, ssor_param, src, dst);
    };
    bool all_zero = true;
    for (double v : o.sol) if (v != 0) { all_zero = false; break; }
    if (!all_zero) { spmv(o, o.sol, g); for (int i = 0; i < n; ++i) g[i] -= o.rhs[i]; }
    else for (int i = 0; i < n; ++i) g[i] = -o.rhs[i];
    double res = std::sqrt(dotv(g, g));
    o.last_res = res;
    int it = 0;
    if (res <= tol) return 0;                       // SolverControl::check: success first
    if (max_iter <= 0 || std::isnan(res)) return 0; // failure at step 0 -> -0
    double gh;
    if (!identity) { apply_prec(g, h); for (int i = 0; i < n; ++i) d[i] = -h[i]; gh = dotv(g, h); }
    else { for (int i = 0; i < n; ++i) d[i] = -g[i]; gh = res * res; }
    while (true) {
        ++it;
        spmv(o, d, h);
        double alpha = dotv(d, h);
        alpha = gh / alpha;
        for (int i = 0; i < n; ++i) o.sol[i] += alpha * d[i];
        for (int i = 0; i < n; ++i) g[i] += alpha * h[i];
        res = std::sqrt(dotv(g, g));
        o.last_res = res;
        if (res <= tol) return it;
        if (it >= max_iter || std::isnan(res)) return -it;   // NoConvergence -> -last_step (DealSolver.cpp:455-457)
        double beta = gh;
        if (!identity) {
            apply_prec(g, h);
            gh = dotv(g, h);
            beta = gh / beta;
            for (int i = 0; i < n; ++i) d[i] = beta * d[i] - h[i];
        } else {
            gh = res * res;
            beta = gh / beta;
            for (int i = 0; i < n; ++i) d[i] = beta * d[i] - g[i];
        }
    }
}

// ============================================================================
//  CurrentHeatSolver on the bulk mesh.  Same deal.II semantics as above: parity unpinned.
// ============================================================================

// PhysicalQuantities.cpp:173-185 linear_interp over resistivity_data
double pq_linear_interp(const Oracle& o, double x) {
    const int n = (int) o.res_T.size();
    if (x <= o.res_T[0]) return o.res_rho[0];
    if (x >= o.res_T[n - 1]) return o.res_rho[n - 1];
    const int i1 = (int) (std::lower_bound(o.res_T.begin(), o.res_T.end(), x) - o.res_T.begin());
    const int i2 = i1 - 1;
    return o.res_rho[i2] + (o.res_rho[i1] - o.res_rho[i2]) * (x - o.res_T[i2]) / (o.res_T[i1] - o.res_T[i2]);
}
// PhysicalQuantities.cpp:31-37 evaluate_resistivity, :47-50 sigma
double pq_sigma(const Oracle& o, double T) {
    if (T < o.res_T.front()) T = o.res_T.front();
    if (T > o.res_T.back()) T = o.res_T.back();
    return 1.0 / (10. * pq_linear_interp(o, T));
}
// PhysicalQuantities.cpp:57-64 kappa (Wiedemann-Franz)
double pq_kappa(const Oracle& o, double T) {
    if (T < o.res_T.front()) T = o.res_T.front();
    if (T > o.res_T.back()) T = o.res_T.back();
    return o.lorentz * T * pq_sigma(o, T);
}

// CurrentHeatSolver.cpp:509-513 setup(temperature): heat.dirichlet_bc_value = T; both setup_system()
// (DealSolver.cpp:368-387: solution = dirichlet_bc_value, boundary_values cleared)
void ch_setup(Oracle& o, double T_ambient) {
    o.ch_T_ambient = T_ambient;
    o.ch_current.assign(o.n_dofs, 0.0);
    o.ch_heat.assign(o.n_dofs, T_ambient);
    o.boundary_values.clear();
}

// CurrentHeatSolver.cpp:420-449 CurrentSolver::assemble + :451-484 assemble_local_cell (sigma = 1)
void current_assemble(Oracle& o, const double* face_bc) {
    std::fill(o.val.begin(), o.val.end(), 0.0);
    std::fill(o.rhs.begin(), o.rhs.end(), 0.0);
    assemble_matrix(o);
    assemble_rhs_faces(o, BID_COPPER, face_bc);
    o.boundary_values.clear();
    append_dirichlet(o, BID_BOTTOM, 0.0);                  // current.dirichlet_bc_value = 0 (DealSolver.cpp:38)
    o.sol = o.ch_current;
    apply_dirichlet(o);
    o.ch_current = o.sol;
}

// CurrentHeatSolver.cpp:105-152 HeatSolver::assemble(delta_time) + :346-393 assemble_local_cell (implicit Euler)
void heat_assemble(Oracle& o, double delta_time, const double* face_bc) {
    const double cu_rho_cp = 3.4496e-24;                   // CurrentHeatSolver.h:118
    const double gamma = cu_rho_cp * (1.0 / delta_time);
    std::fill(o.val.begin(), o.val.end(), 0.0);
    std::fill(o.rhs.begin(), o.rhs.end(), 0.0);
    const int n_cells = (int) o.cells.size();
    for (int c = 0; c < n_cells; ++c) {
        int dofs[8];
        for (int i = 0; i < 8; ++i) dofs[i] = o.vertex2dof[o.cells[c][i]];
        double Ke[8][8] = {}, Fe[8] = {};
        double JxW[8]; V3 g[8][8]; double prev_T[8]; V3 pot_grad[8];
        for (int q = 0; q < 8; ++q) {
            cell_geometry(o, c, q, JxW[q], g[q]);
            prev_T[q] = 0; pot_grad[q] = V3();
            for (int i = 0; i < 8; ++i) {                  // get_function_values / get_function_gradients
                prev_T[q] += o.ch_heat[dofs[i]] * REFCUBE.N[q][i];
                pot_grad[q] = pot_grad[q] + g[q][i] * o.ch_current[dofs[i]];
            }
        }
        for (int q = 0; q < 8; ++q) {
            const double kappa = pq_kappa(o, prev_T[q]);
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j)
                    Ke[i][j] += JxW[q] * (gamma * REFCUBE.N[q][i] * REFCUBE.N[q][j] + kappa * dot(g[q][i], g[q][j]));
        }
        for (int q = 0; q < 8; ++q) {
            const double pot_grad_squared = dot(pot_grad[q], pot_grad[q]);
            const double sigma = pq_sigma(o, prev_T[q]);
            for (int i = 0; i < 8; ++i) Fe[i] += JxW[q] * REFCUBE.N[q][i] * (gamma * prev_T[q] + sigma * pot_grad_squared);
        }
        for (int i = 0; i < 8; ++i) {                      // DealSolver.cpp:64-72 copy_global_cell
            o.rhs[dofs[i]] += Fe[i];
            for (int j = 0; j < 8; ++j) o.val[csr_pos(o, dofs[i], dofs[j])] += Ke[i][j];
        }
    }
    assemble_rhs_faces(o, BID_COPPER, face_bc);
    o.boundary_values.clear();
    append_dirichlet(o, BID_BOTTOM, o.ch_T_ambient);
    o.sol = o.ch_heat;
    apply_dirichlet(o);
    o.ch_heat = o.sol;
}

// ============================================================================
//  Interpolation half  (pinned against the compiled reference)
// ============================================================================

// InterpolatorCells.cpp:523-629 LinearTetrahedra::precompute + :720-742 narrow_search_to(VACUUM)
void precompute_tets(Oracle& o) {
    const int nt = o.n_tet;
    o.t_det0.assign(nt, 0); for (int k = 0; k < 4; ++k) o.t_det[k].assign(nt, V4());
    o.t_cent.assign(nt, V3()); o.t_mark.assign(nt, 0); o.t_nbr.assign(nt, {});
    o.decay_factor = -1.0 / o.tet_edgemax;
    std::vector<std::vector<int>> node2tets(o.n_nodes);
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < 4; ++k) node2tets[o.tet4[4 * t + k]].push_back(t);
    for (int t = 0; t < nt; ++t) {
        const int* nn = &o.tet_nbr[4 * t];
        for (int k = 0; k < 4; ++k) if (nn[k] >= 0) o.t_nbr[t].push_back(nn[k]);
        for (int k = 0; k < 4; ++k)
            for (int nb : node2tets[o.tet4[4 * t + k]])
                if (nb != t && nb != nn[0] && nb != nn[1] && nb != nn[2] && nb != nn[3]) o.t_nbr[t].push_back(nb);
        // TetgenCells.h:142-151 get_centroid
        V3 c;
        for (int k = 0; k < 4; ++k) c = c + o.xyz[o.tet4[4 * t + k]];
        o.t_cent[t] = c * (1.0 / 4);
        const V3 v1 = o.xyz[o.tet4[4 * t]], v2 = o.xyz[o.tet4[4 * t + 1]], v3 = o.xyz[o.tet4[4 * t + 2]], v4 = o.xyz[o.tet4[4 * t + 3]];
        const double d0 = det4(v1, v2, v3, v4);
        o.t_det0[t] = 1.0 / d0;
        double d1, d2, d3, d4;
        d1 = det2({v2.y, v3.y, v4.y}, {v2.z, v3.z, v4.z});
        d2 = det2({v2.x, v3.x, v4.x}, {v2.z, v3.z, v4.z});
        d3 = det2({v2.x, v3.x, v4.x}, {v2.y, v3.y, v4.y});
        d4 = det3({v2.x, v3.x, v4.x}, {v2.y, v3.y, v4.y}, {v2.z, v3.z, v4.z});
        o.t_det[0][t] = {d1, -d2, d3, -d4};
        d1 = det2({v1.y, v3.y, v4.y}, {v1.z, v3.z, v4.z});
        d2 = det2({v1.x, v3.x, v4.x}, {v1.z, v3.z, v4.z});
        d3 = det2({v1.x, v3.x, v4.x}, {v1.y, v3.y, v4.y});
        d4 = det3({v1.x, v3.x, v4.x}, {v1.y, v3.y, v4.y}, {v1.z, v3.z, v4.z});
        o.t_det[1][t] = {-d1, d2, -d3, d4};
        d1 = det2({v1.y, v2.y, v4.y}, {v1.z, v2.z, v4.z});
        d2 = det2({v1.x, v2.x, v4.x}, {v1.z, v2.z, v4.z});
        d3 = det2({v1.x, v2.x, v4.x}, {v1.y, v2.y, v4.y});
        d4 = det3({v1.x, v2.x, v4.x}, {v1.y, v2.y, v4.y}, {v1.z, v2.z, v4.z});
        o.t_det[2][t] = {d1, -d2, d3, -d4};
        d1 = det2({v1.y, v2.y, v3.y}, {v1.z, v2.z, v3.z});
        d2 = det2({v1.x, v2.x, v3.x}, {v1.z, v2.z, v3.z});
        d3 = det2({v1.x, v2.x, v3.x}, {v1.y, v2.y, v3.y});
        d4 = det3(v1, v2, v3);
        o.t_det[3][t] = {-d1, d2, -d3, d4};
        o.t_mark[t] = o.tet_marker[t] != TYPE_VACUUM;    // narrow_search_to(VACUUM), :730-732
    }
}

// InterpolatorCells.cpp:631-649
bool tet_point_in_cell(const Oracle& o, V3 p, int i) {
    const V4 pt = {p.x, p.y, p.z, 1};
    if (o.t_det0[i] * dot4(pt, o.t_det[0][i]) < -ZERO) return false;
    if (o.t_det0[i] * dot4(pt, o.t_det[1][i]) < -ZERO) return false;
    if (o.t_det0[i] * dot4(pt, o.t_det[2][i]) < -ZERO) return false;
    if (o.t_det0[i] * dot4(pt, o.t_det[3][i]) < -ZERO) return false;
    return true;
}

// InterpolatorCells.cpp:651-661
void tet_shape_functions(const Oracle& o, V3 p, int t, double bcc[4]) {
    const V4 pt = {p.x, p.y, p.z, 1};
    for (int k = 0; k < 4; ++k) bcc[k] = ZERO + o.t_det0[t] * dot4(pt, o.t_det[k][t]);
}

// InterpolatorCells.cpp:269-307 InterpolatorCells<dim>::locate_cell, instantiated for tets
int tet_locate_cell(const Oracle& o, V3 p, int guess) {
    const int n_cells = o.n_tet;
    if (guess >= 0) {
        if (tet_point_in_cell(o, p, guess)) return guess;
        for (int c : o.t_nbr[guess]) if (tet_point_in_cell(o, p, c)) return c;
    }
    double min_d2 = 1e100; int min_index = 0;
    for (int c = 0; c < n_cells; ++c) {
        if (o.t_mark[c] == 0 && tet_point_in_cell(o, p, c)) return c;
        const double d2 = dist2(p, o.t_cent[c]);
        if (d2 < min_d2) { min_d2 = d2; min_index = c; }
    }
    return -min_index;
}

// InterpolatorCells.cpp:1205-1267 LinearHexahedra::precompute
void precompute_hexs(Oracle& o) {
    const int nh = o.n_hex;
    for (int k = 0; k < 8; ++k) o.h_f[k].assign(nh, V3());
    o.h_cent.assign(nh, V3());
    for (int h = 0; h < nh; ++h) {
        V3 c;
        for (int k = 0; k < 8; ++k) c = c + o.xyz[o.hex8[8 * h + k]];
        o.h_cent[h] = c * (1.0 / 8);
        V3 x[8];
        for (int k = 0; k < 8; ++k) x[k] = o.xyz[o.hex8[8 * h + k]];
        const V3 x1 = x[0], x2 = x[1], x3 = x[2], x4 = x[3], x5 = x[4], x6 = x[5], x7 = x[6], x8 = x[7];
        o.h_f[0][h] = (x1 + x2 + x3 + x4 + x5 + x6 + x7 + x8) / 8.0;
        o.h_f[1][h] = ((x1 * -1) + x2 + x3 - x4 - x5 + x6 + x7 - x8) / 8.0;
        o.h_f[2][h] = ((x1 * -1) - x2 + x3 + x4 - x5 - x6 + x7 + x8) / 8.0;
        o.h_f[3][h] = ((x1 * -1) - x2 - x3 - x4 + x5 + x6 + x7 + x8) / 8.0;
        o.h_f[4][h] = (x1 - x2 + x3 - x4 + x5 - x6 + x7 - x8) / 8.0;
        o.h_f[5][h] = (x1 - x2 - x3 + x4 - x5 + x6 + x7 - x8) / 8.0;
        o.h_f[6][h] = (x1 + x2 - x3 - x4 - x5 - x6 + x7 + x8) / 8.0;
        o.h_f[7][h] = ((x1 * -1) + x2 - x3 + x4 + x5 - x6 + x7 - x8) / 8.0;
    }
}

// InterpolatorCells.cpp:1507-1528
bool hex_point_in_cell(const Oracle& o, V3 p, int cell) {
    double b[4];
    tet_shape_functions(o, p, cell / 4, b);
    if (b[0] >= 0 && b[1] >= 0 && b[2] >= 0 && b[3] >= 0) {
        switch (cell % 4) {
            case 0: return b[0] >= b[1] && b[0] >= b[2] && b[0] >= b[3];
            case 1: return b[1] >= b[0] && b[1] >= b[2] && b[1] >= b[3];
            case 2: return b[2] >= b[0] && b[2] >= b[1] && b[2] >= b[3];
            case 3: return b[3] >= b[0] && b[3] >= b[1] && b[3] >= b[2];
        }
    }
    return false;
}

// InterpolatorCells.cpp:1536-1562
int hex_locate_cell(const Oracle& o, V3 p, int guess) {
    int tet = guess / 4;
    tet = tet_locate_cell(o, p, tet);
    int sign = 1;
    if (tet < 0) sign = -1;
    tet = std::abs(tet);
    double b[4];
    tet_shape_functions(o, p, tet, b);
    if (b[0] >= b[1] && b[0] >= b[2] && b[0] >= b[3]) return sign * (4 * tet + 0);
    if (b[1] >= b[0] && b[1] >= b[2] && b[1] >= b[3]) return sign * (4 * tet + 1);
    if (b[2] >= b[0] && b[2] >= b[1] && b[2] >= b[3]) return sign * (4 * tet + 2);
    if (b[3] >= b[0] && b[3] >= b[1] && b[3] >= b[2]) return sign * (4 * tet + 3);
    return -1;
}

// InterpolatorCells.cpp:1363-1416 shape_fun_grads(point, hex)
void hex_shape_fun_grads(const Oracle& o, V3 point, int hex, V3 sfg[8]) {
    double u, v, w;
    hex_nat_coords(o, point, hex, u, v, w);
    V3 xyz[8];
    for (int i = 0; i < 8; ++i) xyz[i] = o.xyz[o.hex8[8 * hex + i]];
    V3 dN[8] = {
        {-(1 - v) * (1 - w), -(1 - u) * (1 - w), -(1 - u) * (1 - v)},
        { (1 - v) * (1 - w), -(1 + u) * (1 - w), -(1 + u) * (1 - v)},
        { (1 + v) * (1 - w),  (1 + u) * (1 - w), -(1 + u) * (1 + v)},
        {-(1 + v) * (1 - w),  (1 - u) * (1 - w), -(1 - u) * (1 + v)},
        {-(1 - v) * (1 + w), -(1 - u) * (1 + w),  (1 - u) * (1 - v)},
        { (1 - v) * (1 + w), -(1 + u) * (1 + w),  (1 + u) * (1 - v)},
        { (1 + v) * (1 + w),  (1 + u) * (1 + w),  (1 + u) * (1 + v)},
        {-(1 + v) * (1 + w),  (1 - u) * (1 + w),  (1 - u) * (1 + v)}};
    for (int i = 0; i < 8; ++i) dN[i] = dN[i] * 0.125;
    V3 J[3];
    for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 3; ++i) J[i] = J[i] + xyz[k] * comp(dN[k], i);
    double Jdet = det3(J[0], J[1], J[2]);
    Jdet = 1.0 / Jdet;
    auto Jc = [&](int a, int b) { return comp(J[a], b); };
    V3 Jinv[3] = {
        {Jc(1,1)*Jc(2,2)-Jc(2,1)*Jc(1,2), Jc(2,0)*Jc(1,2)-Jc(1,0)*Jc(2,2), Jc(1,0)*Jc(2,1)-Jc(2,0)*Jc(1,1)},
        {Jc(2,1)*Jc(0,2)-Jc(0,1)*Jc(2,2), Jc(0,0)*Jc(2,2)-Jc(2,0)*Jc(0,2), Jc(2,0)*Jc(0,1)-Jc(0,0)*Jc(2,1)},
        {Jc(0,1)*Jc(1,2)-Jc(1,1)*Jc(0,2), Jc(1,0)*Jc(0,2)-Jc(0,0)*Jc(1,2), Jc(0,0)*Jc(1,1)-Jc(1,0)*Jc(0,1)}};
    for (int i = 0; i < 3; ++i) Jinv[i] = Jinv[i] * Jdet;
    for (int k = 0; k < 8; ++k) {
        sfg[k] = V3();
        for (int i = 0; i < 3; ++i) sfg[k] = sfg[k] + Jinv[i] * comp(dN[k], i);
    }
}

// InterpolatorCells.cpp:1319-1332 + :1462-1505 shape_fun_grads(hex, node), and
// InterpolatorCells.cpp:425-439 interp_gradient(cell, node)
V3 hex_nodal_gradient(const Oracle& o, int hex, int node) {
    static const double UVW[8][3] = {{-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};
    static const int NNN[8][3] = {{1,3,4},{0,2,5},{3,1,6},{2,0,7},{5,7,0},{4,6,1},{7,5,2},{6,4,3}};
    const double* uvw = UVW[node];
    const int* nnn = NNN[node];
    const int* shex = &o.hex8[8 * hex];
    const V3 vec0 = o.xyz[shex[node]];
    V3 J[3] = {(vec0 - o.xyz[shex[nnn[0]]]) * (uvw[0] * 0.5),
               (vec0 - o.xyz[shex[nnn[1]]]) * (uvw[1] * 0.5),
               (vec0 - o.xyz[shex[nnn[2]]]) * (uvw[2] * 0.5)};
    double Jdet = det3(J[0], J[1], J[2]);
    Jdet = 1.0 / Jdet;
    auto Jc = [&](int a, int b) { return comp(J[a], b); };
    V3 Jinv[3] = {
        {Jc(1,1)*Jc(2,2)-Jc(2,1)*Jc(1,2), Jc(2,1)*Jc(0,2)-Jc(0,1)*Jc(2,2), Jc(0,1)*Jc(1,2)-Jc(1,1)*Jc(0,2)},
        {Jc(2,0)*Jc(1,2)-Jc(1,0)*Jc(2,2), Jc(0,0)*Jc(2,2)-Jc(2,0)*Jc(0,2), Jc(1,0)*Jc(0,2)-Jc(0,0)*Jc(1,2)},
        {Jc(1,0)*Jc(2,1)-Jc(2,0)*Jc(1,1), Jc(2,0)*Jc(0,1)-Jc(0,0)*Jc(2,1), Jc(0,0)*Jc(1,1)-Jc(1,0)*Jc(0,1)}};
    for (int i = 0; i < 3; ++i) Jinv[i] = Jinv[i] * Jdet;
    V3 sfg[8];
    const V3 uvwv = {uvw[0], uvw[1], uvw[2]};
    for (int i = 0; i < 3; ++i) {
        compref(sfg[node], i) = 0.5 * dot(Jinv[i], uvwv);
        for (int j = 0; j < 3; ++j) compref(sfg[nnn[j]], i) = -0.5 * comp(Jinv[i], j) * uvw[j];
    }
    V3 r;
    for (int i = 0; i < 8; ++i) r = r - sfg[i] * o.nodal[shex[i]].s2;
    return r;
}

// InterpolatorCells.cpp:410-423 interp_gradient(point, cell) for hexes
V3 hex_interp_gradient(const Oracle& o, V3 point, int hex) {
    V3 sfg[8];
    hex_shape_fun_grads(o, point, hex, sfg);
    V3 r;
    for (int i = 0; i < 8; ++i) r = r - sfg[i] * o.nodal[o.hex8[8 * hex + i]].s2;
    return r;
}

// InterpolatorCells.cpp:1585-1637 LinearTriangles::precompute
void precompute_tris(Oracle& o) {
    const int n = o.n_tri;
    o.r_vert0.assign(n, V3()); o.r_edge1 = o.r_edge2 = o.r_pvec = o.r_norm = o.r_cent = o.r_vert0;
    o.r_maxd.assign(n, 0); o.r_nbr.assign(n, {});
    std::vector<std::vector<int>> node2tris(o.n_nodes);
    for (int t = 0; t < n; ++t) for (int k = 0; k < 3; ++k) node2tris[o.tri3[3 * t + k]].push_back(t);
    for (int t = 0; t < n; ++t) {
        for (int k = 0; k < 3; ++k)
            for (int nb : node2tris[o.tri3[3 * t + k]]) if (nb != t) o.r_nbr[t].push_back(nb);
        const V3 v0 = o.xyz[o.tri3[3 * t]], v1 = o.xyz[o.tri3[3 * t + 1]], v2 = o.xyz[o.tri3[3 * t + 2]];
        const V3 e1 = v1 - v0, e2 = v2 - v0;
        const V3 pv = cross(o.tri_norm[t], e2);
        const double i_det = 1.0 / dot(e1, pv);
        o.r_vert0[t] = v0; o.r_edge1[t] = e1 * i_det; o.r_edge2[t] = e2; o.r_pvec[t] = pv * i_det;
        o.r_norm[t] = o.tri_norm[t];
        o.r_maxd[t] = std::sqrt(e2.x * e2.x + e2.y * e2.y + e2.z * e2.z);
        o.r_cent[t] = (v0 + v1 + v2) / 3.0;      // TetgenCells.cpp:402
    }
}

// InterpolatorCells.cpp:1639-1650
bool tri_point_in_cell(const Oracle& o, V3 p, int f) {
    const V3 tvec = p - o.r_vert0[f];
    const double u = dot(tvec, o.r_pvec[f]);
    if (u < -ZERO || u > 1 + ZERO) return false;
    const V3 qvec = cross(tvec, o.r_edge1[f]);
    const double v = dot(qvec, o.r_norm[f]);
    if (v < -ZERO || u + v > 1 + ZERO) return false;
    return std::fabs(dot(qvec, o.r_edge2[f])) < o.r_maxd[f];
}

// InterpolatorCells.cpp:1652-1660
void tri_shape_functions(const Oracle& o, V3 p, int f, double bcc[3]) {
    const V3 tvec = p - o.r_vert0[f];
    const V3 qvec = cross(tvec, o.r_edge1[f]);
    const double v = dot(tvec, o.r_pvec[f]);
    const double w = dot(qvec, o.r_norm[f]);
    const double u = 1.0 - v - w;
    bcc[0] = ZERO + u; bcc[1] = ZERO + v; bcc[2] = ZERO + w;
}

// InterpolatorCells.cpp:1743-1747
double tri_fast_distance(const Oracle& o, V3 p, int f) {
    const V3 tvec = p - o.r_vert0[f];
    const V3 qvec = cross(tvec, o.r_edge1[f]);
    return dot(o.r_edge2[f], qvec);
}

// InterpolatorCells.cpp:269-307 instantiated for triangles (markers are all 0 for lintri)
int tri_locate_cell(const Oracle& o, V3 p, int guess) {
    if (guess >= 0) {
        if (tri_point_in_cell(o, p, guess)) return guess;
        for (int c : o.r_nbr[guess]) if (tri_point_in_cell(o, p, c)) return c;
    }
    double min_d2 = 1e100; int min_index = 0;
    for (int c = 0; c < o.n_tri; ++c) {
        if (tri_point_in_cell(o, p, c)) return c;
        const double d2 = dist2(p, o.r_cent[c]);
        if (d2 < min_d2) { min_d2 = d2; min_index = c; }
    }
    return -min_index;
}

// InterpolatorCells.cpp:1922-1946 LinearQuadrangles::point_in_cell
bool quad_point_in_cell(const Oracle& o, V3 p, int cell) {
    double b[3];
    tri_shape_functions(o, p, cell / 3, b);
    if (b[0] >= 0 && b[1] >= 0 && b[2] >= 0) {
        switch (cell % 3) {
            case 0: return b[0] >= b[1] && b[0] >= b[2];
            case 1: return b[1] >= b[0] && b[1] >= b[2];
            case 2: return b[2] >= b[0] && b[2] >= b[1];
        }
    }
    return false;
}

// InterpolatorCells.cpp:1954-1982 LinearQuadrangles::locate_cell
int quad_locate_cell(const Oracle& o, V3 p, int guess) {
    int tri = guess / 3;
    tri = tri_locate_cell(o, p, tri);
    int sign = 1;
    if (tri < 0) sign = -1;
    tri = std::abs(tri);
    double b[3];
    tri_shape_functions(o, p, tri, b);
    if (b[0] >= b[1] && b[0] >= b[2]) return sign * (3 * tri + 0);
    if (b[1] >= b[0] && b[1] >= b[2]) return sign * (3 * tri + 1);
    if (b[2] >= b[0] && b[2] >= b[1]) return sign * (3 * tri + 2);
    return -1;
}

// InterpolatorCells.cpp:1151-1173 QuadraticTetrahedra::calc_cell, :1873-1895 QuadraticTriangles::calc_cell
void precompute_quadratic(Oracle& o) {
    auto common = [](const std::vector<int>& a, const std::vector<int>& b) {   // :447-452 common_entry
        for (int i : a) for (int j : b) if (i == j) return i;
        return -1;
    };
    o.qtet.assign(o.n_tet, {});
    for (int t = 0; t < o.n_tet; ++t) {
        std::array<int, 10> c{};
        if (o.n_hex > t) {
            std::vector<int> en[4];
            for (int i = 0; i < 4; ++i)
                for (int k = 0; k < 8; ++k) {
                    const int hn = o.hex8[8 * (4 * t + i) + k];
                    if (o.node_marker[hn] == NODE_EDGE) en[i].push_back(hn);
                }
            for (int k = 0; k < 4; ++k) c[k] = o.tet4[4 * t + k];
            c[4] = common(en[0], en[1]); c[5] = common(en[1], en[2]); c[6] = common(en[2], en[0]);
            c[7] = common(en[0], en[3]); c[8] = common(en[1], en[3]); c[9] = common(en[2], en[3]);
        }
        o.qtet[t] = c;
    }
    o.qtri.assign(o.n_tri, {});
    for (int f = 0; f < o.n_tri; ++f) {
        std::array<int, 6> c{};
        if (o.n_quad > 0) {
            std::vector<int> en[3];
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 4; ++k) {
                    const int qn = o.quad4[4 * (3 * f + i) + k];
                    if (o.node_marker[qn] == NODE_EDGE) en[i].push_back(qn);
                }
            for (int k = 0; k < 3; ++k) c[k] = o.tri3[3 * f + k];
            c[3] = common(en[0], en[1]); c[4] = common(en[1], en[2]); c[5] = common(en[2], en[0]);
        }
        o.qtri[f] = c;
    }
}

template <int N>
Sol weighted(const Oracle& o, const int* nodes, const double* w) {   // InterpolatorCells.cpp:359-380
    Sol r;
    for (int i = 0; i < N; ++i) {
        const Sol& s = o.nodal[nodes[i]];
        r.v = r.v + s.v * w[i];
        r.s1 += s.s1 * w[i];
        r.s2 += s.s2 * w[i];
    }
    return r;
}

Sol tet_interp(const Oracle& o, V3 p, int c) {
    const int cell = std::abs(c);
    double w[4];
    tet_shape_functions(o, p, cell, w);
    return weighted<4>(o, &o.tet4[4 * cell], w);
}

// InterpolatorCells.cpp:795-816 QuadraticTetrahedra::shape_functions
Sol qtet_interp(const Oracle& o, V3 p, int c) {
    const int cell = std::abs(c);
    double b[4];
    tet_shape_functions(o, p, cell, b);
    const double b1 = b[0], b2 = b[1], b3 = b[2], b4 = b[3];
    const double w[10] = {b1 * (2 * b1 - 1), b2 * (2 * b2 - 1), b3 * (2 * b3 - 1), b4 * (2 * b4 - 1),
                          4 * b1 * b2, 4 * b2 * b3, 4 * b3 * b1, 4 * b1 * b4, 4 * b2 * b4, 4 * b3 * b4};
    return weighted<10>(o, o.qtet[cell].data(), w);
}

Sol hex_interp(const Oracle& o, V3 p, int c) {
    const int cell = std::abs(c);
    double w[8];
    hex_shape_functions(o, p, cell, w);
    return weighted<8>(o, &o.hex8[8 * cell], w);
}

// InterpolatorCells.cpp:1662-1686 LinearTriangles::interp_solution
Sol tri_interp(const Oracle& o, V3 p, int t) {
    const int tri = std::abs(t);
    const int t0 = o.tri2tet[2 * tri], t1 = o.tri2tet[2 * tri + 1];
    const double d = std::fabs(tri_fast_distance(o, p, tri));
    if (d <= 100.0 * ZERO) return tet_interp(o, p, t0);
    if (tet_point_in_cell(o, p, t0)) return tet_interp(o, p, t0);
    const int tet = tet_locate_cell(o, p, t1);
    return tet_interp(o, p, tet);
}

// InterpolatorCells.cpp:1838-1861 QuadraticTriangles::interp_solution
Sol qtri_interp(const Oracle& o, V3 p, int t) {
    const int tri = std::abs(t);
    const int t0 = o.tri2tet[2 * tri], t1 = o.tri2tet[2 * tri + 1];
    const double d = std::fabs(tri_fast_distance(o, p, tri));
    if (d <= 100.0 * ZERO) return qtet_interp(o, p, t0);
    if (tet_point_in_cell(o, p, t0)) return qtet_interp(o, p, t0);
    const int tet = tet_locate_cell(o, p, t1);
    return qtet_interp(o, p, tet);
}

// InterpolatorCells.cpp:1898-1920 LinearQuadrangles::interp_solution
Sol quad_interp(const Oracle& o, V3 p, int q) {
    const int quad = std::abs(q);
    const int h0 = o.quad2hex[2 * quad], h1 = o.quad2hex[2 * quad + 1];
    const double d = std::fabs(tri_fast_distance(o, p, quad / 3));
    if (d <= 100.0 * ZERO) return hex_interp(o, p, h0);
    if (hex_point_in_cell(o, p, h0)) return hex_interp(o, p, h0);
    const int hex = hex_locate_cell(o, p, h1);
    return hex_interp(o, p, hex);
}

int locate_any(const Oracle& o, int dim, int rank, V3 p, int guess) {
    if (dim == 2) return rank == 3 ? quad_locate_cell(o, p, guess) : tri_locate_cell(o, p, guess);
    return rank == 3 ? hex_locate_cell(o, p, guess) : tet_locate_cell(o, p, guess);
}

Sol interp_any(const Oracle& o, int dim, int rank, V3 p, int cell) {
    if (dim == 2) return rank == 1 ? tri_interp(o, p, cell) : (rank == 2 ? qtri_interp(o, p, cell) : quad_interp(o, p, cell));
    return rank == 1 ? tet_interp(o, p, cell) : (rank == 2 ? qtet_interp(o, p, cell) : hex_interp(o, p, cell));
}

}  // namespace

// ============================================================================
//  C API (ctypes)
// ============================================================================
extern "C" {

// the OpenMP team of vmult (bench.py sets it explicitly: launchers such as torchrun export OMP_NUM_THREADS=1)
void fo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int fo_get_max_threads() { return omp_get_max_threads(); }

void* fo_create() { return new Oracle(); }
void fo_destroy(void* h) { delete (Oracle*) h; }

// 1 (the reference build) or 2; read by the next fo_import_mesh
void fo_set_fe_degree(void* h, int degree) { ((Oracle*) h)->fe_degree = degree == 2 ? 2 : 1; }
void fo_get_cell_dofs27(void* h, int* out) {
    Oracle& o = *(Oracle*) h;
    for (size_t c = 0; c < o.cdofs.size(); ++c) std::copy(o.cdofs[c].begin(), o.cdofs[c].end(), out + 27 * c);
}
int fo_import_mesh(void* h, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    return import_mesh(*(Oracle*) h, xyz, n_nodes, hex8, hex_marker, n_hex);
}
void fo_setup(void* h, double field, double potential, int anode_dirichlet) { setup(*(Oracle*) h, field, potential, anode_dirichlet); }
void fo_assemble(void* h, int first_time, const double* pxyz, const int* pcell, long n, double charge_factor) {
    assemble(*(Oracle*) h, first_time, pxyz, pcell, n, charge_factor);
}
int fo_solve(void* h, int max_iter, double tol, double ssor, int precond, double* res) {
    Oracle& o = *(Oracle*) h;
    const int it = solve_cg(o, max_iter, tol, ssor, precond);
    if (res) *res = o.last_res;
    return it;
}
// ---- CurrentHeatSolver (bulk mesh) ----
int fo_import_mesh_kind(void* h, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex, int kind) {
    return import_mesh(*(Oracle*) h, xyz, n_nodes, hex8, hex_marker, n_hex, kind);
}
void fo_ch_set_physics(void* h, const double* T, const double* rho, int n, double lorentz) {
    Oracle& o = *(Oracle*) h; o.res_T.assign(T, T + n); o.res_rho.assign(rho, rho + n); o.lorentz = lorentz;
}
double fo_ch_sigma(void* h, double T) { return pq_sigma(*(Oracle*) h, T); }
double fo_ch_kappa(void* h, double T) { return pq_kappa(*(Oracle*) h, T); }
void fo_ch_setup(void* h, double T_ambient) { ch_setup(*(Oracle*) h, T_ambient); }
void fo_current_assemble(void* h, const double* face_bc) { current_assemble(*(Oracle*) h, face_bc); }
void fo_heat_assemble(void* h, double delta_time, const double* face_bc) { heat_assemble(*(Oracle*) h, delta_time, face_bc); }
// EmissionSolver::solve (CurrentHeatSolver.h:41): solve_cg(conf->n_cg, conf->cg_tolerance, conf->ssor_param) on the system assembled last
int fo_ch_solve(void* h, int which, int max_iter, double tol, double ssor, int precond, double* res) {
    Oracle& o = *(Oracle*) h;
    o.sol = which ? o.ch_heat : o.ch_current;
    const int it = solve_cg(o, max_iter, tol, ssor, precond);
    (which ? o.ch_heat : o.ch_current) = o.sol;
    if (res) *res = o.last_res;
    return it;
}
// which: 0 = current potential, 1 = temperature; dof order
void fo_ch_get_solution(void* h, int which, double* out) {
    Oracle& o = *(Oracle*) h; const auto& v = which ? o.ch_heat : o.ch_current; std::copy(v.begin(), v.end(), out);
}
void fo_ch_set_solution(void* h, int which, const double* in) {
    Oracle& o = *(Oracle*) h; auto& v = which ? o.ch_heat : o.ch_current; v.assign(in, in + o.n_dofs);
}
// makes `which` the solution that fo_export_solution / fo_export_solution_grad / fo_check_limits read
void fo_ch_select(void* h, int which) { Oracle& o = *(Oracle*) h; o.sol = which ? o.ch_heat : o.ch_current; }
// DealSolver.cpp:229-245 export_surface_centroids: centres of the copper_surface faces, cell/face iteration order
int fo_surface_centroids(void* h, double* xyz3) {
    Oracle& o = *(Oracle*) h; int n = 0;
    for (const auto& bf : o.bfaces) {
        if (bf.id != BID_COPPER) continue;
        if (xyz3) {
            V3 s;
            for (int v = 0; v < 4; ++v) s = s + cell_vertex(o, bf.cell, FACE_VERTS[bf.face][v]);
            s = s / 4.0;
            xyz3[3 * n] = s.x; xyz3[3 * n + 1] = s.y; xyz3[3 * n + 2] = s.z;
        }
        ++n;
    }
    return n;
}

// out: n_dofs, n_cells, nnz, n_vertices, n_boundary_faces
void fo_sizes(void* h, long* out) {
    Oracle& o = *(Oracle*) h;
    out[0] = o.n_dofs; out[1] = (long) o.cells.size(); out[2] = (long) o.col.size();
    out[3] = (long) o.vert2node.size(); out[4] = (long) o.bfaces.size();
}
void fo_get_csr(void* h, int* rowptr, int* col, double* val, double* val_save) {
    Oracle& o = *(Oracle*) h;
    if (rowptr) std::copy(o.rowptr.begin(), o.rowptr.end(), rowptr);
    if (col) std::copy(o.col.begin(), o.col.end(), col);
    if (val) std::copy(o.val.begin(), o.val.end(), val);
    if (val_save) std::copy(o.val_save.begin(), o.val_save.end(), val_save);
}
void fo_get_vectors(void* h, double* rhs, double* sol, int* vertex2dof, int* vert2node) {
    Oracle& o = *(Oracle*) h;
    if (rhs) std::copy(o.rhs.begin(), o.rhs.end(), rhs);
    if (sol) std::copy(o.sol.begin(), o.sol.end(), sol);
    if (vertex2dof) std::copy(o.vertex2dof.begin(), o.vertex2dof.end(), vertex2dof);
    if (vert2node) std::copy(o.vert2node.begin(), o.vert2node.end(), vert2node);
}
void fo_set_solution(void* h, const double* sol_dof) { Oracle& o = *(Oracle*) h; std::copy(sol_dof, sol_dof + o.n_dofs, o.sol.begin()); }
void fo_get_bfaces(void* h, int* cell, int* face, int* id) {
    Oracle& o = *(Oracle*) h;
    for (size_t i = 0; i < o.bfaces.size(); ++i) { cell[i] = o.bfaces[i].cell; face[i] = o.bfaces[i].face; id[i] = o.bfaces[i].id; }
}
void fo_get_cells(void* h, int* cells8) {
    Oracle& o = *(Oracle*) h;
    for (size_t c = 0; c < o.cells.size(); ++c) for (int k = 0; k < 8; ++k) cells8[8 * c + k] = o.cells[c][k];
}
// DealSolver.cpp:269-278 export_solution (vertex order)
void fo_export_solution(void* h, double* phi_vertex) {
    Oracle& o = *(Oracle*) h;
    for (size_t v = 0; v < o.vertex2dof.size(); ++v) phi_vertex[v] = o.sol[o.vertex2dof[v]];
}
// PoissonSolver.cpp:141-149 export_charge_dens: charge_density is reinit'ed to zero unless a file is
// being written (PoissonSolver.cpp:198-207); file output is outside the hot path -> zeros
// DealSolver.cpp:280-301 export_solution_grad with :317-341 calc_vertex2dof: vertex2cell / vertex2node are overwritten cell
// by cell (the last cell holding the vertex wins); the gradient of the solution is evaluated at the QGauss<3>(2) points
// of that cell and the one with index vertex2node[v] is taken (the reference indexes quadrature points by local vertex)
void fo_export_solution_grad(void* h, double* grad3_vertex) {
    Oracle& o = *(Oracle*) h;
    const int n_vert = (int) o.vert2node.size();
    std::vector<int> v2cell(n_vert, 0), v2node(n_vert, 0);
    for (int c = 0; c < (int) o.cells.size(); ++c)
        for (int i = 0; i < 8; ++i) { v2cell[o.cells[c][i]] = c; v2node[o.cells[c][i]] = i; }
    for (int v = 0; v < n_vert; ++v) {
        double JxW; V3 g[8];
        cell_geometry(o, v2cell[v], v2node[v], JxW, g);
        V3 s;
        for (int k = 0; k < 8; ++k) s = s + g[k] * o.sol[o.vertex2dof[o.cells[v2cell[v]][k]]];
        grad3_vertex[3 * v] = -1.0 * s.x; grad3_vertex[3 * v + 1] = -1.0 * s.y; grad3_vertex[3 * v + 2] = -1.0 * s.z;
    }
}
// operator<<(ostream&, const DealSolver&) counts (DealSolver.h:107-117): #faces, #edges of the solver mesh
void fo_mesh_counts(void* h, long* n_faces, long* n_edges) {
    Oracle& o = *(Oracle*) h;
    static const int E[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    std::set<std::pair<int, int>> edges;
    std::set<std::array<int, 4>> faces;
    for (const auto& c : o.cells) {
        for (const auto& e : E) edges.insert({std::min(c[e[0]], c[e[1]]), std::max(c[e[0]], c[e[1]])});
        for (int f = 0; f < 6; ++f) {
            std::array<int, 4> k;
            for (int v = 0; v < 4; ++v) k[v] = c[FACE_VERTS[f][v]];
            std::sort(k.begin(), k.end());
            faces.insert(k);
        }
    }
    *n_faces = (long) faces.size(); *n_edges = (long) edges.size();
}
void fo_export_charge_dens(void* h, double* rho_vertex) {
    Oracle& o = *(Oracle*) h;
    for (size_t v = 0; v < o.vertex2dof.size(); ++v)
        rho_vertex[v] = o.charge_density.size() == (size_t) o.n_dofs ? o.charge_density[o.vertex2dof[v]] : 0.0;
}
void fo_set_write_time(void* h, int on) { ((Oracle*) h)->write_time = on; }
// DealSolver.cpp:157-167 check_limits
int fo_check_limits(void* h, double lo, double hi, double* mn, double* mx) {
    Oracle& o = *(Oracle*) h;
    double a = 1e100, b = -1e100;
    for (double s : o.sol) { a = std::min(a, s); b = std::max(b, s); }
    *mn = a; *mx = b;
    return a < lo || b > hi;
}
// DealSolver.cpp:169-173 get_cell_vol = cell->measure(); for hexes equal to sum of JxW over QGauss(2)
// only up to quadrature exactness, so deal.II's closed-form measure is restated via 2x2x2 Gauss of detJ
// (exact for trilinear maps: detJ is a polynomial of degree <= 2 per variable).
double fo_cell_vol(void* h, int c) {
    Oracle& o = *(Oracle*) h;
    double vol = 0, JxW; V3 g[8];
    for (int q = 0; q < 8; ++q) { cell_geometry(o, c, q, JxW, g); vol += JxW; }
    return vol;
}

// Interpolator.cpp:28-77 initialize(mesh, 0, VACUUM) and the precompute() calls it makes
void fo_interp_initialize(void* h, const int* node_marker, const int* tet4, const int* tet_nbr, const int* tet_marker, int n_tet,
                          const int* tri3, const int* tri2tet, const double* tri_norm, int n_tri,
                          const int* quad4, const int* quad2hex, int n_quad, double tet_edgemax,
                          const int* voro_off, const int* voro_list, int n_voro) {
    Oracle& o = *(Oracle*) h;
    o.node_marker.assign(node_marker, node_marker + o.n_nodes);
    o.n_tet = n_tet; o.n_tri = n_tri; o.n_quad = n_quad;
    o.tet4.assign(tet4, tet4 + 4 * (size_t) n_tet); o.tet_nbr.assign(tet_nbr, tet_nbr + 4 * (size_t) n_tet);
    o.tet_marker.assign(tet_marker, tet_marker + n_tet);
    o.tri3.assign(tri3, tri3 + 3 * (size_t) n_tri); o.tri2tet.assign(tri2tet, tri2tet + 2 * (size_t) n_tri);
    o.tri_norm.resize(n_tri);
    for (int i = 0; i < n_tri; ++i) o.tri_norm[i] = {tri_norm[3 * i], tri_norm[3 * i + 1], tri_norm[3 * i + 2]};
    o.quad4.assign(quad4, quad4 + 4 * (size_t) n_quad); o.quad2hex.assign(quad2hex, quad2hex + 2 * (size_t) n_quad);
    o.tet_edgemax = tet_edgemax;
    o.voro_off.assign(voro_off, voro_off + n_voro + 1);
    o.voro_list.assign(voro_list, voro_list + voro_off[n_voro]);
    precompute_tets(o);
    precompute_hexs(o);
    precompute_tris(o);
    precompute_quadratic(o);
    o.q_cent.assign(n_quad, V3());
    o.nodal.assign(o.n_nodes, Sol());              // empty_value = 0
    o.node2cells.assign(o.n_nodes, {});
    for (int hx = 0; hx < o.n_hex; ++hx)
        if (o.hex_marker[hx] > 0)
            for (int k = 0; k < 8; ++k) o.node2cells[o.hex8[8 * hx + k]].push_back({hx, k});
}

// Interpolator.cpp:172-190 extract_solution(fem, smoothen): store_solution (:103-123),
// store_elfield (:125-140), average_nodal_fields (:142-170)
// the precomputed cell tables, flattened like the product's device records (tests compare them bit for bit):
// tet17 = {det0, det1[4], det2[4], det3[4], det4[4]}, tri16 = {vert0, edge1, edge2, pvec, norm, max_distance}, hex24 = f0..f7
void fo_get_tables(void* h, double* tet17, double* tet_cent3, int* tet_mark, double* hex24, double* tri16, double* tri_cent3,
                   int* qtet10, int* qtri6) {
    const Oracle& o = *(Oracle*) h;
    for (int t = 0; t < o.n_tet; ++t) {
        double* r = tet17 + 17 * (size_t) t;
        r[0] = o.t_det0[t];
        for (int k = 0; k < 4; ++k) { const V4& d = o.t_det[k][t]; r[1 + 4 * k] = d.x; r[2 + 4 * k] = d.y; r[3 + 4 * k] = d.z; r[4 + 4 * k] = d.w; }
        tet_cent3[3 * t] = o.t_cent[t].x; tet_cent3[3 * t + 1] = o.t_cent[t].y; tet_cent3[3 * t + 2] = o.t_cent[t].z;
        tet_mark[t] = o.t_mark[t];
        for (int k = 0; k < 10; ++k) qtet10[10 * (size_t) t + k] = o.qtet[t][k];
    }
    for (int c = 0; c < o.n_hex; ++c)
        for (int k = 0; k < 8; ++k) { const V3& f = o.h_f[k][c]; double* r = hex24 + 24 * (size_t) c + 3 * k; r[0] = f.x; r[1] = f.y; r[2] = f.z; }
    for (int t = 0; t < o.n_tri; ++t) {
        double* r = tri16 + 16 * (size_t) t;
        const V3* v[5] = {&o.r_vert0[t], &o.r_edge1[t], &o.r_edge2[t], &o.r_pvec[t], &o.r_norm[t]};
        for (int k = 0; k < 5; ++k) { r[3 * k] = v[k]->x; r[3 * k + 1] = v[k]->y; r[3 * k + 2] = v[k]->z; }
        r[15] = o.r_maxd[t];
        tri_cent3[3 * t] = o.r_cent[t].x; tri_cent3[3 * t + 1] = o.r_cent[t].y; tri_cent3[3 * t + 2] = o.r_cent[t].z;
        for (int k = 0; k < 6; ++k) qtri6[6 * (size_t) t + k] = o.qtri[t][k];
    }
}

void fo_extract_solution(void* h, int smoothen) {
    Oracle& o = *(Oracle*) h;
    for (int i = 0; i < o.n_nodes; ++i) {
        Sol s;
        if (o.node2vert[i] >= 0) {      // store_solution(charge_dens, potential): Interpolator.cpp:175-179
            const int d = o.vertex2dof[o.node2vert[i]];
            s.s1 = o.charge_density.size() == (size_t) o.n_dofs ? o.charge_density[d] : 0.0; s.s2 = o.sol[d];
        }
        o.nodal[i] = s;
    }
    for (int node = 0; node < o.n_nodes; ++node) {
        if (o.node2vert[node] < 0) continue;
        V3 mean;
        const int n_fields = (int) o.node2cells[node].size();
        if (n_fields > 0) {
            for (auto& p : o.node2cells[node]) mean = mean + hex_nodal_gradient(o, p.first, p.second);
            mean = mean * (1.0 / n_fields);
        }
        o.nodal[node].v = mean;
    }
    if (smoothen) {
        const int n_voro = (int) o.voro_off.size() - 1;
        for (int i = 0; i < n_voro; ++i) {
            if (o.voro_off[i + 1] == o.voro_off[i]) continue;
            const V3 tetnode = o.xyz[i];
            V3 vec; double w_sum = 0;
            for (int k = o.voro_off[i]; k < o.voro_off[i + 1]; ++k) {
                const int nb = o.voro_list[k];
                const double w = std::exp(o.decay_factor * std::sqrt(dist2(tetnode, o.xyz[nb])));
                w_sum += w;
                vec = vec + o.nodal[nb].v * w;
            }
            if (w_sum > 0) { vec = vec * (1.0 / w_sum); o.nodal[i].v = vec; }
        }
    }
}

void fo_set_nodal(void* h, const double* sol5) {
    Oracle& o = *(Oracle*) h;
    o.nodal.resize(o.n_nodes);
    for (int i = 0; i < o.n_nodes; ++i) o.nodal[i] = {{sol5[5 * i], sol5[5 * i + 1], sol5[5 * i + 2]}, sol5[5 * i + 3], sol5[5 * i + 4]};
}
void fo_get_nodal(void* h, double* sol5) {
    Oracle& o = *(Oracle*) h;
    for (int i = 0; i < o.n_nodes; ++i) {
        sol5[5 * i] = o.nodal[i].v.x; sol5[5 * i + 1] = o.nodal[i].v.y; sol5[5 * i + 2] = o.nodal[i].v.z;
        sol5[5 * i + 3] = o.nodal[i].s1; sol5[5 * i + 4] = o.nodal[i].s2;
    }
}

// SolutionReader.cpp:136-165 calc_full_interpolation + :43-65 locate_interpolate +
// InterpolatorCells.cpp:442-445 (cell = locate_cell(point, abs(cell)); chained guess starting at -1)
void fo_locate_interpolate(void* h, int dim, int rank, long n, const double* xyz, int* cells, double* sol5) {
    Oracle& o = *(Oracle*) h;
    int cell = -1;
    for (long i = 0; i < n; ++i) {
        const V3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        cell = locate_any(o, dim, rank, p, std::abs(cell));
        const Sol s = interp_any(o, dim, rank, p, cell);
        cells[i] = cell;
        sol5[5 * i] = s.v.x; sol5[5 * i + 1] = s.v.y; sol5[5 * i + 2] = s.v.z; sol5[5 * i + 3] = s.s1; sol5[5 * i + 4] = s.s2;
    }
}

// SolutionReader.cpp:167-190 calc_interpolation with known cells + :91-112 interp_solution
void fo_interpolate(void* h, int dim, int rank, long n, const double* xyz, const int* cells, double* sol5) {
    Oracle& o = *(Oracle*) h;
    for (long i = 0; i < n; ++i) {
        const V3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const Sol s = interp_any(o, dim, rank, p, std::abs(cells[i]));
        sol5[5 * i] = s.v.x; sol5[5 * i + 1] = s.v.y; sol5[5 * i + 2] = s.v.z; sol5[5 * i + 3] = s.s1; sol5[5 * i + 4] = s.s2;
    }
}

// Pic.cpp:186-196 update_point_cell
void fo_particle_cells(void* h, long n, const double* xyz, int* cell_inout) {
    Oracle& o = *(Oracle*) h;
    for (long i = 0; i < n; ++i) {
        // a negative (lost) cell has no reference-defined guess (deal2femocs is unchecked in Release);
        // the oracle and the product both start such particles from hex 0
        int fc = cell_inout[i] < 0 ? 0 : o.cell2hex[cell_inout[i]];
        fc = hex_locate_cell(o, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, fc);
        cell_inout[i] = fc < 0 ? -1 : o.hex2cell[fc];
    }
}

// Pic.cpp:198-209 field lookup
void fo_particle_field(void* h, long n, const double* xyz, const int* cells, double* E3) {
    Oracle& o = *(Oracle*) h;
    for (long i = 0; i < n; ++i) {
        const V3 E = hex_interp_gradient(o, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, o.cell2hex[cells[i]]);
        E3[3 * i] = E.x; E3[3 * i + 1] = E.y; E3[3 * i + 2] = E.z;
    }
}

void fo_particle_weights(void* h, long n, const double* xyz, const int* cells, double* w8) {
    Oracle& o = *(Oracle*) h;
    for (long i = 0; i < n; ++i)
        hex_shape_functions_dealii(o, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, o.cell2hex[cells[i]], w8 + 8 * i);
}

void fo_linhex_locate(void* h, long n, const double* xyz, int* cell_inout) {
    Oracle& o = *(Oracle*) h;
    for (long i = 0; i < n; ++i) cell_inout[i] = hex_locate_cell(o, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, cell_inout[i]);
}

void fo_nodal_gradient(void* h, long n, const int* hex, const int* node, double* E3) {
    Oracle& o = *(Oracle*) h;
    for (long i = 0; i < n; ++i) {
        const V3 E = hex_nodal_gradient(o, hex[i], node[i]);
        E3[3 * i] = E.x; E3[3 * i + 1] = E.y; E3[3 * i + 2] = E.z;
    }
}

}  // extern "C"
