// FE_Q(2) variant of the field solver (north star: "element-by-element Q1/Q2 stiffness assembly"; BASELINE config 2
// "nanotip_big Q2 Laplace solve").  Option "fe_degree" = 2, set before fb_import_mesh.
//
// The reference fixes the element at compile time (include/DealSolver.h:130-131: shape_degree = 1, quadrature_degree =
// shape_degree + 1); what is built here is what its call sites do when that constant reads 2:
//   src/DealSolver.cpp:368-387   setup_system     -> distribute_dofs of FE_Q(2) (vertex, line, quad and hex dofs, first
//                                                    touch, deal.II's per-cell order) + make_sparsity_pattern
//   src/PoissonSolver.cpp:213-263 assemble_parallel -> 27 x 27 cell matrices, QGauss<3>(3), MappingQ1 geometry
//   src/DealSolver.cpp:389-430   assemble_rhs      -> 9 face shape functions, QGauss<2>(3)
//   src/DealSolver.cpp:432-435   append_dirichlet  -> all 9 dofs of a boundary face
//   src/DealSolver.cpp:317-341   calc_vertex2dof   -> export_solution returns the VERTEX dofs, so the interpolator half
//                                                    of the path is untouched
// Everything behind the CSR pattern (Dirichlet mask, block-JDS SpMV, CG, preconditioners, check_limits) is the Q1 code.
// The space-charge scatter of PoissonSolver.cpp:276-296 (shape_degree != 1) is the FE_Q(2) branch of k_space_charge
// (interp_kernels.cu).  Not provided for FE_Q(2): partitioned meshes, the bulk (current / heat) solvers, export_solution_grad.
#include <omp.h>

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"

namespace {

// local node l = i + 3 j + 9 k of the reference cube, (i, j, k) in {0, 1, 2}^3 <-> xi = i / 2; listed in the order deal.II
// hands out dofs on a cell: 8 vertices, 12 lines, 6 quads (GeometryInfo<3> numbering), the interior
const int DEAL_ORDER[27] = {0, 2, 6, 8, 18, 20, 24, 26, 3, 5, 1, 7, 21, 23, 19, 25, 9, 11, 15, 17, 12, 14, 10, 16, 4, 22, 13};

// entity -> dof tables of the numbering: flat open-addressing maps (linear probing, power-of-two size, never erased);
// std::unordered_map spent 120 ns per look-up here, 4.4e5 look-ups on nanotip_big
struct Key2 { int a, b; bool operator==(const Key2& o) const { return a == o.a && b == o.b; } };
struct Key4 {
    int v[4];
    bool operator==(const Key4& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
};
inline uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
inline uint64_t hash_of(const Key2& k) { return mix64(((uint64_t) (uint32_t) k.a << 32) | (uint32_t) k.b); }
inline uint64_t hash_of(const Key4& k) {
    return mix64((((uint64_t) (uint32_t) k.v[0] << 32) | (uint32_t) k.v[1]) ^ mix64(((uint64_t) (uint32_t) k.v[2] << 32) | (uint32_t) k.v[3]));
}
template <class K>
struct FlatMap {
    std::vector<K> keys; std::vector<int> vals; size_t mask;
    explicit FlatMap(size_t expected) {
        size_t n = 64;
        while (n < 2 * expected) n <<= 1;
        keys.resize(n); vals.assign(n, -1); mask = n - 1;
    }
    int& slot(const K& k) {                  // the value of key k (-1 when k is new: the caller assigns it)
        size_t i = hash_of(k) & mask;
        while (vals[i] >= 0 && !(keys[i] == k)) i = (i + 1) & mask;
        if (vals[i] < 0) keys[i] = k;
        return vals[i];
    }
};

// the local nodes of face f (deal.II face numbering: 2 * axis + side), lexicographic in the two free axes
inline void face_nodes(int f, int out[9]) {
    const int axis = f >> 1, fixed = (f & 1) * 2;
    const int stride[3] = {1, 3, 9};
    const int a0 = axis == 0 ? 1 : 0, a1 = axis == 2 ? 1 : 2;
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) out[a + 3 * b] = fixed * stride[axis] + a * stride[a0] + b * stride[a1];
}

}  // namespace

// Second half of the import for fe_degree 2 (replaces the FE_Q(1) numbering / sparsity / Dirichlet sets of
// fb_host_import_phase2).  Consumes c->h_cv (8 lexicographic vertex ids per cell) and c->bfaces (with ids).
int fb_host_q2_phase2(fb_ctx* c) {
    if (c->part_n_owned >= 0) return c->fail(FB_ERR_ARG, "fe_degree 2 runs on un-partitioned meshes (native sizes)");
    if (c->mesh_kind != 0) return c->fail(FB_ERR_ARG, "fe_degree 2 is provided for the field solver (vacuum mesh) only");
    const bool verbose = getenv("FB_VERBOSE") != nullptr;
    double t_lap = omp_get_wtime();
    auto lap = [&](const char* what) { if (verbose) { const double t = omp_get_wtime(); fprintf(stderr, "[fb] FE_Q(2) import: %s %.2f ms\n", what, 1e3 * (t - t_lap)); t_lap = t; } };
    const int n_vert = c->n_vert, n_cells = c->n_cells;
    const std::vector<int>& cv = c->h_cv;
    c->vertex2dof.assign(n_vert, -1);
    c->cells27.assign(27 * (size_t) n_cells, -1);
    // cube vertices of the vertex / line / quad / cell every local node sits on (the same for all cells)
    int ent_of[27][8], ent_n[27];
    for (int l = 0; l < 27; ++l) {
        const int ijk[3] = {l % 3, (l / 3) % 3, l / 9};
        ent_n[l] = 0;
        for (int v = 0; v < 8; ++v) {
            bool on = true;
            for (int d = 0; d < 3; ++d) on = on && (ijk[d] == 1 || ijk[d] == 2 * ((v >> d) & 1));
            if (on) ent_of[l][ent_n[l]++] = v;
        }
    }
    if (27L * n_cells > 2147483000L) return c->fail(FB_ERR_MESH, "fe_degree 2: dof count exceeds the 32-bit index range");
    FlatMap<Key2> line_dof(4 * (size_t) n_cells);           // (a hexahedral mesh has ~3 lines and ~3 quads per cell)
    FlatMap<Key4> quad_dof(4 * (size_t) n_cells);
    int n_dofs = 0;
    for (int ce = 0; ce < n_cells; ++ce) {
        const int* v8 = &cv[8 * (size_t) ce];
        int* out = &c->cells27[27 * (size_t) ce];
        for (int t = 0; t < 27; ++t) {
            const int l = DEAL_ORDER[t], ne = ent_n[l];
            int* s;
            int fresh = -1;
            if (ne == 1) s = &c->vertex2dof[v8[ent_of[l][0]]];
            else if (ne == 2) {
                const int a = v8[ent_of[l][0]], b = v8[ent_of[l][1]];
                s = &line_dof.slot(Key2{std::min(a, b), std::max(a, b)});
            } else if (ne == 4) {
                int e[4] = {v8[ent_of[l][0]], v8[ent_of[l][1]], v8[ent_of[l][2]], v8[ent_of[l][3]]};
                std::sort(e, e + 4);
                s = &quad_dof.slot(Key4{{e[0], e[1], e[2], e[3]}});
            } else s = &fresh;
            if (*s < 0) *s = n_dofs++;
            out[l] = *s;
        }
    }
    lap("numbering");
    c->n_dofs = c->n_cols = (int) n_dofs;
    c->dof2vertex.assign(n_dofs, -1);
    for (int v = 0; v < n_vert; ++v) c->dof2vertex[c->vertex2dof[v]] = v;
    c->cells_dof.resize(cv.size());                 // the 8 VERTEX dofs per cell: geometry look-ups of the kernels
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long) cv.size(); ++i) c->cells_dof[i] = c->vertex2dof[cv[i]];

    // support points (trilinear image of the reference lattice): vertex, line middle, quad centre, cell centre
    c->q2_xyz.assign(3 * (size_t) n_dofs, 0.0);
#pragma omp parallel for schedule(static)
    for (int ce = 0; ce < n_cells; ++ce) {
        const double* p[8];
        for (int v = 0; v < 8; ++v) p[v] = &c->xyz[3 * (size_t) c->vert2node[cv[8 * (size_t) ce + v]]];
        for (int l = 0; l < 27; ++l) {
            const double xi[3] = {0.5 * (l % 3), 0.5 * ((l / 3) % 3), 0.5 * (l / 9)};
            double s[3] = {0, 0, 0};
            for (int v = 0; v < 8; ++v) {
                const double w = ((v & 1) ? xi[0] : 1 - xi[0]) * ((v & 2) ? xi[1] : 1 - xi[1]) * ((v & 4) ? xi[2] : 1 - xi[2]);
                for (int d = 0; d < 3; ++d) s[d] += w * p[v][d];
            }
            double* dst = &c->q2_xyz[3 * (size_t) c->cells27[27 * (size_t) ce + l]];      // every writer stores the same point
            dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2];
        }
    }

    lap("support points");
    // sparsity: dof -> cells adjacency, then every row gathers the 27 dofs of its cells (sorted, distinct)
    std::vector<int> d2c_off(n_dofs + 1, 0);
    for (size_t i = 0; i < c->cells27.size(); ++i) ++d2c_off[c->cells27[i] + 1];
    for (long r = 0; r < n_dofs; ++r) d2c_off[r + 1] += d2c_off[r];
    std::vector<int> d2c(d2c_off[n_dofs]), fill(d2c_off.begin(), d2c_off.end() - 1);
    for (int ce = 0; ce < n_cells; ++ce)
        for (int l = 0; l < 27; ++l) d2c[fill[c->cells27[27 * (size_t) ce + l]]++] = ce;
    lap("  dof->cells");
    c->rowptr.assign(n_dofs + 1, 0);
    // the 27 dofs of every cell in ascending order: 7 of 8 dofs (lines, quads, interiors) touch at most a handful of cells,
    // whose sorted lists are merged; only the high-valence vertex rows go through a stamp array + sort
    std::vector<int> sorted27(c->cells27);
#pragma omp parallel for schedule(static)
    for (int ce = 0; ce < n_cells; ++ce) std::sort(sorted27.begin() + 27 * (size_t) ce, sorted27.begin() + 27 * (size_t) (ce + 1));
    lap("  sorted cells");
    {
        const int nt = omp_get_max_threads();
        std::vector<std::vector<int>> arena(nt);
        std::vector<int> first_row(nt, -1), cnt(n_dofs, 0);
#pragma omp parallel num_threads(nt)
        {
            const int t = omp_get_thread_num();
            std::vector<int>& mine = arena[t];
            mine.reserve((size_t) n_dofs / nt * 72 + 4096);          // (~60 entries per row: no re-allocation on the way)
            std::vector<int> buf, tmp, stamp;
#pragma omp for schedule(static)
            for (int r = 0; r < (int) n_dofs; ++r) {
                if (first_row[t] < 0) first_row[t] = r;
                const int k = d2c_off[r + 1] - d2c_off[r];
                if (k <= 8) {
                    const int* first = &sorted27[27 * (size_t) d2c[d2c_off[r]]];
                    buf.assign(first, first + 27);
                    for (int q = d2c_off[r] + 1; q < d2c_off[r + 1]; ++q) {
                        const int* cd = &sorted27[27 * (size_t) d2c[q]];
                        tmp.resize(buf.size() + 27);
                        tmp.erase(std::set_union(buf.begin(), buf.end(), cd, cd + 27, tmp.begin()), tmp.end());
                        buf.swap(tmp);
                    }
                } else {
                    if (stamp.empty()) stamp.assign(n_dofs, -1);
                    buf.clear();
                    for (int q = d2c_off[r]; q < d2c_off[r + 1]; ++q) {
                        const int* cd = &c->cells27[27 * (size_t) d2c[q]];
                        for (int j = 0; j < 27; ++j)
                            if (stamp[cd[j]] != r) { stamp[cd[j]] = r; buf.push_back(cd[j]); }
                    }
                    std::sort(buf.begin(), buf.end());
                }
                cnt[r] = (int) buf.size();
                mine.insert(mine.end(), buf.begin(), buf.end());
            }
        }
        lap("  rows");
        long tot = 0;
        for (int r = 0; r < n_dofs; ++r) {
            tot += cnt[r];
            if (tot > 2147483647L) return c->fail(FB_ERR_MESH, "nnz exceeds 32-bit index range");
            c->rowptr[r + 1] = (int) tot;
        }
        c->nnz = tot;
        c->col.resize(tot);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int t = 0; t < nt; ++t)
            if (first_row[t] >= 0) std::copy(arena[t].begin(), arena[t].end(), c->col.begin() + c->rowptr[first_row[t]]);
    }

    lap("sparsity");
    // Dirichlet candidates: every dof of a copper_surface (2) / vacuum_top (8) face; Neumann faces: 9 dofs per top face
    std::vector<unsigned char> on_cu(n_dofs, 0), on_top(n_dofs, 0);
    c->topfaces9.clear();
    for (const auto& bf : c->bfaces) {
        int fn[9]; face_nodes(bf.face, fn);
        for (int k = 0; k < 9; ++k) {
            const int d = c->cells27[27 * (size_t) bf.cell + fn[k]];
            if (bf.id == 2) on_cu[d] = 1;
            if (bf.id == 8) { on_top[d] = 1; c->topfaces9.push_back(d); }
        }
    }
    c->copper_dofs.clear(); c->top_dofs.clear();
    for (int d = 0; d < (int) n_dofs; ++d) { if (on_cu[d]) c->copper_dofs.push_back(d); if (on_top[d]) c->top_dofs.push_back(d); }
    std::vector<int>().swap(c->h_cv); std::vector<int>().swap(c->h_v2c_off); std::vector<int>().swap(c->h_v2c);
    std::vector<unsigned char>().swap(c->h_isb);
    c->mesh_ok = true;
    return FB_OK;
}

namespace fb {

__device__ __forceinline__ void lagrange2(double x, double L[3], double dL[3]) {
    L[0] = 2.0 * (x - 0.5) * (x - 1.0); L[1] = 4.0 * x * (1.0 - x); L[2] = 2.0 * x * (x - 0.5);
    dL[0] = 4.0 * x - 3.0; dL[1] = 4.0 - 8.0 * x; dL[2] = 4.0 * x - 1.0;
}
__device__ __forceinline__ double gauss3_point(int q) { return q == 1 ? 0.5 : (q == 0 ? 0.5 - 0.38729833462074168852 : 0.5 + 0.38729833462074168852); }
__device__ __forceinline__ double gauss3_weight(int q) { return q == 1 ? 8.0 / 18.0 : 5.0 / 18.0; }

// One CTA per hexahedron.  27 threads invert the trilinear Jacobian at the Gauss points, the block then tabulates the
// 27 x 27 physical gradients (17.5 KB of shared memory) and every thread contracts its share of the 729 entries over the
// quadrature points IN POINT ORDER (the order of the reference's q -> i -> j loop) before adding it to its CSR slot,
// found by bisection of the sorted row.  The cell matrix never exists in memory.
__global__ void __launch_bounds__(256) k_q2_stiffness(int n_cells, const int* __restrict__ cells27, const double* __restrict__ vxyz,
                                                      const int* __restrict__ rowptr, const int* __restrict__ col, double* __restrict__ val) {
    __shared__ double s_x[8][3], s_inv[27][9], s_w[27], s_L[3][3], s_dL[3][3];
    __shared__ double s_G[27][27][3];
    __shared__ int s_dof[27];
    const int ce = blockIdx.x, tid = threadIdx.x;
    if (ce >= n_cells) return;
    if (tid < 27) s_dof[tid] = cells27[27 * (size_t) ce + tid];
    if (tid >= 32 && tid < 35) { double L[3], dL[3]; lagrange2(gauss3_point(tid - 32), L, dL); for (int a = 0; a < 3; ++a) { s_L[tid - 32][a] = L[a]; s_dL[tid - 32][a] = dL[a]; } }
    __syncthreads();
    if (tid < 24) {
        const int v = tid / 3, e = tid % 3;
        const int l = 2 * (v & 1) + 6 * ((v >> 1) & 1) + 18 * (v >> 2);
        s_x[v][e] = vxyz[3 * (size_t) s_dof[l] + e];
    }
    __syncthreads();
    if (tid < 27) {
        const int q0 = tid % 3, q1 = (tid / 3) % 3, q2 = tid / 9;
        const double xi[3] = {gauss3_point(q0), gauss3_point(q1), gauss3_point(q2)};
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            double f[3], df[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) { const bool hi = (v >> d) & 1; f[d] = hi ? xi[d] : 1.0 - xi[d]; df[d] = hi ? 1.0 : -1.0; }
            const double dn[3] = {df[0] * f[1] * f[2], f[0] * df[1] * f[2], f[0] * f[1] * df[2]};
#pragma unroll
            for (int e = 0; e < 3; ++e) { J[0][e] += s_x[v][0] * dn[e]; J[1][e] += s_x[v][1] * dn[e]; J[2][e] += s_x[v][2] * dn[e]; }
        }
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][0] * J[2][2] - J[1][2] * J[2][0], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02;
        double* inv = s_inv[tid];                    // inv[3 e + d] = d xi_e / d x_d
        inv[0] = c00 / det;                                     inv[1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; inv[2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        inv[3] = -c01 / det;                                    inv[4] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[5] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        inv[6] = c02 / det;                                     inv[7] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det; inv[8] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        s_w[tid] = det * gauss3_weight(q0) * gauss3_weight(q1) * gauss3_weight(q2);
    }
    __syncthreads();
    for (int p = tid; p < 729; p += blockDim.x) {
        const int q = p / 27, a = p % 27;
        const int q0 = q % 3, q1 = (q / 3) % 3, q2 = q / 9, i = a % 3, j = (a / 3) % 3, k = a / 9;
        const double r0 = s_dL[q0][i] * s_L[q1][j] * s_L[q2][k], r1 = s_L[q0][i] * s_dL[q1][j] * s_L[q2][k], r2 = s_L[q0][i] * s_L[q1][j] * s_dL[q2][k];
        const double* inv = s_inv[q];
        s_G[q][a][0] = r0 * inv[0] + r1 * inv[3] + r2 * inv[6];
        s_G[q][a][1] = r0 * inv[1] + r1 * inv[4] + r2 * inv[7];
        s_G[q][a][2] = r0 * inv[2] + r1 * inv[5] + r2 * inv[8];
    }
    __syncthreads();
    for (int p = tid; p < 729; p += blockDim.x) {
        const int a = p / 27, b = p % 27;
        double sum = 0;
#pragma unroll 9
        for (int q = 0; q < 27; ++q)
            sum += s_w[q] * (s_G[q][a][0] * s_G[q][b][0] + s_G[q][a][1] * s_G[q][b][1] + s_G[q][a][2] * s_G[q][b][2]);
        const int row = s_dof[a], want = s_dof[b];
        int lo = rowptr[row], hi = rowptr[row + 1] - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < want) lo = mid + 1; else hi = mid; }
        atomicAdd(&val[lo], sum);
    }
}

// Neumann load of the top faces with the 9 face shape functions and 3 x 3 Gauss points; the face geometry is bilinear in
// its 4 corners (entries 0, 2, 6, 8 of the face-lexicographic dof list)
__global__ void k_q2_neumann(int n_faces, const int* __restrict__ face_dofs9, const double* __restrict__ vxyz, double bc_value,
                             double* __restrict__ rhs) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    int d[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] = face_dofs9[9 * (size_t) f + i];
    double P[4][3];
    const int corner[4] = {0, 2, 6, 8};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 3; ++e) P[i][e] = vxyz[3 * (size_t) d[corner[i]] + e];
    double r[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < 9; ++q) {
        const double s = gauss3_point(q % 3), t = gauss3_point(q / 3);
        double ds[3], dt[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            ds[e] = (P[1][e] - P[0][e]) * (1 - t) + (P[3][e] - P[2][e]) * t;
            dt[e] = (P[2][e] - P[0][e]) * (1 - s) + (P[3][e] - P[1][e]) * s;
        }
        const double nx = ds[1] * dt[2] - ds[2] * dt[1], ny = ds[2] * dt[0] - ds[0] * dt[2], nz = ds[0] * dt[1] - ds[1] * dt[0];
        const double JxW = sqrt(nx * nx + ny * ny + nz * nz) * gauss3_weight(q % 3) * gauss3_weight(q / 3);
        double Ls[3], Lt[3], unused[3];
        lagrange2(s, Ls, unused); lagrange2(t, Lt, unused);
#pragma unroll
        for (int i = 0; i < 9; ++i) r[i] += Ls[i % 3] * Lt[i / 3] * bc_value * JxW;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) atomicAdd(&rhs[d[i]], r[i]);
}

void launch_q2_stiffness(fb_ctx* c) {
    k_q2_stiffness<<<c->n_cells, 256, 0, c->stream>>>(c->n_cells, c->d_cells27.p, c->d_vxyz.p, c->d_rowptr.p, c->d_col.p, c->d_val_save.p);
    c->launches++;
}

void launch_q2_neumann(fb_ctx* c) {
    if (c->n_top_faces == 0) return;
    k_q2_neumann<<<(c->n_top_faces + 127) / 128, 128, 0, c->stream>>>(c->n_top_faces, c->d_topfaces9.p, c->d_vxyz.p, c->applied_field, c->d_rhs.p);
    c->launches++;
}

}  // namespace fb
