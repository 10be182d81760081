// CUDA kernels (sm_100a) of the solution transfer: nodal field extraction, cell location
// (bit-exact with the reference's sequential chained-guess loop), interpolation, and the
// PIC particle look-ups.  THIS FILE IS COMPILED WITH -fmad=false: the reference is built for
// baseline x86-64 (no FMA contraction) and cell indices are compared bit-for-bit, so every
// expression below keeps the operation order of src/InterpolatorCells.cpp.
//
// Reference behaviour replaced (paths relative to the reference root):
//   src/Interpolator.cpp:103-190            store_solution / store_elfield / average_nodal_fields
//   src/InterpolatorCells.cpp:269-307       InterpolatorCells<dim>::locate_cell
//   src/InterpolatorCells.cpp:631-661       LinearTetrahedra::point_in_cell / shape_functions
//   src/InterpolatorCells.cpp:1272-1416     LinearHexahedra Newton map, shape functions, gradients
//   src/InterpolatorCells.cpp:1462-1562     nodal gradients, point_in_cell, locate_cell
//   src/InterpolatorCells.cpp:1639-1686     LinearTriangles
//   src/InterpolatorCells.cpp:1838-1982     QuadraticTriangles / LinearQuadrangles
//   src/SolutionReader.cpp:43-65,136-190    the per-point driver loops
//   src/Pic.cpp:186-209                     particle cell update and field look-up
//   src/PoissonSolver.cpp:299-319           space-charge right-hand side
//
// These kernels are latency/L2-bound (mesh tables of the native meshes are a few MB and stay
// in the 126 MB L2); the mandatory HBM traffic is 24 B in + 44 B out per point.
#include <cooperative_groups.h>

#include <algorithm>

#include "kernels.h"

namespace fb {

#define FB_ZERO 1e-15      // InterpolatorCells.h:241

struct P3 { double x, y, z; };
__device__ __forceinline__ P3 operator+(P3 a, P3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ P3 operator-(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ P3 operator*(P3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot3(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ P3 cross3(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double det3(P3 a, P3 b, P3 c) {          // InterpolatorCells.cpp:477-480
    return a.x * (b.y * c.z - c.y * b.z) - b.x * (a.y * c.z - c.y * a.z) + c.x * (a.y * b.z - b.y * a.z);
}
__device__ __forceinline__ double dist2(P3 a, const double* __restrict__ b) {    // Primitives.h:247-252
    const double xx = a.x - b[0], yy = a.y - b[1], zz = a.z - b[2];
    return xx * xx + yy * yy + zz * zz;
}
__device__ __forceinline__ P3 ldp(const double* __restrict__ a, long i) { return {a[3 * i], a[3 * i + 1], a[3 * i + 2]}; }
__device__ __forceinline__ P3 p3(const double* a) { return {a[0], a[1], a[2]}; }
__device__ __forceinline__ double comp(const P3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Uniform grid over the tetrahedra (built on the device by fb_interp_initialize): per bucket the searchable tetrahedra
// whose slightly enlarged bounding box overlaps it (`off`/`list`) and the tetrahedra whose centroid lies in it
// (`coff`/`clist`).  It only PRUNES the linear scan of locate_cell: see grid_scan.
struct CellGridDev {
    double lo[3], h[3], inv_h[3]; int g[3]; int n_buckets;
    const int* off; const int* list; const int* coff; const int* clist;
};

struct Tables {                 // device view of the interpolator tables
    const double* nxyz; const double* nodal; const int* hex8;
    const TetRec* tet; const double* tet_cent; const int* tet_mark; const int* tet_nbr_off; const int* tet_nbr; const int* tet4; int n_tet;
    const TriRec* tri; const double* tri_cent; const int* tri_nbr_off; const int* tri_nbr; const int* tri2tet; int n_tri;
    const HexRec* hex; const int* quad2hex; const int* qtet; const int* qtri;
    CellGridDev grid;                                      // uniform-grid filter over the tetrahedra (n_buckets == 0: none)
};

// ---- tetrahedra ----
__device__ __forceinline__ double tet_bary(const TetRec& R, P3 p, int k) {
    return R.det0 * (p.x * R.d[k][0] + p.y * R.d[k][1] + p.z * R.d[k][2] + R.d[k][3]);     // Vec4(point,1).dotProduct
}
__device__ __forceinline__ bool tet_in(const TetRec& R, P3 p) {                          // :631-649
    if (tet_bary(R, p, 0) < -FB_ZERO) return false;
    if (tet_bary(R, p, 1) < -FB_ZERO) return false;
    if (tet_bary(R, p, 2) < -FB_ZERO) return false;
    if (tet_bary(R, p, 3) < -FB_ZERO) return false;
    return true;
}
__device__ __forceinline__ void tet_sf(const TetRec& R, P3 p, double b[4]) {              // :651-661
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = FB_ZERO + tet_bary(R, p, k);
}

// ---- triangles ----
__device__ __forceinline__ bool tri_in(const TriRec& R, P3 p) {                           // :1639-1650
    const P3 tvec = p - p3(R.vert0);
    const double u = dot3(tvec, p3(R.pvec));
    if (u < -FB_ZERO || u > 1 + FB_ZERO) return false;
    const P3 qvec = cross3(tvec, p3(R.edge1));
    const double v = dot3(qvec, p3(R.norm));
    if (v < -FB_ZERO || u + v > 1 + FB_ZERO) return false;
    return fabs(dot3(qvec, p3(R.edge2))) < R.maxd;
}
__device__ __forceinline__ void tri_sf(const TriRec& R, P3 p, double b[3]) {              // :1652-1660
    const P3 tvec = p - p3(R.vert0);
    const P3 qvec = cross3(tvec, p3(R.edge1));
    const double v = dot3(tvec, p3(R.pvec));
    const double w = dot3(qvec, p3(R.norm));
    const double u = 1.0 - v - w;
    b[0] = FB_ZERO + u; b[1] = FB_ZERO + v; b[2] = FB_ZERO + w;
}
__device__ __forceinline__ double tri_fast_distance(const TriRec& R, P3 p) {              // :1743-1747
    const P3 tvec = p - p3(R.vert0);
    const P3 qvec = cross3(tvec, p3(R.edge1));
    return dot3(p3(R.edge2), qvec);
}

// cell families for the generic locate_cell (:269-307)
struct TetFam {
    typedef TetRec Rec;
    static __device__ __forceinline__ const Rec* recs(const Tables& T) { return T.tet; }
    static __device__ __forceinline__ const double* cents(const Tables& T) { return T.tet_cent; }
    static __device__ __forceinline__ int count(const Tables& T) { return T.n_tet; }
    static __device__ __forceinline__ bool searchable(const Tables& T, int c) { return T.tet_mark[c] == 0; }
    static __device__ __forceinline__ const int* nbr_off(const Tables& T) { return T.tet_nbr_off; }
    static __device__ __forceinline__ const int* nbr(const Tables& T) { return T.tet_nbr; }
    static __device__ __forceinline__ bool in(const Rec& R, P3 p) { return tet_in(R, p); }
};
struct TriFam {
    typedef TriRec Rec;
    static __device__ __forceinline__ const Rec* recs(const Tables& T) { return T.tri; }
    static __device__ __forceinline__ const double* cents(const Tables& T) { return T.tri_cent; }
    static __device__ __forceinline__ int count(const Tables& T) { return T.n_tri; }
    static __device__ __forceinline__ bool searchable(const Tables&, int) { return true; }   // lintri markers are all 0
    static __device__ __forceinline__ const int* nbr_off(const Tables& T) { return T.tri_nbr_off; }
    static __device__ __forceinline__ const int* nbr(const Tables& T) { return T.tri_nbr; }
    static __device__ __forceinline__ bool in(const Rec& R, P3 p) { return tri_in(R, p); }
};

// the linear-scan tail of locate_cell for ONE point, straight from global/L2 memory
template <class Fam>
__device__ int scan_all(const Tables& T, P3 p) {
    double min_d2 = 1e100; int min_index = 0;
    const int n = Fam::count(T);
    for (int c = 0; c < n; ++c) {
        if (Fam::searchable(T, c) && Fam::in(Fam::recs(T)[c], p)) return c;
        const double d2 = dist2(p, Fam::cents(T) + 3 * (long) c);
        if (d2 < min_d2) { min_d2 = d2; min_index = c; }
    }
    return -min_index;
}

// The same linear scan done by a whole thread block for ONE point (all threads pass the same p and get the same
// answer).  Warp w tests cells [32 w, 32 w + 32), then strides by the block size; the first hit in index order is
// the minimum over the warps of their own first hit (a warp stops once a hit at a smaller index is known); the
// nearest centroid is the lexicographic minimum of (distance^2, index), i.e. the first index that attains the
// minimum -- exactly what the sequential loop returns.  Used for the few points whose guess and neighbours fail:
// one block per such point (k_finish_scanned, k_particle_scanned) instead of one thread walking all cells.
template <class Fam>
__device__ int block_scan_all(const Tables& T, P3 p) {
    __shared__ int s_first;
    __shared__ double s_d2[32];
    __shared__ int s_idx[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = Fam::count(T);
    __syncthreads();                          // previous use of the scratch is over
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();
    double min_d2 = 1e100; int min_index = 0;
    for (int base = warp * 32; base < n; base += nwarp * 32) {
        if (base > *((volatile int*) &s_first)) break;
        const int c = base + lane;
        bool hit = false;
        if (c < n) {
            hit = Fam::searchable(T, c) && Fam::in(Fam::recs(T)[c], p);
            const double d2 = dist2(p, Fam::cents(T) + 3 * (long) c);
            if (d2 < min_d2) { min_d2 = d2; min_index = c; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) { if (lane == 0) atomicMin(&s_first, base + __ffs(m) - 1); break; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double d2o = __shfl_xor_sync(0xffffffffu, min_d2, o);
        const int io = __shfl_xor_sync(0xffffffffu, min_index, o);
        if (d2o < min_d2 || (d2o == min_d2 && io < min_index)) { min_d2 = d2o; min_index = io; }
    }
    if (lane == 0) { s_d2[warp] = min_d2; s_idx[warp] = min_index; }
    __syncthreads();
    if (s_first != 0x7fffffff) return s_first;
    double d2 = s_d2[0]; int mi = s_idx[0];
    for (int w = 1; w < nwarp; ++w)
        if (s_d2[w] < d2 || (s_d2[w] == d2 && s_idx[w] < mi)) { d2 = s_d2[w]; mi = s_idx[w]; }
    return -mi;
}

// guess -> neighbours -> (fallback) ; returns true on a hit
template <class Fam>
__device__ __forceinline__ bool try_guess(const Tables& T, P3 p, int guess, int& cell) {
    if (guess < 0) return false;
    if (Fam::in(Fam::recs(T)[guess], p)) { cell = guess; return true; }
    const int* off = Fam::nbr_off(T); const int* nb = Fam::nbr(T);
    for (int q = off[guess]; q < off[guess + 1]; ++q) {
        const int c = nb[q];
        if (Fam::in(Fam::recs(T)[c], p)) { cell = c; return true; }
    }
    return false;
}

template <class Fam>
__device__ __forceinline__ int locate_cell(const Tables& T, P3 p, int guess) {
    int cell;
    if (try_guess<Fam>(T, p, guess, cell)) return cell;
    return scan_all<Fam>(T, p);
}

// ---------------------------------------------------------------------------------------
// Guess-free linear scan for MANY points.  A block of 8 warps serves 32 points (4 per warp); the cell
// records are staged through shared memory in tiles (a block reads each record from L2 once) and the
// 32 lanes of a warp test 32 different cells of the tile against the warp's 4 points.  First hit in
// index order = lowest set bit of the first non-empty ballot; nearest centroid = lexicographic minimum
// of (distance^2, index): both identical to the sequential loop of InterpolatorCells.cpp:292-306.
// ---------------------------------------------------------------------------------------
template <class Fam, int TILE>
__global__ void __launch_bounds__(256) k_scan_cells(Tables T, long n, const double* __restrict__ pts, int* __restrict__ out) {
    constexpr int PPW = 4, WARPS = 8;
    constexpr int WORDS = sizeof(typename Fam::Rec) / 8, WP = WORDS | 1;     // odd stride: conflict-free 8-byte reads
    __shared__ double s_rec[TILE * WP];
    __shared__ double s_cent[TILE * 3];
    __shared__ int s_ok[TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long i0 = ((long) blockIdx.x * WARPS + warp) * PPW;
    P3 p[PPW];
    bool found[PPW];
    int result[PPW], min_index[PPW];
    double min_d2[PPW];
#pragma unroll
    for (int q = 0; q < PPW; ++q) {
        const bool active = i0 + q < n;
        p[q] = active ? ldp(pts, i0 + q) : P3{0, 0, 0};
        found[q] = !active; result[q] = 0; min_index[q] = 0; min_d2[q] = 1e100;
    }
    const int n_cells = Fam::count(T);
    for (int base = 0; base < n_cells; base += TILE) {
        const int cnt = min(TILE, n_cells - base);
        __syncthreads();
        const double* src = (const double*) (Fam::recs(T) + base);
        for (int w = threadIdx.x; w < cnt * WORDS; w += blockDim.x) s_rec[(w / WORDS) * WP + (w % WORDS)] = src[w];
        for (int w = threadIdx.x; w < cnt * 3; w += blockDim.x) s_cent[w] = Fam::cents(T)[3 * (long) base + w];
        for (int w = threadIdx.x; w < cnt; w += blockDim.x) s_ok[w] = Fam::searchable(T, base + w);
        __syncthreads();
        for (int sub = 0; sub < cnt; sub += 32) {
            const int k = sub + lane;
            const bool valid = k < cnt;
            const typename Fam::Rec& R = *reinterpret_cast<const typename Fam::Rec*>(&s_rec[(valid ? k : 0) * WP]);
            const bool ok = valid && s_ok[valid ? k : 0];
#pragma unroll
            for (int q = 0; q < PPW; ++q) {
                if (found[q]) continue;                      // warp-uniform
                const bool hit = ok && Fam::in(R, p[q]);
                if (valid) {
                    const double d2 = dist2(p[q], s_cent + 3 * k);
                    if (d2 < min_d2[q]) { min_d2[q] = d2; min_index[q] = base + k; }
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) { found[q] = true; result[q] = base + sub + __ffs(m) - 1; }
            }
        }
        bool all = true;
#pragma unroll
        for (int q = 0; q < PPW; ++q) all &= found[q];
        if (__syncthreads_and(all)) break;
    }
#pragma unroll
    for (int q = 0; q < PPW; ++q) {
        if (i0 + q >= n) continue;
        double d2 = min_d2[q]; int mi = min_index[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double d2o = __shfl_xor_sync(0xffffffffu, d2, o);
            const int io = __shfl_xor_sync(0xffffffffu, mi, o);
            if (d2o < d2 || (d2o == d2 && io < mi)) { d2 = d2o; mi = io; }
        }
        if (lane == 0) out[i0 + q] = found[q] ? result[q] : -mi;
    }
}


// ---------------------------------------------------------------------------------------
// The guess-free answer of locate_cell for tetrahedra through the uniform grid -- bit-identical to the linear scan:
//   * hit: a tetrahedron that contains p (barycentric test with its 1e-15 tolerance) has p inside its bounding box
//     enlarged by the build margin, hence sits in the list of p's bucket; the FIRST hit in index order of the scan is
//     the minimum index over the hits in that list (lists are unordered: every candidate is tested);
//   * miss: the scan returns the lexicographic minimum of (distance^2 to the centroid, index) over ALL tetrahedra.
//     Centroids are binned one bucket each; shells of buckets around p's (clamped) bucket are visited until the best
//     distance is strictly below the distance to everything unvisited (the faces of the visited box that are not grid
//     boundaries, minus a safety margin far above round-off); ties keep the smaller index, as the ascending loop does.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int grid_bucket(const CellGridDev& G, double x, int d) {
    const int b = (int) floor((x - G.lo[d]) * G.inv_h[d]);
    return min(max(b, 0), G.g[d] - 1);
}

// one WARP per point: all 32 lanes pass the same p and obtain the same answer
__device__ int grid_scan(const Tables& T, P3 p) {
    const CellGridDev& G = T.grid;
    const int lane = threadIdx.x & 31;
    const double pc[3] = {p.x, p.y, p.z};
    int b[3]; bool inside = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        b[d] = grid_bucket(G, pc[d], d);
        inside = inside && pc[d] >= G.lo[d] && pc[d] <= G.lo[d] + G.h[d] * G.g[d];
    }
    if (inside) {
        const int bk = (b[2] * G.g[1] + b[1]) * G.g[0] + b[0];
        int hit = 0x7fffffff;
        for (int k = G.off[bk] + lane; k < G.off[bk + 1]; k += 32) {
            const int c = G.list[k];
            if (c < hit && tet_in(T.tet[c], p)) hit = c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hit = min(hit, __shfl_xor_sync(0xffffffffu, hit, o));
        if (hit != 0x7fffffff) return hit;
    }
    double best = 1e100; int bi = 0;
    const double safety = 1e-9 * fmin(G.h[0], fmin(G.h[1], G.h[2]));
    for (int r = 0;; ++r) {
        const int z0 = max(0, b[2] - r), z1 = min(G.g[2] - 1, b[2] + r);
        const int y0 = max(0, b[1] - r), y1 = min(G.g[1] - 1, b[1] + r);
        const int x0 = max(0, b[0] - r), x1 = min(G.g[0] - 1, b[0] + r);
        const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, total = nx * ny * (z1 - z0 + 1);
        for (int q = lane; q < total; q += 32) {                   // the lanes share the box of shell r; its interior was visited before
            const int x = x0 + q % nx, y = y0 + (q / nx) % ny, z = z0 + q / (nx * ny);
            if (max(abs(x - b[0]), max(abs(y - b[1]), abs(z - b[2]))) != r) continue;
            const int bk = (z * G.g[1] + y) * G.g[0] + x;
            for (int k = G.coff[bk]; k < G.coff[bk + 1]; ++k) {
                const int c = G.clist[k];
                const double d2 = dist2(p, T.tet_cent + 3 * (long) c);
                if (d2 < best || (d2 == best && c < bi)) { best = d2; bi = c; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double d2o = __shfl_xor_sync(0xffffffffu, best, o);
            const int io = __shfl_xor_sync(0xffffffffu, bi, o);
            if (d2o < best || (d2o == best && io < bi)) { best = d2o; bi = io; }
        }
        double dmin = 1e300;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (b[d] - r > 0) dmin = fmin(dmin, pc[d] - (G.lo[d] + G.h[d] * (b[d] - r)));
            if (b[d] + r < G.g[d] - 1) dmin = fmin(dmin, (G.lo[d] + G.h[d] * (b[d] + r + 1)) - pc[d]);
        }
        if (dmin == 1e300) break;                         // the whole grid has been visited
        dmin -= safety;
        if (dmin > 0 && best < dmin * dmin) break;
    }
    return -bi;
}

__global__ void __launch_bounds__(128) k_scan_grid(Tables T, long n, const double* __restrict__ pts, int* __restrict__ out) {
    const long i = ((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;         // warp-uniform
    if (i >= n) return;
    const int c = grid_scan(T, ldp(pts, i));
    if ((threadIdx.x & 31) == 0) out[i] = c;
}

// ---- grid build: count, (host-free) scan, fill ----
__device__ __forceinline__ void tet_bucket_range(const Tables& T, const CellGridDev& G, int t, double margin, int lo[3], int hi[3], int cb[3]) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double* v = T.nxyz + 3 * (long) T.tet4[4 * (long) t + k];
#pragma unroll
        for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], v[d]); mx[d] = fmax(mx[d], v[d]); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = grid_bucket(G, mn[d] - margin, d); hi[d] = grid_bucket(G, mx[d] + margin, d);
        cb[d] = grid_bucket(G, T.tet_cent[3 * (long) t + d], d);
    }
}
template <bool FILL>
__global__ void __launch_bounds__(128) k_grid_build(Tables T, CellGridDev G, double margin, int* __restrict__ cnt, int* __restrict__ ccnt,
                                                    int* __restrict__ list, int* __restrict__ clist, int cap) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T.n_tet) return;
    int lo[3], hi[3], cb[3];
    tet_bucket_range(T, G, t, margin, lo, hi, cb);
    const int cbk = (cb[2] * G.g[1] + cb[1]) * G.g[0] + cb[0];
    const int pos = atomicAdd(&ccnt[cbk], 1);
    if (FILL) clist[pos] = t;
    if (T.tet_mark[t] != 0) return;                       // not searchable: never a hit of the linear scan
    for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int x = lo[0]; x <= hi[0]; ++x) {
                const int bk = (z * G.g[1] + y) * G.g[0] + x;
                const int q = atomicAdd(&cnt[bk], 1);
                if (FILL && q < cap) list[q] = t;
            }
}
// exclusive scan of two count arrays by one block (n up to a few 1e5); cursors start at the offsets
__global__ void __launch_bounds__(1024) k_grid_scan(int n, int* __restrict__ cnt, int* __restrict__ off, int* __restrict__ ccnt, int* __restrict__ coff,
                                                    int* __restrict__ totals) {
    __shared__ int s_sum[1024];
    for (int which = 0; which < 2; ++which) {
        int* c = which ? ccnt : cnt; int* o = which ? coff : off;
        const int per = (n + blockDim.x - 1) / blockDim.x;
        const int a = threadIdx.x * per, e = min(n, a + per);
        int s = 0;
        for (int i = a; i < e; ++i) s += c[i];
        s_sum[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x == 0) { int run = 0; for (int i = 0; i < (int) blockDim.x; ++i) { const int v = s_sum[i]; s_sum[i] = run; run += v; } o[n] = run; totals[which] = run; }
        __syncthreads();
        int run = s_sum[threadIdx.x];
        for (int i = a; i < e; ++i) { const int v = c[i]; o[i] = run; c[i] = run; run += v; }     // c becomes the fill cursor
        __syncthreads();
    }
}

// One Jacobi sweep of the chained-guess recurrence  T_i = F(p_i, |T_{i-1}|), T_{-1} = first_guess:
// a point is recomputed only if its predecessor changed in the previous sweep.
template <class Fam>
__global__ void __launch_bounds__(128) k_chain_sweep(Tables T, long n, const double* __restrict__ pts, const int* __restrict__ scan,
                                                     const int* __restrict__ prev, int* __restrict__ next,
                                                     const unsigned char* __restrict__ dirty_prev, unsigned char* __restrict__ dirty_next,
                                                     int first_guess, int first_sweep, int* __restrict__ changed, long chain_len) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool start = chain_len > 0 ? (i % chain_len == 0) : (i == 0);     // first point of an independent chain
    const bool need = first_sweep || (!start && dirty_prev[i - 1]);
    if (!need) { next[i] = prev[i]; dirty_next[i] = 0; return; }
    const int guess = start ? first_guess : abs(prev[i - 1]);
    int cell;
    if (!try_guess<Fam>(T, ldp(pts, i), guess, cell)) cell = scan[i];
    next[i] = cell;
    const bool ch = (cell != prev[i]);
    dirty_next[i] = ch;
    if (ch) *changed = 1;
}

// The whole fix-point iteration of the chained-guess recurrence in ONE cooperative launch: Jacobi
// sweeps separated by grid-wide barriers, convergence flag on the device (three rotating flags so
// that resetting the next flag never races with readers of the previous one), result copied to
// `out`.  No host round trip; flags[3] returns the number of sweeps.
template <class Fam>
__global__ void __launch_bounds__(128) k_chain_fixpoint(Tables T, long n, const double* __restrict__ pts, const int* __restrict__ scan,
                                                        int* bufA, int* bufB, unsigned char* dirtyA, unsigned char* dirtyB,
                                                        int first_guess, int* flags, int* __restrict__ out, long chain_len) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const long tid = (long) blockIdx.x * blockDim.x + threadIdx.x, stride = (long) gridDim.x * blockDim.x;
    const int* prev = scan; int* next = bufA;                  // the first sweep reads `scan`, which is never written
    unsigned char* dcur = dirtyA; unsigned char* dprev = dirtyB;   // dprev is not read in the first sweep
    int sweep = 0;
    while (true) {
        if (tid == 0) flags[(sweep + 1) % 3] = 0;
        bool any = false;
        for (long i = tid; i < n; i += stride) {
            // buffers written by other CTAs in the previous sweep are read through L2 (ld.global.cg)
            // chain_len > 0: the points form independent chains of that length (EmissionReader::emission_line: every
            // line of 32 points is a fresh SolutionReader, i.e. restarts from the first guess)
            const bool start = chain_len > 0 ? (i % chain_len == 0) : (i == 0);
            const bool need = sweep == 0 || (!start && __ldcg(&dprev[i - 1]));
            const int old = __ldcg(&prev[i]);
            if (!need) { next[i] = old; dcur[i] = 0; continue; }
            const int guess = start ? first_guess : abs(__ldcg(&prev[i - 1]));
            int cell;
            if (!try_guess<Fam>(T, ldp(pts, i), guess, cell)) cell = scan[i];
            next[i] = cell;
            const bool ch = (cell != old);
            dcur[i] = ch;
            any |= ch;
        }
        if (any) atomicOr(&flags[sweep % 3], 1);
        grid.sync();
        const int changed = *((volatile int*) &flags[sweep % 3]);
        prev = next;
        next = (prev == bufA) ? bufB : bufA;
        unsigned char* t = dcur; dcur = dprev; dprev = t;
        ++sweep;
        if (!changed) break;
    }
    for (long i = tid; i < n; i += stride) out[i] = __ldcg(&prev[i]);
    if (tid == 0) flags[3] = sweep;
}

// ---- hexahedra -------------------------------------------------------------------------
__device__ void hex_nat_coords(const HexRec& H, P3 point, double& u, double& v, double& w) {     // :1272-1317
    const P3 f0 = point - p3(H.f[0]);
    const P3 f1 = p3(H.f[1]), f2 = p3(H.f[2]), f3 = p3(H.f[3]), f4 = p3(H.f[4]), f5 = p3(H.f[5]), f6 = p3(H.f[6]), f7 = p3(H.f[7]);
    u = 0; v = 0; w = 0;
    for (int i = 0; i < 20; ++i) {
        const P3 f = (f0 - f1 * u - f2 * v - f3 * w - f4 * (u * v) - f5 * (u * w) - f6 * (v * w) - f7 * (u * v * w));
        const P3 fu = f1 + f4 * v + f5 * w + f7 * (v * w);
        const P3 fv = f2 + f4 * u + f6 * w + f7 * (u * w);
        const P3 fw = f3 + f5 * u + f6 * v + f7 * (u * v);
        double D = det3(fu, fv, fw);
        D = 1.0 / D;
        const double du = det3(f, fv, fw) * D;
        const double dv = det3(fu, f, fw) * D;
        const double dw = det3(fu, fv, f) * D;
        u += du; v += dv; w += dw;
        if (du * du + dv * dv + dw * dw < FB_ZERO) return;
    }
}

__device__ __forceinline__ void hex_sf(const HexRec& H, P3 point, double sf[8]) {                 // :1334-1353
    double u, v, w;
    hex_nat_coords(H, point, u, v, w);
    sf[0] = (1 - u) * (1 - v) * (1 - w) / 8.0;
    sf[1] = (1 + u) * (1 - v) * (1 - w) / 8.0;
    sf[2] = (1 + u) * (1 + v) * (1 - w) / 8.0;
    sf[3] = (1 - u) * (1 + v) * (1 - w) / 8.0;
    sf[4] = (1 - u) * (1 - v) * (1 + w) / 8.0;
    sf[5] = (1 + u) * (1 - v) * (1 + w) / 8.0;
    sf[6] = (1 + u) * (1 + v) * (1 + w) / 8.0;
    sf[7] = (1 - u) * (1 + v) * (1 + w) / 8.0;
}

__device__ __forceinline__ bool hex_in(const Tables& T, P3 p, int cell) {                         // :1507-1528
    double b[4];
    tet_sf(T.tet[cell / 4], p, b);
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

// tail of LinearHexahedra::locate_cell (:1544-1561): signed tet -> signed hex
__device__ __forceinline__ int hex_from_tet(const Tables& T, P3 p, int tet_signed) {
    const int sign = tet_signed < 0 ? -1 : 1;
    const int tet = abs(tet_signed);
    double b[4];
    tet_sf(T.tet[tet], p, b);
    if (b[0] >= b[1] && b[0] >= b[2] && b[0] >= b[3]) return sign * (4 * tet + 0);
    if (b[1] >= b[0] && b[1] >= b[2] && b[1] >= b[3]) return sign * (4 * tet + 1);
    if (b[2] >= b[0] && b[2] >= b[1] && b[2] >= b[3]) return sign * (4 * tet + 2);
    if (b[3] >= b[0] && b[3] >= b[1] && b[3] >= b[2]) return sign * (4 * tet + 3);
    return -1;
}
// tail of LinearQuadrangles::locate_cell (:1961-1981)
__device__ __forceinline__ int quad_from_tri(const Tables& T, P3 p, int tri_signed) {
    const int sign = tri_signed < 0 ? -1 : 1;
    const int tri = abs(tri_signed);
    double b[3];
    tri_sf(T.tri[tri], p, b);
    if (b[0] >= b[1] && b[0] >= b[2]) return sign * (3 * tri + 0);
    if (b[1] >= b[0] && b[1] >= b[2]) return sign * (3 * tri + 1);
    if (b[2] >= b[0] && b[2] >= b[1]) return sign * (3 * tri + 2);
    return -1;
}

__device__ __forceinline__ int hex_locate(const Tables& T, P3 p, int guess_hex) {                 // :1536-1562
    return hex_from_tet(T, p, locate_cell<TetFam>(T, p, guess_hex / 4));
}

// ---- weighted sums of nodal Solutions (:359-380) ----
template <int N>
__device__ __forceinline__ void weighted(const Tables& T, const int* __restrict__ nodes, const double* w, double out[5]) {
    double vx = 0, vy = 0, vz = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double* s = T.nodal + 5 * (long) nodes[i];
        vx += s[0] * w[i]; vy += s[1] * w[i]; vz += s[2] * w[i];
        s1 += s[3] * w[i]; s2 += s[4] * w[i];
    }
    out[0] = vx; out[1] = vy; out[2] = vz; out[3] = s1; out[4] = s2;
}

__device__ void tet_interp(const Tables& T, P3 p, int c, double out[5]) {
    const int cell = abs(c);
    double w[4];
    tet_sf(T.tet[cell], p, w);
    weighted<4>(T, T.tet4 + 4 * (long) cell, w, out);
}
__device__ void qtet_interp(const Tables& T, P3 p, int c, double out[5]) {                        // :795-816
    const int cell = abs(c);
    double b[4];
    tet_sf(T.tet[cell], p, b);
    const double b1 = b[0], b2 = b[1], b3 = b[2], b4 = b[3];
    const double w[10] = {b1 * (2 * b1 - 1), b2 * (2 * b2 - 1), b3 * (2 * b3 - 1), b4 * (2 * b4 - 1),
                          4 * b1 * b2, 4 * b2 * b3, 4 * b3 * b1, 4 * b1 * b4, 4 * b2 * b4, 4 * b3 * b4};
    weighted<10>(T, T.qtet + 10 * (long) cell, w, out);
}
__device__ void hex_interp(const Tables& T, P3 p, int c, double out[5]) {
    const int cell = abs(c);
    double w[8];
    hex_sf(T.hex[cell], p, w);
    weighted<8>(T, T.hex8 + 8 * (long) cell, w, out);
}

// dim = 2 variants hop from the surface cell to the adjacent 3D cell (:1662-1686, :1838-1861, :1898-1920).
// Split in two so that the linear-scan tail of locate_cell (:292-306) can be done warp-cooperatively in between:
//   surface_resolve : cheap part; returns false when guess and neighbours failed (the scan decides),
//   surface_finish  : interpolation in the resolved 3D cell.
// `cell3d` is a tetrahedron (signed after a scan) except for rank 3 when the quadrangle's own hexahedron is taken.
__device__ bool surface_resolve(const Tables& T, int rank, P3 p, int c, int& cell3d, bool& is_hex) {
    const int cell = abs(c);
    is_hex = false;
    if (rank == 3) {
        const int h0 = T.quad2hex[2 * cell], h1 = T.quad2hex[2 * cell + 1];
        const double d = fabs(tri_fast_distance(T.tri[cell / 3], p));
        if (d <= 100.0 * FB_ZERO || hex_in(T, p, h0)) { cell3d = h0; is_hex = true; return true; }
        return try_guess<TetFam>(T, p, h1 / 4, cell3d);            // hex_locate: tet search from the other hexahedron
    }
    const int t0 = T.tri2tet[2 * cell], t1 = T.tri2tet[2 * cell + 1];
    const double d = fabs(tri_fast_distance(T.tri[cell], p));
    if (d <= 100.0 * FB_ZERO || tet_in(T.tet[t0], p)) { cell3d = t0; return true; }
    return try_guess<TetFam>(T, p, t1, cell3d);
}
__device__ void surface_finish(const Tables& T, int rank, P3 p, int cell3d, bool is_hex, double out[5]) {
    if (rank == 3) hex_interp(T, p, is_hex ? cell3d : hex_from_tet(T, p, cell3d), out);
    else if (rank == 1) tet_interp(T, p, cell3d, out);
    else qtet_interp(T, p, cell3d, out);
}

// final stage of locate_interpolate: base-family cell -> reported cell + interpolation.  Surface points whose
// hand-over to the 3D cell needs the linear scan are appended to `needy` and finished by k_finish_scanned.
__global__ void __launch_bounds__(128) k_finish_interp(Tables T, int dim, int rank, long n, const double* __restrict__ pts,
                                                       const int* __restrict__ base_cells, int cells_are_final,
                                                       int* __restrict__ cells_out, double* __restrict__ sol5,
                                                       int* __restrict__ needy_count, int* __restrict__ needy_idx) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const P3 p = ldp(pts, i);
    int cell = base_cells[i];
    if (!cells_are_final && rank == 3) cell = (dim == 2) ? quad_from_tri(T, p, cell) : hex_from_tet(T, p, cell);
    if (cells_out) cells_out[i] = cell;
    double out[5];
    if (dim == 2) {
        int cell3d; bool is_hex;
        if (!surface_resolve(T, rank, p, cell, cell3d, is_hex)) { needy_idx[atomicAdd(needy_count, 1)] = (int) i; return; }
        surface_finish(T, rank, p, cell3d, is_hex, out);
    }
    else if (rank == 1) tet_interp(T, p, cell, out);
    else if (rank == 2) qtet_interp(T, p, cell, out);
    else hex_interp(T, p, cell, out);
#pragma unroll
    for (int k = 0; k < 5; ++k) sol5[5 * i + k] = out[k];
}

// one block per deferred surface point: block-cooperative linear scan over the tetrahedra, then the interpolation
__global__ void __launch_bounds__(256) k_finish_scanned(Tables T, int rank, const double* __restrict__ pts,
                                                        const int* __restrict__ needy_count, const int* __restrict__ needy_idx,
                                                        double* __restrict__ sol5) {
    const int count = *needy_count;
    if (T.grid.n_buckets > 0) {         // grid filter: the scan of one point is a handful of candidates -> one WARP per deferred point
        for (long e = ((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < count; e += ((long) gridDim.x * blockDim.x) >> 5) {
            const long i = needy_idx[e];
            const P3 p = ldp(pts, i);
            const int tet = grid_scan(T, p);
            if ((threadIdx.x & 31) == 0) {
                double out[5];
                surface_finish(T, rank, p, tet, false, out);
#pragma unroll
                for (int k = 0; k < 5; ++k) sol5[5 * i + k] = out[k];
            }
        }
        return;
    }
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const long i = needy_idx[e];
        const P3 p = ldp(pts, i);
        const int tet = block_scan_all<TetFam>(T, p);
        if (threadIdx.x == 0) {
            double out[5];
            surface_finish(T, rank, p, tet, false, out);
#pragma unroll
            for (int k = 0; k < 5; ++k) sol5[5 * i + k] = out[k];
        }
    }
}

// ---- gradients ----------------------------------------------------------------------
__device__ P3 hex_point_gradient(const Tables& T, P3 point, int hex) {                            // :1363-1416, :410-423
    double u, v, w;
    hex_nat_coords(T.hex[hex], point, u, v, w);
    const int* shex = T.hex8 + 8 * (long) hex;
    P3 dN[8] = {
        {-(1 - v) * (1 - w), -(1 - u) * (1 - w), -(1 - u) * (1 - v)},
        { (1 - v) * (1 - w), -(1 + u) * (1 - w), -(1 + u) * (1 - v)},
        { (1 + v) * (1 - w),  (1 + u) * (1 - w), -(1 + u) * (1 + v)},
        {-(1 + v) * (1 - w),  (1 - u) * (1 - w), -(1 - u) * (1 + v)},
        {-(1 - v) * (1 + w), -(1 - u) * (1 + w),  (1 - u) * (1 - v)},
        { (1 - v) * (1 + w), -(1 + u) * (1 + w),  (1 + u) * (1 - v)},
        { (1 + v) * (1 + w),  (1 + u) * (1 + w),  (1 + u) * (1 + v)},
        {-(1 + v) * (1 + w),  (1 - u) * (1 + w),  (1 - u) * (1 + v)}};
#pragma unroll
    for (int i = 0; i < 8; ++i) dN[i] = dN[i] * 0.125;
    P3 J[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const P3 x = ldp(T.nxyz, shex[k]);
        J[0] = J[0] + x * dN[k].x; J[1] = J[1] + x * dN[k].y; J[2] = J[2] + x * dN[k].z;
    }
    double Jdet = det3(J[0], J[1], J[2]);
    Jdet = 1.0 / Jdet;
    P3 Jinv[3] = {
        {J[1].y * J[2].z - J[2].y * J[1].z, J[2].x * J[1].z - J[1].x * J[2].z, J[1].x * J[2].y - J[2].x * J[1].y},
        {J[2].y * J[0].z - J[0].y * J[2].z, J[0].x * J[2].z - J[2].x * J[0].z, J[2].x * J[0].y - J[0].x * J[2].y},
        {J[0].y * J[1].z - J[1].y * J[0].z, J[1].x * J[0].z - J[0].x * J[1].z, J[0].x * J[1].y - J[1].x * J[0].y}};
#pragma unroll
    for (int i = 0; i < 3; ++i) Jinv[i] = Jinv[i] * Jdet;
    P3 r = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        P3 g = {0, 0, 0};
        g = g + Jinv[0] * dN[k].x; g = g + Jinv[1] * dN[k].y; g = g + Jinv[2] * dN[k].z;
        r = r - g * T.nodal[5 * (long) shex[k] + 4];
    }
    return r;
}

// closed-form nodal gradient (:1319-1332, :1462-1505) and interp_gradient(cell,node) (:425-439)
__device__ P3 hex_nodal_gradient(const double* __restrict__ nxyz, const double* __restrict__ nodal,
                                 const int* __restrict__ hex8, int hex, int node) {
    const double su = (node == 1 || node == 2 || node == 5 || node == 6) ? 1.0 : -1.0;
    const double sv = (node == 2 || node == 3 || node == 6 || node == 7) ? 1.0 : -1.0;
    const double sw = (node >= 4) ? 1.0 : -1.0;
    int n1, n2, n3;
    switch (node) {
        case 0: n1 = 1; n2 = 3; n3 = 4; break;
        case 1: n1 = 0; n2 = 2; n3 = 5; break;
        case 2: n1 = 3; n2 = 1; n3 = 6; break;
        case 3: n1 = 2; n2 = 0; n3 = 7; break;
        case 4: n1 = 5; n2 = 7; n3 = 0; break;
        case 5: n1 = 4; n2 = 6; n3 = 1; break;
        case 6: n1 = 7; n2 = 5; n3 = 2; break;
        default: n1 = 6; n2 = 4; n3 = 3; break;
    }
    const int* shex = hex8 + 8 * (long) hex;
    const int g0 = shex[node], g1 = shex[n1], g2 = shex[n2], g3 = shex[n3];
    const P3 vec0 = ldp(nxyz, g0);
    const P3 J0 = (vec0 - ldp(nxyz, g1)) * (su * 0.5);
    const P3 J1 = (vec0 - ldp(nxyz, g2)) * (sv * 0.5);
    const P3 J2 = (vec0 - ldp(nxyz, g3)) * (sw * 0.5);
    double Jdet = det3(J0, J1, J2);
    Jdet = 1.0 / Jdet;
    P3 Jinv[3] = {
        {J1.y * J2.z - J2.y * J1.z, J2.y * J0.z - J0.y * J2.z, J0.y * J1.z - J1.y * J0.z},
        {J2.x * J1.z - J1.x * J2.z, J0.x * J2.z - J2.x * J0.z, J1.x * J0.z - J0.x * J1.z},
        {J1.x * J2.y - J2.x * J1.y, J2.x * J0.y - J0.x * J2.y, J0.x * J1.y - J1.x * J0.y}};
#pragma unroll
    for (int i = 0; i < 3; ++i) Jinv[i] = Jinv[i] * Jdet;
    const P3 uvw = {su, sv, sw};
    // sfg[node][i] = 0.5 * Jinv[i].uvw ; sfg[n_j][i] = -0.5 * Jinv[i][j] * uvw[j] ; other nodes 0
    P3 s0, s1, s2, s3;
    s0.x = 0.5 * dot3(Jinv[0], uvw); s0.y = 0.5 * dot3(Jinv[1], uvw); s0.z = 0.5 * dot3(Jinv[2], uvw);
    s1.x = -0.5 * Jinv[0].x * su; s1.y = -0.5 * Jinv[1].x * su; s1.z = -0.5 * Jinv[2].x * su;
    s2.x = -0.5 * Jinv[0].y * sv; s2.y = -0.5 * Jinv[1].y * sv; s2.z = -0.5 * Jinv[2].y * sv;
    s3.x = -0.5 * Jinv[0].z * sw; s3.y = -0.5 * Jinv[1].z * sw; s3.z = -0.5 * Jinv[2].z * sw;
    // vector_i -= sfg[i] * phi_i for i = 0..7 in local order (zero rows subtract +-0)
    P3 r = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        P3 s = {0, 0, 0};
        if (i == node) s = s0; else if (i == n1) s = s1; else if (i == n2) s = s2; else if (i == n3) s = s3;
        r = r - s * nodal[5 * (long) shex[i] + 4];
    }
    return r;
}

// Interpolator::store_solution (:103-123): Solution(Vec3(0), charge_dens, potential) on vacuum nodes
__global__ void k_store_solution(int n_nodes, const int* __restrict__ node2vert, const int* __restrict__ vertex2dof,
                                 const double* __restrict__ x, const double* __restrict__ rho, double* __restrict__ nodal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int v = node2vert[i];
    double* s = nodal + 5 * (long) i;
    s[0] = 0; s[1] = 0; s[2] = 0;
    s[3] = (v >= 0 && rho) ? rho[vertex2dof[v]] : 0.0;
    s[4] = v >= 0 ? x[vertex2dof[v]] : 0.0;
}

// Interpolator::store_elfield (:125-140): mean over incident vacuum hexes of the nodal gradient
__global__ void __launch_bounds__(128) k_nodal_field(int n_nodes, const int* __restrict__ node2vert, const int* __restrict__ n2c_off,
                                                     const int* __restrict__ n2c_list, const double* __restrict__ nxyz,
                                                     const int* __restrict__ hex8, double* nodal) {
    // reads only the potentials (slot 4), writes only the vectors (slots 0..2): race free
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || node2vert[i] < 0) return;
    P3 mean = {0, 0, 0};
    const int lo = n2c_off[i], hi = n2c_off[i + 1];
    if (hi > lo) {
        for (int q = lo; q < hi; ++q) {
            const int e = n2c_list[q];
            mean = mean + hex_nodal_gradient(nxyz, nodal, hex8, e >> 3, e & 7);
        }
        mean = mean * (1.0 / (hi - lo));
    }
    nodal[5 * (long) i] = mean.x; nodal[5 * (long) i + 1] = mean.y; nodal[5 * (long) i + 2] = mean.z;
}

// Interpolator::average_nodal_fields (:142-170): only tet nodes are written, only centroid-type
// nodes are read (TetgenMesh.cpp:819-837), so the in-place update is order independent.
__global__ void __launch_bounds__(128) k_smooth(int n_voro, const int* __restrict__ off, const int* __restrict__ list,
                                                const double* __restrict__ nxyz, double decay, double* __restrict__ nodal) {
    // one warp per tet node: the lanes stride its (long, uneven) list of pseudo-Voronoi neighbours; the weighted sums
    // are reduced with shuffles (summation order differs from the sequential loop by ~1e-16 relative)
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_voro) return;
    const int lo = off[i], hi = off[i + 1];
    if (hi == lo) return;
    const P3 tetnode = ldp(nxyz, i);
    P3 vec = {0, 0, 0};
    double w_sum = 0;
    for (int q = lo + lane; q < hi; q += 32) {
        const int nb = list[q];
        const double w = exp(decay * sqrt(dist2(tetnode, nxyz + 3 * (long) nb)));
        w_sum += w;
        vec = vec + P3{nodal[5 * (long) nb], nodal[5 * (long) nb + 1], nodal[5 * (long) nb + 2]} * w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w_sum += __shfl_xor_sync(0xffffffffu, w_sum, o);
        vec.x += __shfl_xor_sync(0xffffffffu, vec.x, o); vec.y += __shfl_xor_sync(0xffffffffu, vec.y, o); vec.z += __shfl_xor_sync(0xffffffffu, vec.z, o);
    }
    if (lane == 0 && w_sum > 0) {
        vec = vec * (1.0 / w_sum);
        nodal[5 * (long) i] = vec.x; nodal[5 * (long) i + 1] = vec.y; nodal[5 * (long) i + 2] = vec.z;
    }
}

// ---- PIC particles --------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_particle_cells(Tables T, long n, const double* __restrict__ pts, const int* __restrict__ cell2hex, int n_cells,
                                                        const int* __restrict__ hex2cell, int* __restrict__ cell_inout,
                                                        int* __restrict__ needy_count, int* __restrict__ needy_idx, int lost_marker) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const P3 p = ldp(pts, i);
    const int c = cell_inout[i];
    if (c == lost_marker) { cell_inout[i] = -1; return; }     // left the simulation box in k_pic_move: no search (Pic.cpp:177-180)
    // a stale cell id from before a re-mesh (Pic.cpp:188-191 anticipates it; deal2femocs answers -2 past the end of its
    // map, InterpolatorCells.h:437-448) is no guess at all: start from cell 0 like a negative one
    const int guess = (c < 0 || c >= n_cells) ? 0 : cell2hex[c];
    int tet;
    // hex_locate (:1536-1562): tetrahedron from the guess / its neighbours; particles that left the neighbourhood
    // are deferred to the block-cooperative scan of k_particle_scanned
    if (!try_guess<TetFam>(T, p, guess / 4, tet)) { needy_idx[atomicAdd(needy_count, 1)] = (int) i; return; }
    const int fc = hex_from_tet(T, p, tet);
    cell_inout[i] = fc < 0 ? -1 : hex2cell[fc];
}

__global__ void __launch_bounds__(256) k_particle_scanned(Tables T, const double* __restrict__ pts, const int* __restrict__ hex2cell,
                                                          const int* __restrict__ needy_count, const int* __restrict__ needy_idx,
                                                          int* __restrict__ cell_inout) {
    const int count = *needy_count;
    if (T.grid.n_buckets > 0) {         // grid filter: one warp per deferred particle
        for (long e = ((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < count; e += ((long) gridDim.x * blockDim.x) >> 5) {
            const long i = needy_idx[e];
            const P3 p = ldp(pts, i);
            const int tet = grid_scan(T, p);
            if ((threadIdx.x & 31) == 0) {
                const int fc = hex_from_tet(T, p, tet);
                cell_inout[i] = fc < 0 ? -1 : hex2cell[fc];
            }
        }
        return;
    }
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const long i = needy_idx[e];
        const P3 p = ldp(pts, i);
        const int tet = block_scan_all<TetFam>(T, p);
        if (threadIdx.x == 0) {
            const int fc = hex_from_tet(T, p, tet);
            cell_inout[i] = fc < 0 ? -1 : hex2cell[fc];
        }
    }
}

__global__ void __launch_bounds__(128) k_particle_field(Tables T, long n, const double* __restrict__ pts, const int* __restrict__ cell2hex, int n_cells,
                                                        const int* __restrict__ cells, double* __restrict__ E3) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cells[i];
    if (c < 0 || c >= n_cells) { E3[3 * i] = 0.0; E3[3 * i + 1] = 0.0; E3[3 * i + 2] = 0.0; return; }      // no cell, no field (never out of bounds)
    const P3 E = hex_point_gradient(T, ldp(pts, i), cell2hex[c]);
    E3[3 * i] = E.x; E3[3 * i + 1] = E.y; E3[3 * i + 2] = E.z;
}

// ---- PIC push --------------------------------------------------------
// Pic<3>::update_position (src/Pic.cpp:151-184) without the cell search: pos += vel dt (separate multiply and add,
// as the reference's Vec3 operators compile for baseline x86-64), periodic images (src/Macros.cpp:41-48) or the
// x/y box test, z < zmax; particles that left the box get cell = FB_PIC_LOST and are not searched.
#define FB_PIC_LOST (-2)
__device__ __forceinline__ double periodic_image(double p, double mx, double mn) {
    const double from_max = p - mx;
    if (from_max > 0) return mn + from_max;
    const double from_min = p - mn;
    if (from_min < 0) return mx + from_min;
    return p;
}
__global__ void __launch_bounds__(256) k_pic_move(long n, double* __restrict__ pos, const double* __restrict__ vel, int* __restrict__ cell,
                                                  double dt, double xmin, double xmax, double ymin, double ymax, double zmax, int periodic) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = pos[3 * i] + vel[3 * i] * dt, y = pos[3 * i + 1] + vel[3 * i + 1] * dt;
    const double z = pos[3 * i + 2] + vel[3 * i + 2] * dt;
    bool b1 = true, b2 = true;
    const bool b3 = z < zmax;
    if (periodic) { x = periodic_image(x, xmax, xmin); y = periodic_image(y, ymax, ymin); }
    else { b1 = x > xmin && x < xmax; b2 = y > ymin && y < ymax; }
    pos[3 * i] = x; pos[3 * i + 1] = y; pos[3 * i + 2] = z;
    if (!(b1 && b2 && b3)) cell[i] = FB_PIC_LOST;
}

// Pic<3>::update_velocities (src/Pic.cpp:198-209): vel += interp_gradient(pos, deal2femocs(cell)) * (dt q/m)
__global__ void __launch_bounds__(128) k_pic_velocities(Tables T, long n, const double* __restrict__ pts, const int* __restrict__ cell2hex, int n_cells,
                                                        const int* __restrict__ cells, double* __restrict__ vel, double dt_q_over_m) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cells[i];
    if (c < 0 || c >= n_cells) return;          // a particle without a valid cell is not accelerated (never out of bounds)
    const P3 E = hex_point_gradient(T, ldp(pts, i), cell2hex[c]);
    vel[3 * i] += E.x * dt_q_over_m; vel[3 * i + 1] += E.y * dt_q_over_m; vel[3 * i + 2] += E.z * dt_q_over_m;
}

// ParticleSpecies::clear_lost (src/ParticleSpecies.cpp:16-31): stable removal of the particles with cell == -1.
// Three passes: kept particles per block of 1024 -> exclusive scan of the block counts (one block) -> scatter.
constexpr int COMPACT_BLOCK = 1024;
__global__ void __launch_bounds__(256) k_compact_count(long n, const int* __restrict__ cell, int* __restrict__ block_count) {
    const long base = (long) blockIdx.x * COMPACT_BLOCK;
    int cnt = 0;
    for (int k = threadIdx.x; k < COMPACT_BLOCK; k += 256) { const long i = base + k; cnt += (i < n && cell[i] != -1); }
    __shared__ int s[8];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s[w]; block_count[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_compact_scan(int nb, int* __restrict__ block_count, long* __restrict__ total) {
    // exclusive scan of nb counts by one block (chunks of 1024 with a running carry)
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? block_count[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) block_count[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_compact_scatter(long n, const double* __restrict__ pos, const double* __restrict__ vel,
                                                         const int* __restrict__ cell, const int* __restrict__ block_offset,
                                                         double* __restrict__ pos_out, double* __restrict__ vel_out, int* __restrict__ cell_out) {
    // 4 rounds of 256 particles; inside a round the rank of a kept particle = kept in lower warps + lower lanes
    __shared__ int s_w[8];
    __shared__ int s_round;
    const long base = (long) blockIdx.x * COMPACT_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_round = block_offset[blockIdx.x];
    __syncthreads();
    for (int r = 0; r < COMPACT_BLOCK / 256; ++r) {
        const long i = base + r * 256 + threadIdx.x;
        const bool keep = i < n && cell[i] != -1;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int off = s_round;
        for (int w = 0; w < warp; ++w) off += s_w[w];
        if (keep) {
            const long o = off + __popc(m & ((1u << lane) - 1));
            pos_out[3 * o] = pos[3 * i]; pos_out[3 * o + 1] = pos[3 * i + 1]; pos_out[3 * o + 2] = pos[3 * i + 2];
            vel_out[3 * o] = vel[3 * i]; vel_out[3 * o + 1] = vel[3 * i + 1]; vel_out[3 * o + 2] = vel[3 * i + 2];
            cell_out[o] = cell[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_w[w]; s_round += t; }
        __syncthreads();
    }
}

// PoissonSolver<3>::assemble_space_charge_fast (PoissonSolver.cpp:299-319): 8 scatter-adds per particle.
// FROM_VERTS = false reads the LinearHexahedra coefficient table of the interpolator (192 B/hex);
// FROM_VERTS = true rebuilds f0..f7 (InterpolatorCells.cpp:1205-1245) from the 8 cell vertices, for
// solver-only meshes that have no tetrahedral tables (the refined benchmark meshes).
template <bool FROM_VERTS>
__global__ void __launch_bounds__(128) k_space_charge(const HexRec* __restrict__ hexrec, long n, const double* __restrict__ pts,
                                                      const int* __restrict__ pcell, int n_cells, const int* __restrict__ cell2hex,
                                                      const int* __restrict__ cells_dof, const double* __restrict__ vxyz,
                                                      double charge_factor, double* __restrict__ rhs,
                                                      const int* __restrict__ gcell2local, int n_cells_global, int n_rows,
                                                      const int* __restrict__ cells27) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cell = pcell[i];
    if (gcell2local) {          // partitioned mesh: particles carry global cell ids; cells of other ranks are skipped
        if (cell < 0 || cell >= n_cells_global) return;
        cell = gcell2local[cell];
    }
    if (cell < 0 || cell >= n_cells) return;
    double sf[8];
    if (FROM_VERTS) {
        const int lex[8] = {0, 1, 5, 4, 2, 3, 7, 6};   // femocs/UCD local vertex k sits at lexicographic position lex[k]
        P3 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = ldp(vxyz, cells_dof[8 * (long) cell + lex[k]]);
        HexRec H;
        const P3 f0 = (x[0] + x[1] + x[2] + x[3] + x[4] + x[5] + x[6] + x[7]) * 0.125;
        const P3 f1 = ((x[0] * -1.0) + x[1] + x[2] - x[3] - x[4] + x[5] + x[6] - x[7]) * 0.125;
        const P3 f2 = ((x[0] * -1.0) - x[1] + x[2] + x[3] - x[4] - x[5] + x[6] + x[7]) * 0.125;
        const P3 f3 = ((x[0] * -1.0) - x[1] - x[2] - x[3] + x[4] + x[5] + x[6] + x[7]) * 0.125;
        const P3 f4 = (x[0] - x[1] + x[2] - x[3] + x[4] - x[5] + x[6] - x[7]) * 0.125;
        const P3 f5 = (x[0] - x[1] - x[2] + x[3] - x[4] + x[5] + x[6] - x[7]) * 0.125;
        const P3 f6 = (x[0] + x[1] - x[2] - x[3] - x[4] - x[5] + x[6] + x[7]) * 0.125;
        const P3 f7 = ((x[0] * -1.0) + x[1] - x[2] + x[3] + x[4] - x[5] + x[6] - x[7]) * 0.125;
        const P3 f[8] = {f0, f1, f2, f3, f4, f5, f6, f7};
#pragma unroll
        for (int k = 0; k < 8; ++k) { H.f[k][0] = f[k].x; H.f[k][1] = f[k].y; H.f[k][2] = f[k].z; }
        hex_sf(H, ldp(pts, i), sf);
    } else {
        hex_sf(hexrec[cell2hex[cell]], ldp(pts, i), sf);
    }
    const int perm[8] = {0, 1, 4, 5, 3, 2, 7, 6};       // shape_funs_dealii (:1355-1358)
    if (cells27) {
        // FE_Q(2) (PoissonSolver.cpp:276-296 -> DealSolver::shape_funs, DealSolver.cpp:75-110): the 27 shape values at the
        // particle's unit-cell point, clamped to the cell (project_to_unit_cell); the point follows from the trilinear
        // weights: xi_d = sum of the weights of the vertices with bit d set
        double xi[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double w = sf[perm[k]];
            if (k & 1) xi[0] += w;
            if (k & 2) xi[1] += w;
            if (k & 4) xi[2] += w;
        }
        double L[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double x = fmin(1.0, fmax(0.0, xi[d]));
            L[d][0] = 2.0 * (x - 0.5) * (x - 1.0); L[d][1] = 4.0 * x * (1.0 - x); L[d][2] = 2.0 * x * (x - 0.5);
        }
        for (int a = 0; a < 27; ++a)
            atomicAdd(&rhs[cells27[27 * (long) cell + a]], L[0][a % 3] * L[1][(a / 3) % 3] * L[2][a / 9] * charge_factor);
        return;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dof = cells_dof[8 * (long) cell + k];
        if (dof < n_rows) atomicAdd(&rhs[dof], sf[perm[k]] * charge_factor);       // rows of ghost dofs belong to their owner
    }
}

__global__ void k_pack_points(long n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                              int stride, double* __restrict__ out) {
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * i] = x[i * stride]; out[3 * i + 1] = y[i * stride]; out[3 * i + 2] = z[i * stride];
}

// ---------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------
static Tables make_tables(const fb_ctx* c) {
    Tables T;
    T.nxyz = c->d_nxyz.p; T.nodal = c->d_nodal.p; T.hex8 = c->d_hex8.p;
    T.tet = c->d_tet.p; T.tet_cent = c->d_tet_cent.p; T.tet_mark = c->d_tet_mark.p; T.tet_nbr_off = c->d_tet_nbr_off.p;
    T.tet_nbr = c->d_tet_nbr.p; T.tet4 = c->d_tet4.p; T.n_tet = c->n_tet;
    T.tri = c->d_tri.p; T.tri_cent = c->d_tri_cent.p; T.tri_nbr_off = c->d_tri_nbr_off.p; T.tri_nbr = c->d_tri_nbr.p;
    T.tri2tet = c->d_tri2tet.p; T.n_tri = c->n_tri;
    T.hex = c->d_hex.p; T.quad2hex = c->d_quad2hex.p; T.qtet = c->d_qtet.p; T.qtri = c->d_qtri.p;
    CellGridDev& G = T.grid;
    G.n_buckets = c->grid_on ? c->grid_g[0] * c->grid_g[1] * c->grid_g[2] : 0;
    for (int d = 0; d < 3; ++d) { G.lo[d] = c->grid_lo[d]; G.h[d] = c->grid_h[d]; G.inv_h[d] = 1.0 / c->grid_h[d]; G.g[d] = c->grid_g[d]; }
    G.off = c->d_grid_off.p; G.list = c->d_grid_list.p; G.coff = c->d_grid_coff.p; G.clist = c->d_grid_clist.p;
    return T;
}

// Uniform grid over the tetrahedra, built on the device (two passes over the bucket ranges: count, fill); the host only
// fixes the box (all mesh nodes) and the bucket edge (about 4 buckets per tetrahedron, at most 2^21 buckets).
int launch_build_cell_grid(fb_ctx* c, const double* bb_lo, const double* bb_hi) {
    c->grid_on = false;
    if (c->n_tet < 64 || c->cell_grid == 0) return FB_OK;           // tiny meshes: the plain scan is as fast
    double ext[3], vol = 1;
    for (int d = 0; d < 3; ++d) { ext[d] = std::max(bb_hi[d] - bb_lo[d], 1e-9); vol *= ext[d]; }
    double h = std::cbrt(vol / (4.0 * c->n_tet));
    long nb = 1;
    for (int it = 0; it < 8; ++it) {
        nb = 1;
        for (int d = 0; d < 3; ++d) { c->grid_g[d] = (int) std::min(512.0, std::max(1.0, std::ceil(ext[d] / h))); nb *= c->grid_g[d]; }
        if (nb <= (1L << 21)) break;
        h *= 1.3;
    }
    const double margin = 1e-6 * std::max(ext[0], std::max(ext[1], ext[2]));
    for (int d = 0; d < 3; ++d) { c->grid_lo[d] = bb_lo[d] - 2 * margin; c->grid_h[d] = (ext[d] + 4 * margin) / c->grid_g[d]; }
    cudaStream_t s = c->stream;
    FB_CUDA(c, c->d_grid_cnt.alloc(2 * (size_t) nb + 4));
    int* cnt = c->d_grid_cnt.p; int* ccnt = cnt + nb; int* totals = ccnt + nb;
    FB_CUDA(c, c->d_grid_off.alloc(nb + 1)); FB_CUDA(c, c->d_grid_coff.alloc(nb + 1));
    FB_CUDA(c, c->d_grid_clist.alloc(c->n_tet));
    FB_CUDA(c, c->d_grid_list.alloc(std::max<size_t>(c->d_grid_list.n, 48 * (size_t) c->n_tet)));      // capacity; checked below
    const int g = (c->n_tet + 127) / 128;
    int h_tot[2] = {0, 0};
    for (int attempt = 0; attempt < 2; ++attempt) {
        FB_CUDA(c, cudaMemsetAsync(cnt, 0, (2 * (size_t) nb + 4) * sizeof(int), s));
        c->grid_on = true;
        Tables T = make_tables(c);
        c->grid_on = false;
        const int cap = (int) std::min<size_t>(c->d_grid_list.n, 0x7fffffff);
        k_grid_build<false><<<g, 128, 0, s>>>(T, T.grid, margin, cnt, ccnt, nullptr, nullptr, 0);
        k_grid_scan<<<1, 1024, 0, s>>>((int) nb, cnt, c->d_grid_off.p, ccnt, c->d_grid_coff.p, totals);
        k_grid_build<true><<<g, 128, 0, s>>>(T, T.grid, margin, cnt, ccnt, c->d_grid_list.p, c->d_grid_clist.p, cap);
        c->launches += 3;
        FB_CUDA(c, cudaMemcpyAsync(h_tot, totals, sizeof h_tot, cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
        if (h_tot[0] <= cap) break;
        FB_CUDA(c, c->d_grid_list.alloc((size_t) h_tot[0]));        // graded mesh with huge far-field cells: exact size, once more
    }
    c->grid_on = true; c->grid_entries = h_tot[0];
    if (getenv("FB_VERBOSE")) fprintf(stderr, "[fb] cell grid %d x %d x %d, %d list entries for %d tetrahedra\n", c->grid_g[0], c->grid_g[1], c->grid_g[2], h_tot[0], c->n_tet);
    return FB_OK;
}

void launch_extract(fb_ctx* c, int smoothen) {
    const int g = (c->n_nodes + 127) / 128;
    k_store_solution<<<(c->n_nodes + 255) / 256, 256, 0, c->stream>>>(c->n_nodes, c->d_node2vert.p, c->d_vertex2dof.p, c->d_x.p, c->rho_valid ? c->d_rho.p : nullptr,
                                                                                c->d_nodal.p);
    k_nodal_field<<<g, 128, 0, c->stream>>>(c->n_nodes, c->d_node2vert.p, c->d_n2c_off.p, c->d_n2c_list.p, c->d_nxyz.p, c->d_hex8.p,
                                            c->d_nodal.p);
    c->launches += 2;
    if (smoothen && c->n_voro > 0) {
        k_smooth<<<(c->n_voro + 3) / 4, 128, 0, c->stream>>>(c->n_voro, c->d_voro_off.p, c->d_voro_list.p, c->d_nxyz.p, c->decay_factor, c->d_nodal.p);
        c->launches++;
    }
}

void launch_pack_points(fb_ctx* c, long n, const double* x, const double* y, const double* z, int stride, double* out) {
    k_pack_points<<<(unsigned) ((n + 255) / 256), 256, 0, c->stream>>>(n, x, y, z, stride, out);
    c->launches++;
}

// chained-guess location for n points already on the device: brute-force guess-free scan, then the
// fix-point of the guess chain in one cooperative launch.  Result (base-family cells) in c->d_scan2.
template <class Fam>
static int chain_fixpoint(fb_ctx* c, const Tables& T, long n, const double* d_pts, int first_guess, long chain_len) {
    auto kern = k_chain_fixpoint<Fam>;
    if (c->chain_blocks_per_sm == 0) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, 0);
        c->chain_blocks_per_sm = std::max(1, nb);
    }
    const long want = (n + 127) / 128;
    const int grid = (int) std::min<long>(want, (long) c->chain_blocks_per_sm * c->n_sm);
    const double* a2 = d_pts; const int* a3 = c->d_scan.p; int* a4 = c->d_cellsA.p; int* a5 = c->d_cellsB.p;
    unsigned char* a6 = c->d_dirtyA.p; unsigned char* a7 = c->d_dirtyB.p; int a8 = first_guess; int* a9 = c->d_flag.p; int* a10 = c->d_scan2.p;
    Tables t = T; long nn = n; long cl = chain_len;
    void* args[] = {&t, &nn, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9, &a10, &cl};
    cudaError_t e = cudaLaunchCooperativeKernel((void*) kern, dim3(grid), dim3(128), args, 0, c->stream);
    if (e != cudaSuccess) return c->fail(FB_ERR_CUDA, "locate chain launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return FB_OK;
}

int launch_locate_chain(fb_ctx* c, int dim, int rank, long n, const double* d_pts, int** result, long chain_len) {
    const Tables T = make_tables(c);
    // abs(-1) = 1 is the first guess of the reference loop (SolutionReader.cpp:145, InterpolatorCells.cpp:443);
    // hex/quad ranks divide it by 4/3 (:1539, :1955)
    const int first_guess = (rank == 3) ? 0 : 1;
    const unsigned gs = (unsigned) ((n + 31) / 32);          // 32 points per block of 8 warps
    if (dim == 2) k_scan_cells<TriFam, 128><<<gs, 256, 0, c->stream>>>(T, n, d_pts, c->d_scan.p);
    else if (T.grid.n_buckets > 0) k_scan_grid<<<(unsigned) ((n + 3) / 4), 128, 0, c->stream>>>(T, n, d_pts, c->d_scan.p);
    else k_scan_cells<TetFam, 128><<<gs, 256, 0, c->stream>>>(T, n, d_pts, c->d_scan.p);
    c->launches++;
    cudaMemsetAsync(c->d_flag.p, 0, 4 * sizeof(int), c->stream);
    const int rc = (dim == 2) ? chain_fixpoint<TriFam>(c, T, n, d_pts, first_guess, chain_len)
                              : chain_fixpoint<TetFam>(c, T, n, d_pts, first_guess, chain_len);
    if (rc) return rc;
    *result = c->d_scan2.p;
    return FB_OK;
}

void launch_finish_interp(fb_ctx* c, int dim, int rank, long n, const double* d_pts, const int* d_base, int final_cells,
                          int* d_cells_out, double* d_sol) {
    const Tables T = make_tables(c);
    if (dim == 2) { c->d_needy.alloc((size_t) n + 1); cudaMemsetAsync(c->d_needy.p, 0, sizeof(int), c->stream); }
    k_finish_interp<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(T, dim, rank, n, d_pts, d_base, final_cells, d_cells_out, d_sol,
                                                                        c->d_needy.p, c->d_needy.p + 1);
    c->launches++;
    if (dim == 2) {     // deferred points (device-side count): one block each, grid-stride
        k_finish_scanned<<<(unsigned) std::min<long>(n, 4L * c->n_sm), 256, 0, c->stream>>>(T, rank, d_pts, c->d_needy.p, c->d_needy.p + 1, d_sol);
        c->launches++;
    }
}

void launch_particle_cells(fb_ctx* c, long n, const double* d_pts, int* d_cells, bool after_move) {
    const Tables T = make_tables(c);
    c->d_needy.alloc((size_t) n + 1);
    cudaMemsetAsync(c->d_needy.p, 0, sizeof(int), c->stream);
    k_particle_cells<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(T, n, d_pts, c->d_cell2hex.p, c->n_cells, c->d_hex2cell.p, d_cells,
                                                                         c->d_needy.p, c->d_needy.p + 1, after_move ? FB_PIC_LOST : (int) 0x80000000);
    k_particle_scanned<<<(unsigned) std::min<long>(n, 8L * c->n_sm), 256, 0, c->stream>>>(T, d_pts, c->d_hex2cell.p, c->d_needy.p, c->d_needy.p + 1, d_cells);
    c->launches += 2;
}

void launch_particle_field(fb_ctx* c, long n, const double* d_pts, const int* d_cells, double* d_E) {
    const Tables T = make_tables(c);
    k_particle_field<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(T, n, d_pts, c->d_cell2hex.p, c->n_cells, d_cells, d_E);
    c->launches++;
}

void launch_pic_move(fb_ctx* c, long n, double* d_pos, const double* d_vel, int* d_cells, double dt, const double* box6, int periodic) {
    k_pic_move<<<(unsigned) ((n + 255) / 256), 256, 0, c->stream>>>(n, d_pos, d_vel, d_cells, dt, box6[0], box6[1], box6[2], box6[3], box6[5], periodic);
    c->launches++;
}

void launch_pic_velocities(fb_ctx* c, long n, const double* d_pos, const int* d_cells, double* d_vel, double dt_q_over_m) {
    const Tables T = make_tables(c);
    k_pic_velocities<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(T, n, d_pos, c->d_cell2hex.p, c->n_cells, d_cells, d_vel, dt_q_over_m);
    c->launches++;
}

// stable compaction into (pos_out, vel_out, cell_out); the number of kept particles lands in *d_total (device)
void launch_pic_compact(fb_ctx* c, long n, const double* d_pos, const double* d_vel, const int* d_cells, int* d_block_count,
                        long* d_total, double* pos_out, double* vel_out, int* cell_out) {
    const int nb = (int) ((n + COMPACT_BLOCK - 1) / COMPACT_BLOCK);
    k_compact_count<<<nb, 256, 0, c->stream>>>(n, d_cells, d_block_count);
    k_compact_scan<<<1, 1024, 0, c->stream>>>(nb, d_block_count, d_total);
    k_compact_scatter<<<nb, 256, 0, c->stream>>>(n, d_pos, d_vel, d_cells, d_block_count, pos_out, vel_out, cell_out);
    c->launches += 3;
}

void launch_space_charge(fb_ctx* c, long n, const double* d_pts, const int* d_pcell, double charge_factor) {
    if (n <= 0) return;
    const unsigned g = (unsigned) ((n + 127) / 128);
    const int* q27 = c->imported_degree == 2 ? c->d_cells27.p : nullptr;
    if (c->interp_ok)
        k_space_charge<false><<<g, 128, 0, c->stream>>>(c->d_hex.p, n, d_pts, d_pcell, c->n_cells, c->d_cell2hex.p, c->d_cells.p,
                                                         c->d_vxyz.p, charge_factor, c->d_rhs.p, nullptr, 0, c->n_dofs, q27);
    else
        k_space_charge<true><<<g, 128, 0, c->stream>>>(nullptr, n, d_pts, d_pcell, c->n_cells, c->d_cell2hex.p, c->d_cells.p,
                                                        c->d_vxyz.p, charge_factor, c->d_rhs.p,
                                                        c->world > 1 ? c->d_gcell2local.p : nullptr, c->n_cells_global, c->n_dofs, q27);
    c->launches++;
}

}  // namespace fb
