// CUDA kernels (sm_100a) of the Laplace/Poisson solver: Q1 stiffness assembly, Neumann face
// RHS, Dirichlet elimination, and the fused Jacobi-preconditioned CG.
//
// Reference behaviour replaced (paths relative to the reference root):
//   src/PoissonSolver.cpp:213-263  assemble_parallel / assemble_local_cell (FE_Q(1), QGauss<3>(2))
//   src/DealSolver.cpp:389-430     assemble_rhs (Neumann faces, QGauss<2>(2))
//   src/DealSolver.cpp:432-440     append_dirichlet / apply_dirichlet (MatrixTools::apply_boundary_values)
//   src/DealSolver.cpp:442-458     solve_cg (deal.II SolverCG, absolute residual tolerance)
//   src/DealSolver.cpp:157-167     check_limits
//
// Roofline: everything here is HBM-bound FP64 streaming (<= 0.25 flop/byte); tensor cores
// are not used.  Layout: CSR (double val, int col) with sorted columns, vectors in DoF order.
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace fb {

// ---------------------------------------------------------------------------------------
// block reduction + deterministic cross-block reduction ("last block reduces the partials")
// ---------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ bool reduce_publish(double (&v)[NV], double* __restrict__ partial, unsigned* counter,
                                               double (&total)[NV]) {
    __shared__ double sm[NV][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0;
            for (int w = 0; w < nwarp; ++w) s += sm[k][w];
            partial[(size_t) k * gridDim.x + blockIdx.x] = s;
        }
        __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    // last block: fixed-order reduction of all block partials (deterministic for a given grid)
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) x += __ldcg(&partial[(size_t) k * gridDim.x + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if (lane == 0) sm[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0;
            for (int w = 0; w < nwarp; ++w) s += sm[k][w];
            total[k] = s;
        }
        *counter = 0;
    }
    return threadIdx.x == 0;
}

// ---------------------------------------------------------------------------------------
// Q1 stiffness assembly: one thread per hexahedron, 2x2x2 Gauss, MappingQ1.
//   K_e(i,j) = sum_q JxW_q grad N_i(q) . grad N_j(q)       (PoissonSolver.cpp:249-255)
// scattered into the CSR matrix with FP64 atomics (summation order across cells is not fixed:
// ~1e-16 relative run-to-run variation, far below the 1e-8 parity bar).
// Algorithmic bytes per hex: 32 B connectivity + 192 B coordinates + 64 x 8 B adds.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int find_col(const int* __restrict__ col, int lo, int hi, int c) {
    while (lo < hi) {                       // columns are sorted within a row
        const int mid = (lo + hi) >> 1;
        if (__ldg(&col[mid]) < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128) k_assemble_stiffness(int n_cells, const int* __restrict__ cells,
                                                           const double* __restrict__ vxyz,
                                                           const int* __restrict__ rowptr, const int* __restrict__ col,
                                                           double* __restrict__ val, double* __restrict__ cell_vol) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int dof[8];
    double X[8], Y[8], Z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dof[i] = __ldg(&cells[8 * (size_t) c + i]);
        X[i] = __ldg(&vxyz[3 * (size_t) dof[i]]); Y[i] = __ldg(&vxyz[3 * (size_t) dof[i] + 1]); Z[i] = __ldg(&vxyz[3 * (size_t) dof[i] + 2]);
    }
    double Ke[36];      // upper triangle, row-major: (i,j), j >= i
#pragma unroll
    for (int k = 0; k < 36; ++k) Ke[k] = 0;
    const double ga = 0.5 * (1.0 - 0.57735026918962576451), gb = 0.5 * (1.0 + 0.57735026918962576451);
    double vol = 0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const double xi = (q & 1) ? gb : ga, eta = (q & 2) ? gb : ga, zeta = (q & 4) ? gb : ga;
        double dN[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double fx = (i & 1) ? xi : 1.0 - xi, fy = (i & 2) ? eta : 1.0 - eta, fz = (i & 4) ? zeta : 1.0 - zeta;
            const double sx = (i & 1) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0, sz = (i & 4) ? 1.0 : -1.0;
            dN[i][0] = sx * fy * fz; dN[i][1] = fx * sy * fz; dN[i][2] = fx * fy * sz;
        }
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 3; ++e) { J[0][e] += X[i] * dN[i][e]; J[1][e] += Y[i] * dN[i][e]; J[2][e] += Z[i] * dN[i][e]; }
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][0] * J[2][2] - J[1][2] * J[2][0];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02;
        const double id = 1.0 / det;
        double inv[3][3];
        inv[0][0] = c00 * id;
        inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
        inv[1][0] = -c01 * id;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
        inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
        inv[2][0] = c02 * id;
        inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
        const double JxW = det * 0.125;
        vol += JxW;
        double G[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int d = 0; d < 3; ++d) G[i][d] = dN[i][0] * inv[0][d] + dN[i][1] * inv[1][d] + dN[i][2] * inv[2][d];
        int k = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = i; j < 8; ++j, ++k) Ke[k] += JxW * (G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2]);
    }
    if (cell_vol) cell_vol[c] = vol;
    int k = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ri = dof[i], lo_i = __ldg(&rowptr[ri]), hi_i = __ldg(&rowptr[ri + 1]);
#pragma unroll
        for (int j = i; j < 8; ++j, ++k) {
            atomicAdd(&val[find_col(col, lo_i, hi_i, dof[j])], Ke[k]);
            if (j != i) {
                const int rj = dof[j];
                atomicAdd(&val[find_col(col, __ldg(&rowptr[rj]), __ldg(&rowptr[rj + 1]), ri)], Ke[k]);
            }
        }
    }
}

// Neumann faces (DealSolver.cpp:389-430): b_i += sum_q N_i(q) * bc * JxW_face(q), QGauss<2>(2)
__global__ void k_neumann_faces(int n_faces, const int* __restrict__ face_dofs, const double* __restrict__ vxyz,
                                double bc_value, double* __restrict__ rhs) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    int d[4]; double P[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        d[i] = face_dofs[4 * f + i];
#pragma unroll
        for (int e = 0; e < 3; ++e) P[i][e] = vxyz[3 * (size_t) d[i] + e];
    }
    const double ga = 0.5 * (1.0 - 0.57735026918962576451), gb = 0.5 * (1.0 + 0.57735026918962576451);
    double r[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double s = (q & 1) ? gb : ga, t = (q & 2) ? gb : ga;
        double ds[3], dt[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            ds[e] = (P[1][e] - P[0][e]) * (1 - t) + (P[3][e] - P[2][e]) * t;
            dt[e] = (P[2][e] - P[0][e]) * (1 - s) + (P[3][e] - P[1][e]) * s;
        }
        const double nx = ds[1] * dt[2] - ds[2] * dt[1], ny = ds[2] * dt[0] - ds[0] * dt[2], nz = ds[0] * dt[1] - ds[1] * dt[0];
        const double JxW = sqrt(nx * nx + ny * ny + nz * nz) * 0.25;
        r[0] += (1 - s) * (1 - t) * bc_value * JxW; r[1] += s * (1 - t) * bc_value * JxW;
        r[2] += (1 - s) * t * bc_value * JxW;       r[3] += s * t * bc_value * JxW;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(&rhs[d[i]], r[i]);
}

// mark constrained dofs
__global__ void k_set_bc(int n, const int* __restrict__ dofs, double value, int* __restrict__ flag, double* __restrict__ bcval) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { flag[dofs[i]] = 1; bcval[dofs[i]] = value; }
}

// Symmetric Dirichlet elimination (MatrixTools::apply_boundary_values, eliminate_columns=true):
// val <- val_save with constrained rows/columns zeroed except the diagonal; lift_r = sum_c a_rc v_c;
// dinv = 1/diag.  LANES threads cooperate on one row.
template <int LANES>
__global__ void __launch_bounds__(256) k_apply_bc_matrix(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                         const double* __restrict__ val_save, double* __restrict__ val,
                                                         const int* __restrict__ flag, const double* __restrict__ bcval,
                                                         double* __restrict__ lift, double* __restrict__ dinv,
                                                         int* __restrict__ diagpos) {
    // the row loop is warp-uniform (all 32 lanes take the same number of trips) so that the
    // full-mask shuffles below are always executed by the whole warp
    const int lane = threadIdx.x % LANES;
    const long gtid = (long) blockIdx.x * blockDim.x + threadIdx.x;
    const long warp_row0 = (gtid / 32) * (32 / LANES);
    const long stride = (long) gridDim.x * blockDim.x / LANES;
    for (long rb = warp_row0; rb < n; rb += stride) {
        const long r = rb + (threadIdx.x % 32) / LANES;
        const bool valid = r < n;
        const int lo = valid ? rowptr[r] : 0, hi = valid ? rowptr[r + 1] : 0;
        const bool rc = valid && flag[r] != 0;
        double l = 0;
        for (int k = lo + lane; k < hi; k += LANES) {
            const int c = col[k];
            const double a = val_save[k];
            double out;
            if (c == (int) r) { out = a; dinv[r] = 1.0 / a; diagpos[r] = k; }
            else if (rc) out = 0;
            else if (flag[c]) { l += a * bcval[c]; out = 0; }
            else out = a;
            val[k] = out;
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o, LANES);
        if (lane == 0 && valid) lift[r] = l;
    }
}

// rhs finalisation: b_c = a_cc v_c on constrained rows, b_r -= lift_r elsewhere; x_c = v_c
__global__ void k_apply_bc_rhs(int n, const int* __restrict__ flag, const double* __restrict__ bcval,
                               const double* __restrict__ dinv, const double* __restrict__ lift,
                               double* __restrict__ rhs, double* __restrict__ x) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        if (flag[i]) { rhs[i] = bcval[i] / dinv[i]; x[i] = bcval[i]; }
        else rhs[i] -= lift[i];
    }
}

// ---------------------------------------------------------------------------------------
// Fused CG.  Per iteration (deal.II SolverCG order of operations):
//   k_spmv_dot : h = A d ; alpha = gh / (d.h)                       [12 nnz + 4n + 16n B]
//   k_update   : x += alpha d ; g += alpha h ; res = |g| ; gh' = g.(Dinv g); beta = gh'/gh
//                convergence test (absolute, deal.II SolverControl)  [56n B]
//   k_direction: d = beta d - Dinv g                                 [32n B]
// All scalars live in device memory (CgScalars); no host round trip inside the loop.
// ---------------------------------------------------------------------------------------
template <int LANES, bool INIT>
__global__ void __launch_bounds__(256) k_spmv_dot(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                  const double* __restrict__ val, const double* __restrict__ xin,
                                                  const double* __restrict__ rhs, const double* __restrict__ dinv,
                                                  double* __restrict__ out, double* __restrict__ partial, unsigned* counter,
                                                  CgScalars* __restrict__ cgs, double* __restrict__ alpha_out) {
    if (!INIT && cgs->done) return;
    const int lane = threadIdx.x % LANES;
    const long gtid = (long) blockIdx.x * blockDim.x + threadIdx.x;
    const long warp_row0 = (gtid / 32) * (32 / LANES);       // warp-uniform trip count (full-mask shuffles)
    const long stride = (long) gridDim.x * blockDim.x / LANES;
    double acc[2] = {0, 0};
    for (long rb = warp_row0; rb < n; rb += stride) {
        const long r = rb + (threadIdx.x % 32) / LANES;
        const bool valid = r < n;
        const int lo = valid ? __ldg(&rowptr[r]) : 0, hi = valid ? __ldg(&rowptr[r + 1]) : 0;
        double s = 0;
        for (int k = lo + lane; k < hi; k += LANES) s += __ldg(&val[k]) * xin[__ldg(&col[k])];
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LANES);
        if (lane == 0 && valid) {
            if (INIT) {                      // g = A x - b ; accumulate g.(Dinv g) and g.g
                const double g = s - rhs[r];
                out[r] = g;
                acc[0] += g * g * dinv[r];
                acc[1] += g * g;
            } else {                         // h = A d ; accumulate d.h
                out[r] = s;
                acc[0] += xin[r] * s;
            }
        }
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        if (INIT) {
            cgs->gh = tot[0]; cgs->res2 = tot[1]; cgs->it = 0;
            cgs->done = (tot[1] <= cgs->tol2) ? 1 : ((cgs->max_iter <= 0 || tot[1] != tot[1]) ? 2 : 0);
        } else {
            *alpha_out = cgs->gh / tot[0];
        }
    }
}

__global__ void __launch_bounds__(256) k_update(int n, const double* __restrict__ d, const double* __restrict__ h,
                                                const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ g,
                                                double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                const double* __restrict__ alpha_in, double* __restrict__ beta_out) {
    if (cgs->done) return;
    const double alpha = *alpha_in;
    double acc[2] = {0, 0};
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        x[i] += alpha * d[i];
        const double gi = g[i] + alpha * h[i];
        g[i] = gi;
        acc[0] += gi * gi * dinv[i];
        acc[1] += gi * gi;
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        const int it = cgs->it + 1;
        cgs->it = it;
        cgs->res2 = tot[1];
        *beta_out = tot[0] / cgs->gh;
        cgs->gh = tot[0];
        if (tot[1] <= cgs->tol2) cgs->done = 1;                     // SolverControl: success first,
        else if (it >= cgs->max_iter || tot[1] != tot[1]) cgs->done = 2;   // then failure on max steps / NaN
    }
}

template <bool INIT>
__global__ void __launch_bounds__(256) k_direction(int n, const double* __restrict__ g, const double* __restrict__ dinv,
                                                   double* __restrict__ d, const CgScalars* __restrict__ cgs,
                                                   const double* __restrict__ beta_in) {
    if (cgs->done) return;
    const double beta = INIT ? 0.0 : *beta_in;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x)
        d[i] = (INIT ? 0.0 : beta * d[i]) - dinv[i] * g[i];
}

// min/max of the solution (DealSolver::check_limits)
__global__ void __launch_bounds__(256) k_minmax(int n, const double* __restrict__ x, double* __restrict__ partial,
                                                unsigned* counter, double* __restrict__ out2) {
    double mn = 1e100, mx = -1e100;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        mn = fmin(mn, x[i]); mx = fmax(mx, x[i]);
    }
    __shared__ double smn[32], smx[32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarp; ++w) { mn = fmin(mn, smn[w]); mx = fmax(mx, smx[w]); }
        partial[blockIdx.x] = mn; partial[gridDim.x + blockIdx.x] = mx;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        mn = 1e100; mx = -1e100;
        for (unsigned b = 0; b < gridDim.x; ++b) { mn = fmin(mn, __ldcg(&partial[b])); mx = fmax(mx, __ldcg(&partial[gridDim.x + b])); }
        out2[0] = mn; out2[1] = mx; *counter = 0;
    }
}

__global__ void k_gather(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void k_scatter(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) dst[idx[i]] = src[i];
}

// ---------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------
static inline int grid_for(const fb_ctx* c, long work_items, int block) {
    long g = (work_items + block - 1) / block;
    const long cap = (long) c->n_sm * (2048 / block);
    if (g > cap) g = cap;
    return (int) (g < 1 ? 1 : g);
}

int choose_lanes(const fb_ctx* c) {
    const double avg = c->n_dofs ? (double) c->nnz / c->n_dofs : 1.0;
    if (avg > 48) return 32;
    if (avg > 20) return 8;
    if (avg > 10) return 4;
    return 2;
}

void launch_assemble_stiffness(fb_ctx* c, double* d_cell_vol) {
    const int block = 128;
    k_assemble_stiffness<<<(c->n_cells + block - 1) / block, block, 0, c->stream>>>(
        c->n_cells, c->d_cells.p, c->d_vxyz.p, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, d_cell_vol);
    c->launches++;
}

void launch_neumann(fb_ctx* c) {
    if (c->n_top_faces == 0) return;
    k_neumann_faces<<<(c->n_top_faces + 127) / 128, 128, 0, c->stream>>>(c->n_top_faces, c->d_topfaces.p, c->d_vxyz.p,
                                                                       c->applied_field, c->d_rhs.p);
    c->launches++;
}

void launch_set_bc(fb_ctx* c, const int* d_dofs, int n, double value) {
    if (n == 0) return;
    k_set_bc<<<(n + 255) / 256, 256, 0, c->stream>>>(n, d_dofs, value, c->d_bcflag.p, c->d_bcval.p);
    c->launches++;
}

void launch_apply_bc_matrix(fb_ctx* c) {
    const int g = grid_for(c, (long) c->n_dofs * 8, 256);
    k_apply_bc_matrix<8><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, c->d_val.p,
                                                   c->d_bcflag.p, c->d_bcval.p, c->d_w.p, c->d_dinv.p, c->d_diagpos.p);
    c->launches++;
}

void launch_apply_bc_rhs(fb_ctx* c) {
    const int g = grid_for(c, c->n_dofs, 256);
    k_apply_bc_rhs<<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_bcflag.p, c->d_bcval.p, c->d_dinv.p, c->d_w.p, c->d_rhs.p, c->d_x.p);
    c->launches++;
}

template <bool INIT>
static void spmv_dispatch(fb_ctx* c, int lanes, const double* xin, double* out, double* alpha) {
    const int g = grid_for(c, (long) c->n_dofs * lanes, 256);
    unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
    double* part = c->d_partial.p;
#define FB_SPMV(L) k_spmv_dot<L, INIT><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_rowptr.p, c->d_col.p, c->d_val.p, xin, \
                                                              c->d_rhs.p, c->d_dinv.p, out, part, counter, c->d_cg.p, alpha)
    switch (lanes) {
        case 32: FB_SPMV(32); break;
        case 16: FB_SPMV(16); break;
        case 8: FB_SPMV(8); break;
        case 4: FB_SPMV(4); break;
        default: FB_SPMV(2); break;
    }
#undef FB_SPMV
    c->launches++;
}

// scalars alpha/beta live behind the CgScalars struct in the same allocation (d_cg has 1 + 4 slots)
static inline double* alpha_ptr(fb_ctx* c) { return (double*) (c->d_cg.p + 1); }
static inline double* beta_ptr(fb_ctx* c) { return (double*) (c->d_cg.p + 1) + 1; }

void launch_cg_init(fb_ctx* c, int lanes) {
    spmv_dispatch<true>(c, lanes, c->d_x.p, c->d_g.p, nullptr);
    const int g = grid_for(c, c->n_dofs, 256);
    k_direction<true><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_d.p, c->d_cg.p, nullptr);
    c->launches++;
}

void launch_cg_spmv(fb_ctx* c, int lanes) {       // h = A d, alpha = gh / (d.h)
    spmv_dispatch<false>(c, lanes, c->d_d.p, c->d_h.p, alpha_ptr(c));
}

void launch_cg_vectors(fb_ctx* c) {               // x, g update + dots + convergence test, then new direction
    unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
    const int g = grid_for(c, c->n_dofs, 256);
    k_update<<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter,
                                       c->d_cg.p, alpha_ptr(c), beta_ptr(c));
    k_direction<false><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_d.p, c->d_cg.p, beta_ptr(c));
    c->launches += 2;
}

void launch_cg_iteration(fb_ctx* c, int lanes) {
    launch_cg_spmv(c, lanes);
    launch_cg_vectors(c);
}

void launch_minmax(fb_ctx* c) {
    const int g = grid_for(c, c->n_dofs, 256);
    unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
    k_minmax<<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_x.p, c->d_partial.p, counter, c->d_minmax.p);
    c->launches++;
}

void launch_gather(fb_ctx* c, int n, const int* idx, const double* src, double* dst) {
    k_gather<<<grid_for(c, n, 256), 256, 0, c->stream>>>(n, idx, src, dst);
    c->launches++;
}
void launch_scatter(fb_ctx* c, int n, const int* idx, const double* src, double* dst) {
    k_scatter<<<grid_for(c, n, 256), 256, 0, c->stream>>>(n, idx, src, dst);
    c->launches++;
}

}  // namespace fb
