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

#include <algorithm>
#include <vector>

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
// scalar tail of the two reduction kernels of a CG iteration (executed by ONE thread).  On one GPU the
// last block calls them with the grid totals; with a partitioned mesh the last block only deposits the
// rank-local sums in cgs->red, NCCL all-reduces them and k_cg_scalars_* calls the same code.
// ---------------------------------------------------------------------------------------
template <bool INIT>
__device__ __forceinline__ void cg_scalars_spmv(CgScalars* cgs, const double* tot, double* alpha_out) {
    if (INIT) {
        cgs->gh = tot[0]; cgs->res2 = tot[1]; cgs->it = 0;
        cgs->done = (tot[1] <= cgs->tol2) ? 1 : ((cgs->max_iter <= 0 || tot[1] != tot[1]) ? 2 : 0);
    } else {
        *alpha_out = cgs->gh / tot[0];
    }
}
__device__ __forceinline__ void cg_scalars_update(CgScalars* cgs, const double* tot, double* beta_out) {
    const int it = cgs->it + 1;
    cgs->it = it;
    cgs->res2 = tot[1];
    *beta_out = tot[0] / cgs->gh;
    cgs->gh = tot[0];
    if (tot[1] <= cgs->tol2) cgs->done = 1;                     // SolverControl: success first,
    else if (it >= cgs->max_iter || tot[1] != tot[1]) cgs->done = 2;   // then failure on max steps / NaN
}
// Chebyshev-preconditioned CG: the update kernel only knows |g| (iteration count, convergence); g.z -- and with it
// beta -- comes from the last Chebyshev step
__device__ __forceinline__ void cg_scalars_update_cheb(CgScalars* cgs, const double* tot) {
    const int it = cgs->it + 1;
    cgs->it = it;
    cgs->res2 = tot[1];
    if (tot[1] <= cgs->tol2) cgs->done = 1;
    else if (it >= cgs->max_iter || tot[1] != tot[1]) cgs->done = 2;
}
// ---------------------------------------------------------------------------------------
// Peer-mapped all-reduce of the two sums of a reduction kernel, executed by ONE thread (the one that holds the grid
// totals): the sums go straight into every rank's slot record over NVLink (plain remote stores, then a release store
// of the sequence number at system scope), then the thread polls its own record until all ranks have delivered and
// adds the contributions in rank order -- every rank obtains bit-identical totals, hence identical alpha / beta /
// convergence decisions, without an NCCL call or a kernel boundary.  A slot is rewritten only after its reader has
// consumed it: the writer's next reduction of the same kind lies behind a reduction of the other kind, which needs
// the reader's contribution.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
constexpr long long P2P_SPIN_LIMIT = 8000000000LL;      // ~4 s of SM clocks: a peer that died must not hang the GPU

// One WARP per rank: lane p talks to rank p.  It stores this rank's two sums into p's record, publishes the sequence number
// (st.release.sys orders it behind the sums) and polls its own record for p's contribution; the lanes then add the
// contributions in rank order through shuffles, so every rank obtains the same bits.  (The first version did all of this
// in ONE thread: 8 release stores over NVLink and 8 polls back to back, ~20 us per all-reduce at 8 GPUs.)
__device__ __noinline__ bool p2p_allreduce2(P2pDesc* D, int kind, double* tot) {
    const int lane = threadIdx.x & 31;
    const long long seq = D->red_count[kind] + 1;
    const double t0v = tot[0], t1v = tot[1];
    __syncwarp();
    if (lane == 0) D->red_count[kind] = seq;
    const int me = D->rank, W = D->world;
    double a = 0, b = 0;
    bool ok = true;
    if (lane < W) {
        volatile double* dst = D->slots[lane]->red[kind][me];
        dst[0] = t0v; dst[1] = t1v;
        st_release_sys(&D->slots[lane]->red_seq[kind][me], seq);
        P2pSlots* mine = D->slots[me];
        const long long c0 = clock64();
        while (ld_acquire_sys(&mine->red_seq[kind][lane]) < seq)
            if (clock64() - c0 > P2P_SPIN_LIMIT) { ok = false; break; }
        const volatile double* src = mine->red[kind][lane];
        a = src[0]; b = src[1];
    }
    ok = __all_sync(0xffffffffu, ok);
    double sa = 0, sb = 0;
    for (int p = 0; p < W; ++p) { sa += __shfl_sync(0xffffffffu, a, p); sb += __shfl_sync(0xffffffffu, b, p); }
    __syncwarp();
    if (lane == 0) { tot[0] = sa; tot[1] = sb; }
    __syncwarp();
    return ok;
}

template <bool INIT>
__device__ __forceinline__ void cg_finish_spmv(CgScalars* cgs, const double* tot, double* alpha_out) {
    if (cgs->red) { cgs->red[0] = tot[0]; cgs->red[1] = tot[1]; }
    else cg_scalars_spmv<INIT>(cgs, tot, alpha_out);
}
__device__ __forceinline__ void cg_finish_update(CgScalars* cgs, const double* tot, double* beta_out) {
    if (cgs->red) { cgs->red[0] = tot[0]; cgs->red[1] = tot[1]; }
    else cg_scalars_update(cgs, tot, beta_out);
}
// The scalar kernels (one thread) of the partitioned CG: in peer-mapped mode they ARE the all-reduce.  It lives here
// and not in the epilogue of the reduction kernels on purpose: a call in k_spmv_jds costs that kernel its register
// schedule (78 registers + stack instead of 80 and none, SpMV 0.74 -> 0.85 ms on half of X).
template <bool INIT>
__global__ void __launch_bounds__(32) k_cg_scalars_spmv(CgScalars* cgs, double* alpha_out) {      // one warp
    if (!INIT && cgs->done) return;
    if (cgs->p2p && !p2p_allreduce2(cgs->p2p, 0, cgs->red)) { if (threadIdx.x == 0) cgs->done = 3; return; }
    if (threadIdx.x == 0) cg_scalars_spmv<INIT>(cgs, cgs->red, alpha_out);
}
__global__ void __launch_bounds__(32) k_cg_scalars_update(CgScalars* cgs, double* beta_out) {
    if (cgs->done) return;
    if (cgs->p2p && !p2p_allreduce2(cgs->p2p, 1, cgs->red)) { if (threadIdx.x == 0) cgs->done = 3; return; }
    if (threadIdx.x == 0) cg_scalars_update(cgs, cgs->red, beta_out);
}

// ---------------------------------------------------------------------------------------
// Q1 stiffness assembly, 2x2x2 Gauss, MappingQ1:
//   K_e(i,j) = sum_q JxW_q grad N_i(q) . grad N_j(q)       (PoissonSolver.cpp:249-255)
// Two phases per block of 128 hexahedra:
//   A. one thread per hexahedron integrates the upper triangle of K_e in registers (FP64 pipe, no memory traffic
//      beyond 8 indices + 24 coordinates) and parks it in shared memory;
//   B. the block re-maps itself to 8 lanes per hexahedron: lane i owns ROW dof_i of the global matrix, walks that row's
//      sorted column list ONCE (independent, sector-contiguous loads) and adds its 8 entries where the columns match.
//      The 8 atomics of a lane land in one contiguous row segment.
// (Round 1 had every thread do 64 binary searches of ~5 dependent loads each: 26.9 ms for 2.24e7 hexahedra, 0.09 of the
// HBM roofline; the scatter, not the arithmetic, was the cost.)  Summation order across cells is not fixed (FP64
// atomics): ~1e-16 relative run-to-run variation, far below the 1e-8 parity bar.
// Algorithmic bytes per hex: 32 B connectivity + 192 B coordinates + 64 x 8 B adds.
// ---------------------------------------------------------------------------------------
constexpr int ASM_BLOCK = 128;
constexpr int ASM_STRIDE = 37;        // 36 upper-triangle entries + 1: odd stride in doubles = conflict-free per half-warp

__device__ __forceinline__ void hex_stiffness(const double (&X)[8], const double (&Y)[8], const double (&Z)[8], double (&Ke)[36], double& vol) {
#pragma unroll
    for (int k = 0; k < 36; ++k) Ke[k] = 0;
    const double ga = 0.5 * (1.0 - 0.57735026918962576451), gb = 0.5 * (1.0 + 0.57735026918962576451);
    vol = 0;
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
}

// cell volumes alone (DealSolver::get_cell_vol, DealSolver.cpp:169-173 = sum of JxW)
__global__ void __launch_bounds__(ASM_BLOCK) k_cell_volumes(int n_cells, const int* __restrict__ cells, const double* __restrict__ vxyz,
                                                            double* __restrict__ cell_vol) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    double X[8], Y[8], Z[8], Ke[36], vol;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = __ldg(&cells[8 * (size_t) c + i]);
        X[i] = __ldg(&vxyz[3 * (size_t) d]); Y[i] = __ldg(&vxyz[3 * (size_t) d + 1]); Z[i] = __ldg(&vxyz[3 * (size_t) d + 2]);
    }
    hex_stiffness(X, Y, Z, Ke, vol);
    cell_vol[c] = vol;
}

// phase B of the assembly kernels: lane i of an 8-lane group adds row i of one element matrix (upper triangle parked in
// shared memory by phase A) to global row dof_i
__device__ __forceinline__ void asm_scatter_rows(const double* s_ke, const int* s_dof, int n_rows, const int* __restrict__ rowptr,
                                                 const int* __restrict__ col, double* __restrict__ val) {
    const int tid = threadIdx.x;
    const int lane = tid & 7, sub = tid >> 3;
#pragma unroll 1
    for (int pass = 0; pass < ASM_BLOCK / 16; ++pass) {
        const int h = pass * 16 + sub;
        int d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = s_dof[8 * h + j];
        const int r = d[lane];
        // rows >= n_rows belong to another rank (ghost dofs of a partitioned mesh): that rank assembles them
        if (r < 0 || r >= n_rows) continue;
        double kv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int a = min(lane, j), b = max(lane, j);
            kv[j] = s_ke[h * ASM_STRIDE + (a * 8 - (a * (a - 1)) / 2 + (b - a))];
        }
        const int lo = __ldg(&rowptr[r]), hi = __ldg(&rowptr[r + 1]);
        int found = 0;
        for (int k = lo; k < hi && found < 8; ++k) {
            const int cc = __ldg(&col[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (cc == d[j]) { atomicAdd(&val[k], kv[j]); ++found; }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Scatter map of the assembly (built once per mesh, on the device): map[(8 i + j) n_cells + c] = position of entry
// (dof_i, dof_j) of hexahedron c in the CSR value array, or -1 when row dof_i belongs to another rank.  With it the
// assembly kernels add their 64 entries straight to their slots: no search of the column lists per entry (round 1:
// 64 binary searches per thread, 26.9 ms on X) and no walk of the row per lane (first version of this round: 51 ms --
// the dependent column loads of the walk, not the atomics, were the cost).  Stored entry-major, so that the 64 map
// loads of a warp are coalesced.  256 B per hexahedron (5.7 GB on X).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_asm_map(int n_cells, int n_rows, const int* __restrict__ cells, const int* __restrict__ rowptr,
                                                       const int* __restrict__ col, int* __restrict__ map) {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;          // one thread per (hexahedron, local row)
    const long c = t >> 3; const int i = (int) (t & 7);
    if (c >= n_cells) return;
    int d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = __ldg(&cells[8 * (size_t) c + j]);
    const int r = d[i];
    const bool mine = r < n_rows;
    const int lo0 = mine ? __ldg(&rowptr[r]) : 0, hi0 = mine ? __ldg(&rowptr[r + 1]) : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int lo = lo0, hi = hi0;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&col[mid]) < d[j]) lo = mid + 1; else hi = mid; }
        map[(size_t) (8 * i + j) * n_cells + c] = mine ? lo : -1;
    }
}

__device__ __forceinline__ void asm_scatter_mapped(const double (&Ke)[36], long c, int n_cells, const int* __restrict__ map, double* __restrict__ val) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int a = i < j ? i : j, b = i < j ? j : i;
            const int pos = __ldg(&map[(size_t) (8 * i + j) * n_cells + c]);
            if (pos >= 0) atomicAdd(&val[pos], Ke[a * 8 - (a * (a - 1)) / 2 + (b - a)]);
        }
}

__global__ void __launch_bounds__(ASM_BLOCK) k_assemble_stiffness_mapped(int n_cells, const int* __restrict__ cells, const double* __restrict__ vxyz,
                                                                        const int* __restrict__ map, double* __restrict__ val) {
    const long c = (long) blockIdx.x * ASM_BLOCK + threadIdx.x;
    if (c >= n_cells) return;
    double X[8], Y[8], Z[8], Ke[36], vol;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = __ldg(&cells[8 * (size_t) c + i]);
        X[i] = __ldg(&vxyz[3 * (size_t) d]); Y[i] = __ldg(&vxyz[3 * (size_t) d + 1]); Z[i] = __ldg(&vxyz[3 * (size_t) d + 2]);
    }
    hex_stiffness(X, Y, Z, Ke, vol);
    asm_scatter_mapped(Ke, c, n_cells, map, val);
}

// DealSolver::calc_dof_volumes (DealSolver.cpp:344-366): every dof receives the whole volume (sum of JxW) of each cell
// it belongs to; then charge_density = rhs / dof_volume (PoissonSolver.cpp:198-205)
__global__ void __launch_bounds__(ASM_BLOCK) k_dof_volumes(int n_cells, int n_rows, const int* __restrict__ cells, const double* __restrict__ vxyz,
                                                           double* __restrict__ dof_vol) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    double X[8], Y[8], Z[8], Ke[36], vol; int d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        d[i] = __ldg(&cells[8 * (size_t) c + i]);
        X[i] = __ldg(&vxyz[3 * (size_t) d[i]]); Y[i] = __ldg(&vxyz[3 * (size_t) d[i] + 1]); Z[i] = __ldg(&vxyz[3 * (size_t) d[i] + 2]);
    }
    hex_stiffness(X, Y, Z, Ke, vol);
#pragma unroll
    for (int i = 0; i < 8; ++i) if (d[i] < n_rows) atomicAdd(&dof_vol[d[i]], vol);
}
__global__ void k_charge_density(int n, const double* __restrict__ rhs, const double* __restrict__ dof_vol, double* __restrict__ rho) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) rho[i] = rhs[i] / dof_vol[i];
}

__global__ void __launch_bounds__(ASM_BLOCK) k_assemble_stiffness(int n_cells, int n_rows, const int* __restrict__ cells,
                                                                 const double* __restrict__ vxyz,
                                                                 const int* __restrict__ rowptr, const int* __restrict__ col,
                                                                 double* __restrict__ val) {
    __shared__ double s_ke[ASM_BLOCK * ASM_STRIDE];
    __shared__ int s_dof[ASM_BLOCK * 8];
    const int tid = threadIdx.x;
    const int c = blockIdx.x * ASM_BLOCK + tid;
    // ---- phase A: element matrix of hexahedron c ----
    if (c < n_cells) {
        double X[8], Y[8], Z[8], Ke[36], vol;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int d = __ldg(&cells[8 * (size_t) c + i]);
            s_dof[8 * tid + i] = d;
            X[i] = __ldg(&vxyz[3 * (size_t) d]); Y[i] = __ldg(&vxyz[3 * (size_t) d + 1]); Z[i] = __ldg(&vxyz[3 * (size_t) d + 2]);
        }
        hex_stiffness(X, Y, Z, Ke, vol);
#pragma unroll
        for (int k = 0; k < 36; ++k) s_ke[tid * ASM_STRIDE + k] = Ke[k];
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_dof[8 * tid + i] = -1;
    }
    __syncthreads();
    asm_scatter_rows(s_ke, s_dof, n_rows, rowptr, col, val);
}

// ---------------------------------------------------------------------------------------
// HeatSolver::assemble_local_cell (CurrentHeatSolver.cpp:346-393), implicit Euler:
//   K_e(i,j) = sum_q JxW_q ( gamma N_i N_j + kappa(T_q) grad N_i . grad N_j ),  gamma = cu_rho_cp / dt
//   f_e(i)   = sum_q JxW_q N_i ( gamma T_q + sigma(T_q) |grad phi_q|^2 )
// T_q / grad phi_q = previous temperature / gradient of the current potential at the Gauss point
// (FEValues::get_function_values / get_function_gradients); sigma, kappa from the resistivity table
// (PhysicalQuantities.cpp:31-64,173-185: clamped linear interpolation, Wiedemann-Franz).
// Same two phases as k_assemble_stiffness; the load vector goes out with 8 atomics per hexahedron.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double pq_resistivity(const double* __restrict__ tab_T, const double* __restrict__ tab_rho, int n, double T) {
    T = fmin(fmax(T, __ldg(&tab_T[0])), __ldg(&tab_T[n - 1]));
    if (T <= __ldg(&tab_T[0])) return 10.0 * __ldg(&tab_rho[0]);
    if (T >= __ldg(&tab_T[n - 1])) return 10.0 * __ldg(&tab_rho[n - 1]);
    int lo = 0, hi = n;                                   // std::lower_bound: first entry with tab_T >= T
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&tab_T[mid]) < T) lo = mid + 1; else hi = mid; }
    const double t1 = __ldg(&tab_T[lo]), t2 = __ldg(&tab_T[lo - 1]), r1 = __ldg(&tab_rho[lo]), r2 = __ldg(&tab_rho[lo - 1]);
    return 10.0 * (r2 + (r1 - r2) * (T - t2) / (t1 - t2));
}

__global__ void __launch_bounds__(ASM_BLOCK) k_assemble_heat(int n_cells, int n_rows, const int* __restrict__ cells,
                                                            const double* __restrict__ vxyz, const int* __restrict__ rowptr,
                                                            const int* __restrict__ col, double* __restrict__ val, double* __restrict__ rhs,
                                                            const double* __restrict__ T_prev, const double* __restrict__ phi, double gamma,
                                                            const double* __restrict__ tab_T, const double* __restrict__ tab_rho, int n_tab,
                                                            double lorentz, const int* __restrict__ map) {
    __shared__ double s_ke[ASM_BLOCK * ASM_STRIDE];
    __shared__ int s_dof[ASM_BLOCK * 8];
    const int tid = threadIdx.x;
    const int c = blockIdx.x * ASM_BLOCK + tid;
    if (c < n_cells) {
        double X[8], Y[8], Z[8], Tn[8], Pn[8], Ke[36], Fe[8];
        int dof[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int d = __ldg(&cells[8 * (size_t) c + i]);
            dof[i] = d; s_dof[8 * tid + i] = d;
            X[i] = __ldg(&vxyz[3 * (size_t) d]); Y[i] = __ldg(&vxyz[3 * (size_t) d + 1]); Z[i] = __ldg(&vxyz[3 * (size_t) d + 2]);
            Tn[i] = T_prev[d]; Pn[i] = phi[d];
            Fe[i] = 0;
        }
#pragma unroll
        for (int k = 0; k < 36; ++k) Ke[k] = 0;
        const double ga = 0.5 * (1.0 - 0.57735026918962576451), gb = 0.5 * (1.0 + 0.57735026918962576451);
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const double xi = (q & 1) ? gb : ga, eta = (q & 2) ? gb : ga, zeta = (q & 4) ? gb : ga;
            double N[8], dN[8][3];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double fx = (i & 1) ? xi : 1.0 - xi, fy = (i & 2) ? eta : 1.0 - eta, fz = (i & 4) ? zeta : 1.0 - zeta;
                const double sx = (i & 1) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0, sz = (i & 4) ? 1.0 : -1.0;
                N[i] = fx * fy * fz;
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
            double G[8][3], Tq = 0, gp[3] = {0, 0, 0};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int d = 0; d < 3; ++d) G[i][d] = dN[i][0] * inv[0][d] + dN[i][1] * inv[1][d] + dN[i][2] * inv[2][d];
                Tq += Tn[i] * N[i];
                gp[0] += G[i][0] * Pn[i]; gp[1] += G[i][1] * Pn[i]; gp[2] += G[i][2] * Pn[i];
            }
            const double sigma = 1.0 / pq_resistivity(tab_T, tab_rho, n_tab, Tq);
            const double Tc = fmin(fmax(Tq, __ldg(&tab_T[0])), __ldg(&tab_T[n_tab - 1]));
            const double kappa = lorentz * Tc * sigma;
            const double src = gamma * Tq + sigma * (gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2]);
            int k = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                Fe[i] += JxW * N[i] * src;
#pragma unroll
                for (int j = i; j < 8; ++j, ++k)
                    Ke[k] += JxW * (gamma * N[i] * N[j] + kappa * (G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2]));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) if (dof[i] < n_rows) atomicAdd(&rhs[dof[i]], Fe[i]);
        if (map) asm_scatter_mapped(Ke, c, n_cells, map, val);
        else {
#pragma unroll
            for (int k = 0; k < 36; ++k) s_ke[tid * ASM_STRIDE + k] = Ke[k];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_dof[8 * tid + i] = -1;
    }
    if (map) return;                                       // (uniform: kernel argument)
    __syncthreads();
    asm_scatter_rows(s_ke, s_dof, n_rows, rowptr, col, val);
}

// Neumann faces (DealSolver.cpp:389-430): b_i += sum_q N_i(q) * bc * JxW_face(q), QGauss<2>(2)
__global__ void k_neumann_faces(int n_faces, int n_rows, const int* __restrict__ face_dofs, const double* __restrict__ vxyz,
                                double bc_uniform, const double* __restrict__ face_bc, double* __restrict__ rhs) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    // face_bc: EmissionSolver::get_face_bc (CurrentHeatSolver.h:53-56), one value per face in cell/face iteration order
    const double bc_value = face_bc ? face_bc[f] : bc_uniform;
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
    for (int i = 0; i < 4; ++i) if (d[i] < n_rows) atomicAdd(&rhs[d[i]], r[i]);
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

// rhs finalisation of the ELIMINATED system (test hook fb_get_system only): b_c = a_cc v_c on constrained rows,
// b_r -= lift_r elsewhere
__global__ void k_eliminated_rhs(int n, const int* __restrict__ flag, const double* __restrict__ bcval,
                                 const double* __restrict__ diag_inv, const double* __restrict__ lift,
                                 const double* __restrict__ rhs_raw, double* __restrict__ rhs_out) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x)
        rhs_out[i] = flag[i] ? bcval[i] / diag_inv[i] : rhs_raw[i] - lift[i];
}

// Dirichlet conditions as a MASK on the immutable stiffness matrix K (the reference eliminates rows and columns of a
// copy, DealSolver.cpp:437-440 / PoissonSolver.cpp:157-159; here K is never rewritten):
//   dinv_r = 1 / K_rr on free rows and 0 on constrained rows.  The solver starts from x_c = v_c, so the initial residual
//   g = K x - b of a free row already contains the lift sum_c K_rc v_c; every kernel that writes g forces g_c = 0 where
//   dinv_c = 0, hence z_c = d_c = 0 for the whole solve, x_c stays v_c, and the products K_rc d_c vanish exactly: the
//   iterates are those of the eliminated system (up to the summation order inside a row).
// One thread per row: binary search of the diagonal (columns sorted), also kept as diagpos for the symmetric layout.
__global__ void k_bc_prepare(int n, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                             const int* __restrict__ flag, double* __restrict__ dinv, int* __restrict__ diagpos) {
    for (long r = (long) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long) gridDim.x * blockDim.x) {
        int lo = rowptr[r], hi = rowptr[r + 1];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < (int) r) lo = mid + 1; else hi = mid; }
        diagpos[r] = lo;
        dinv[r] = flag[r] ? 0.0 : 1.0 / val[lo];
    }
}

// x_c = v_c on constrained dofs (the right-hand side of a constrained row is never read)
__global__ void k_bc_solution(int n, const int* __restrict__ flag, const double* __restrict__ bcval, double* __restrict__ rhs,
                              double* __restrict__ x) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x)
        if (flag[i]) { x[i] = bcval[i]; rhs[i] = 0.0; }
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
            if (INIT) {                      // g = A x - b (0 on constrained rows: dinv = 0); accumulate g.(Dinv g) and g.g
                const double g = dinv[r] != 0.0 ? s - rhs[r] : 0.0;
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
        cg_finish_spmv<INIT>(cgs, tot, alpha_out);
    }
}

// ---------------------------------------------------------------------------------------
// Row-block ("CSR-stream") SpMV for HBM-sized systems.  The host cuts the rows into blocks of
// whole rows holding <= SPMV_CHUNK non-zeros and <= 256 rows (ctx: rowblk).  A CTA streams the
// block's val/col arrays with perfectly coalesced, evict-first loads (each thread SPMV_CHUNK/256
// independent loads in flight), multiplies by the gathered vector entries (read-only path, L1/L2
// resident) into shared memory, then one thread per row adds its products in column order --
// the same summation order as a sequential CSR sweep, so the result is deterministic.
// Algorithmic bytes per launch: 12 nnz + 4 (n+1) + 16 n.
// ---------------------------------------------------------------------------------------
template <bool INIT, int THREADS, int PER_THREAD>
__global__ void __launch_bounds__(THREADS) k_spmv_stream(int n_blocks, const int* __restrict__ rowblk, const int* __restrict__ rowptr,
                                                         const int* __restrict__ col, const double* __restrict__ val,
                                                         const double* __restrict__ xin, const double* __restrict__ rhs,
                                                         const double* __restrict__ dinv, double* __restrict__ out,
                                                         double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                         double* __restrict__ alpha_out) {
    if (!INIT && cgs->done) return;
    constexpr int CHUNK = THREADS * PER_THREAD;
    __shared__ double s_prod[CHUNK];
    __shared__ int s_rp[THREADS + 1];
    const int tid = threadIdx.x;
    double acc[2] = {0, 0};
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int r0 = __ldg(&rowblk[b]), nr = __ldg(&rowblk[b + 1]) - r0;
        if (tid <= nr) s_rp[tid] = __ldg(&rowptr[r0 + tid]);
        __syncthreads();
        const int k0 = s_rp[0], k1 = s_rp[nr];
        const int* __restrict__ cb = col + k0;
        const double* __restrict__ vb = val + k0;
        const int cnt = k1 - k0;
        int cidx[PER_THREAD];
        double v[PER_THREAD];
#pragma unroll
        for (int u = 0; u < PER_THREAD; ++u) {
            const int k = tid + u * THREADS;
            cidx[u] = (k < cnt) ? __ldcg(&cb[k]) : -1;
            v[u] = (k < cnt) ? __ldcg(&vb[k]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PER_THREAD; ++u)
            if (cidx[u] >= 0) s_prod[tid + u * THREADS] = v[u] * __ldg(&xin[cidx[u]]);
        __syncthreads();
        if (tid < nr) {
            const int lo = s_rp[tid] - k0, hi = s_rp[tid + 1] - k0;
            double sum = 0;
            for (int j = lo; j < hi; ++j) sum += s_prod[j];
            const int r = r0 + tid;
            if (INIT) {
                const double g = dinv[r] != 0.0 ? sum - rhs[r] : 0.0;
                out[r] = g;
                acc[0] += g * g * dinv[r];
                acc[1] += g * g;
            } else {
                out[r] = sum;
                acc[0] += __ldg(&xin[r]) * sum;
            }
        }
        __syncthreads();
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<INIT>(cgs, tot, alpha_out);
    }
}

template <bool ZERO_H, bool CHEB>
__global__ void __launch_bounds__(256, 6) k_update(int n, const double* __restrict__ d, double* __restrict__ h,
                                                const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ g,
                                                double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                const double* __restrict__ alpha_in, double* __restrict__ beta_out,
                                                double* __restrict__ cheb_p, double* __restrict__ cheb_z, double inv_theta) {
    if (cgs->done) return;
    const double alpha = *alpha_in;
    double acc[2] = {0, 0};
    // two entries per thread and trip with 16-byte accesses; all five loads of a trip are issued before the first
    // store, so every thread keeps 80 B in flight (the scalar loop serialised load -> fma -> store per vector)
    const long n2 = n >> 1, stride = (long) gridDim.x * blockDim.x;
    const double2* __restrict__ d2 = reinterpret_cast<const double2*>(d);
    const double2* __restrict__ i2 = reinterpret_cast<const double2*>(dinv);
    double2* __restrict__ h2 = reinterpret_cast<double2*>(h);
    double2* __restrict__ x2 = reinterpret_cast<double2*>(x);
    double2* __restrict__ g2 = reinterpret_cast<double2*>(g);
#pragma unroll 1
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 dd = __ldg(&d2[i]), di = __ldg(&i2[i]), hh = h2[i];
        double2 xx = x2[i], gg = g2[i];
        xx.x += alpha * dd.x; xx.y += alpha * dd.y;
        // constrained rows (dinv = 0) keep g = 0: the rows of K are not eliminated, see k_bc_prepare
        gg.x = di.x != 0.0 ? gg.x + alpha * hh.x : 0.0; gg.y = di.y != 0.0 ? gg.y + alpha * hh.y : 0.0;
        x2[i] = xx; g2[i] = gg;
        if (ZERO_H) h2[i] = make_double2(0.0, 0.0);   // the symmetric SpMV accumulates into h with reductions: hand it a zeroed vector
        if (CHEB) {                                   // first Chebyshev step: p = Dinv g / theta, z = p
            const double2 pp = make_double2(di.x * gg.x * inv_theta, di.y * gg.y * inv_theta);
            reinterpret_cast<double2*>(cheb_p)[i] = pp; reinterpret_cast<double2*>(cheb_z)[i] = pp;
        }
        acc[0] += gg.x * gg.x * di.x + gg.y * gg.y * di.y;
        acc[1] += gg.x * gg.x + gg.y * gg.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int i = n - 1;
        x[i] += alpha * d[i];
        const double gi = dinv[i] != 0.0 ? g[i] + alpha * h[i] : 0.0;
        g[i] = gi;
        if (ZERO_H) h[i] = 0.0;
        if (CHEB) { const double pp = dinv[i] * gi * inv_theta; cheb_p[i] = pp; cheb_z[i] = pp; }
        acc[0] += gi * gi * dinv[i];
        acc[1] += gi * gi;
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        if (CHEB) cg_scalars_update_cheb(cgs, tot); else cg_finish_update(cgs, tot, beta_out);
    }
}

// ---------------------------------------------------------------------------------------
// Chebyshev polynomial preconditioner  z = p_k(Dinv A) Dinv g  (k = option "cheb_degree" >= 2) on the interval
// [lmax / ratio, lmax] of the Jacobi-scaled operator; lmax = Gershgorin bound max_i sum_j |a_ij| / a_ii computed on
// the device after every assembly.  Three-term recurrence (Saad, Iterative Methods, Alg. 12.1) with
//   theta = (lmax + lmin) / 2, delta = (lmax - lmin) / 2, sigma = theta / delta, rho_0 = 1 / sigma:
//   p_0 = Dinv g / theta, z = p_0;   r_i = r_{i-1} - A p_{i-1};  rho_i = 1 / (2 sigma - rho_{i-1});
//   p_i = rho_i rho_{i-1} p_{i-1} + (2 rho_i / delta) Dinv r_i;  z += p_i            (i = 1 .. k-1)
// The coefficients depend only on (lmax, ratio, i) and are computed on the host.  One step = one SpMV (k_spmv_jds,
// its fused dot product unused) + k_cheb_step [64 B per DoF]; the last step also reduces g.z, from which beta follows.
// The polynomial is fixed and positive on (0, lmax], so M^-1 is SPD and CG stays CG; each iteration costs k SpMVs and
// needs ~sqrt-fewer iterations -- and, on several GPUs, fewer all-reduces per unit of work.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gershgorin(int n, const int* __restrict__ rowptr, const double* __restrict__ val,
                                                    const double* __restrict__ dinv, unsigned long long* __restrict__ out_bits) {
    // 8 lanes per row; warp-uniform trip count; positive doubles order like their bit patterns
    const int lane = threadIdx.x & 7;
    const long gtid = (long) blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long) gridDim.x * blockDim.x / 8;
    double mx = 0;
    for (long rb = (gtid / 32) * 4; rb < n; rb += stride) {
        const long r = rb + (threadIdx.x % 32) / 8;
        double s = 0;
        if (r < n) for (int k = rowptr[r] + lane; k < rowptr[r + 1]; k += 8) s += fabs(val[k]);
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (r < n) mx = fmax(mx, s * dinv[r]);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long) __double_as_longlong(mx));
}

// Power iteration on Dinv A for a sharper lmax than the Gershgorin bound: v <- Dinv A v (unnormalised: <= lmax^its
// growth), Rayleigh quotient (v.Av) / (v.Dv) of the similar symmetric matrix D^-1/2 A D^-1/2 -- a lower bound that
// converges from below, hence the safety factor applied by cheb_prepare.
__global__ void k_power_init(int n, const double* __restrict__ dinv, double* __restrict__ v) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        unsigned x = (unsigned) i * 2654435761u + 12345u;          // deterministic pseudo-random start in [-0.5, 0.5)
        x ^= x >> 16; x *= 2246822519u; x ^= x >> 13;
        v[i] = dinv[i] != 0.0 ? (double) x * (1.0 / 4294967296.0) - 0.5 : 0.0;      // constrained dofs stay out of the iteration
    }
}
__global__ void __launch_bounds__(256) k_power_step(int n, const double* __restrict__ h, const double* __restrict__ dinv,
                                                    double* __restrict__ v, double* __restrict__ partial, unsigned* counter,
                                                    double* __restrict__ out2) {
    double acc[2] = {0, 0};
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const double vi = v[i], hi = h[i], di = dinv[i];
        if (di != 0.0) { acc[0] += vi * hi; acc[1] += vi * vi / di; }
        v[i] = di * hi;
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) { out2[0] = tot[0]; out2[1] = tot[1]; }
}

__global__ void __launch_bounds__(256) k_cheb_start(int n, const double* __restrict__ g, const double* __restrict__ dinv,
                                                    double* __restrict__ p, double* __restrict__ z, double inv_theta,
                                                    const CgScalars* __restrict__ cgs) {
    if (cgs->done) return;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const double pp = dinv[i] * g[i] * inv_theta;
        p[i] = pp; z[i] = pp;
    }
}

// r = r_in - w;  p = c1 p + c2 Dinv r;  z += p.   LAST: also g.z -> beta = g.z / gh (INIT: only gh = g.z)
template <bool LAST, bool INIT>
__global__ void __launch_bounds__(256) k_cheb_step(int n, const double* __restrict__ w, const double* __restrict__ r_in, double* __restrict__ r_out,
                                                   const double* __restrict__ dinv, double* __restrict__ p, double* __restrict__ z,
                                                   const double* __restrict__ g, double c1, double c2,
                                                   double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                   double* __restrict__ beta_out) {
    if (cgs->done) return;
    double acc[1] = {0};
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const double r = r_in[i] - w[i];
        const double pp = c1 * p[i] + c2 * dinv[i] * r;
        const double zz = z[i] + pp;
        z[i] = zz;
        if (!LAST) { r_out[i] = r; p[i] = pp; }
        else acc[0] += g[i] * zz;
    }
    if (LAST) {
        double tot[1];
        if (reduce_publish<1>(acc, partial, counter, tot)) {
            if (!INIT) *beta_out = tot[0] / cgs->gh;
            cgs->gh = tot[0];
        }
    }
}

// d = beta d - z
template <bool INIT>
__global__ void __launch_bounds__(256) k_direction_z(int n, const double* __restrict__ z, double* __restrict__ d,
                                                     const CgScalars* __restrict__ cgs, const double* __restrict__ beta_in) {
    if (cgs->done) return;
    const double beta = INIT ? 0.0 : *beta_in;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x)
        d[i] = (INIT ? 0.0 : beta * d[i]) - z[i];
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

// ---------------------------------------------------------------------------------------
// Windowed streaming SpMV (the HBM-roofline kernel).  Same row blocks as k_spmv_stream, plus two
// host-built tables (fb_host_col_windows): the block's WINDOW = sorted distinct columns it touches,
// and a 16-bit window position per non-zero.  A CTA
//   1. issues its coalesced val (8 B) / col16 (2 B) loads, L1-bypassing (ld.global.cg),
//   2. stages the window of the input vector into shared memory (sorted -> dense sectors; each
//      distinct entry is fetched once per block instead of once per non-zero),
//   3. multiplies out of shared memory (no scattered global gathers: the L1 wavefront limit of the
//      per-non-zero gather is what capped the plain CSR kernels at ~0.45-0.65 of HBM peak),
//   4. adds the products of each row in column order (deterministic, = sequential CSR order).
// DRAM traffic is BELOW the algorithmic 12 nnz + 4(n+1) + 16 n bytes it is rated against
// (10 B per non-zero + windows), which is why its roofline fraction can approach / exceed 1.
// ---------------------------------------------------------------------------------------
template <bool INIT, int THREADS, int PER_THREAD>
__global__ void __launch_bounds__(THREADS) k_spmv_window(int n_blocks, const int* __restrict__ rowblk, const int* __restrict__ rowptr,
                                                         const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                         const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                         const double* __restrict__ xin, const double* __restrict__ rhs,
                                                         const double* __restrict__ dinv, double* __restrict__ out,
                                                         double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                         double* __restrict__ alpha_out, int wcap) {
    if (!INIT && cgs->done) return;
    constexpr int CHUNK = THREADS * PER_THREAD;
    extern __shared__ double s_dyn[];
    double* s_prod = s_dyn;                  // CHUNK products
    double* s_x = s_dyn + CHUNK;             // wcap window entries
    __shared__ int s_rp[THREADS + 1];
    const int tid = threadIdx.x;
    double acc[2] = {0, 0};
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int r0 = __ldg(&rowblk[b]), nr = __ldg(&rowblk[b + 1]) - r0;
        const int w0 = __ldg(&win_off[b]), nw = __ldg(&win_off[b + 1]) - w0;
        if (tid <= nr) s_rp[tid] = __ldg(&rowptr[r0 + tid]);
        const int k0 = __ldg(&rowptr[r0]), cnt = __ldg(&rowptr[r0 + nr]) - k0;
        const unsigned short* __restrict__ cb = col16 + k0;
        const double* __restrict__ vb = val + k0;
        int cidx[PER_THREAD];
        double v[PER_THREAD];
#pragma unroll
        for (int u = 0; u < PER_THREAD; ++u) {
            const int k = tid + u * THREADS;
            cidx[u] = (k < cnt) ? (int) __ldcg(&cb[k]) : -1;
            v[u] = (k < cnt) ? __ldcg(&vb[k]) : 0.0;
        }
        for (int i = tid; i < nw; i += THREADS) s_x[i] = __ldg(&xin[__ldg(&win_list[w0 + i])]);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER_THREAD; ++u)
            if (cidx[u] >= 0) s_prod[tid + u * THREADS] = v[u] * s_x[cidx[u]];
        __syncthreads();
        if (tid < nr) {
            const int lo = s_rp[tid] - k0, hi = s_rp[tid + 1] - k0;
            double sum = 0;
            for (int j = lo; j < hi; ++j) sum += s_prod[j];
            const int r = r0 + tid;
            if (INIT) {
                const double g = dinv[r] != 0.0 ? sum - rhs[r] : 0.0;
                out[r] = g;
                acc[0] += g * g * dinv[r];
                acc[1] += g * g;
            } else {
                out[r] = sum;
                acc[0] += __ldg(&xin[r]) * sum;
            }
        }
        __syncthreads();
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<INIT>(cgs, tot, alpha_out);
    }
}

// ---------------------------------------------------------------------------------------
// Block-JDS SpMV (the HBM-roofline kernel; tables from fb_host_jds_build).  One CTA = one block of
// R = blockDim.x consecutive rows; thread t owns the t-th longest row of the block and walks it
// diagonal by diagonal:  k = base + jd[j] + t  -> val (8 B) and the 16-bit window position are read
// perfectly coalesced with NO padding and NO shared-memory staging of products; the input vector is
// read from the block's window in shared memory (staged once per block from sorted, dense global
// addresses).  Per-row summation is in column order = the sequential CSR order (deterministic).
// L1/LSU work per 32 non-zeros: ~2 (val) + 0.5 (col16) + shared-memory gather, vs ~26 wavefronts for
// the per-non-zero global gather of plain CSR, which is what capped those kernels at 0.45-0.65 of peak.
// DRAM traffic ~ 10 B per non-zero + windows, i.e. BELOW the algorithmic 12 nnz + 4(n+1) + 16 n bytes.
// ---------------------------------------------------------------------------------------
__global__ void k_csr_to_jds(int n, int R, const int* __restrict__ rowptr, const unsigned short* __restrict__ slot,
                             const int* __restrict__ jbase, const int* __restrict__ jdp, const int* __restrict__ jd,
                             const double* __restrict__ val, double* __restrict__ val_jds,
                             const int* __restrict__ diagpos, double* __restrict__ diag) {
    // 8 lanes per row; diagpos != nullptr: symmetric layout, only the entries left of the diagonal (a prefix of
    // the sorted row) are stored and the diagonal goes to diag[]
    const int lane = threadIdx.x & 7;
    for (long r = ((long) blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n; r += ((long) gridDim.x * blockDim.x) >> 3) {
        const int b = (int) (r / R);
        const int base = jbase[b], lo = rowptr[r], hi = diagpos ? diagpos[r] : rowptr[r + 1], t = slot[r];
        if (diagpos && lane == 0) diag[r] = val[hi];
        const int* __restrict__ jdb = jd + jdp[b];
        for (int k = lo + lane; k < hi; k += 8) val_jds[(long) base + jdb[k - lo] + t] = val[k];
    }
}

// R rows per block, R / 2 threads: thread t owns the rows in slots 2t and 2t+1 (len0 >= len1) and fetches
// their entries of a diagonal with ONE 16-byte value load and ONE 4-byte window-position load.
// CS: the matrix stream is read with ld.global.cs (evict-first) so that it does not push the input vector,
// which the windows of neighbouring blocks re-read, out of L2 (option "spmv_kernel" 304).
// evict-first loads as volatile asm: ptxas keeps them in program order, which pins the software pipeline of the main
// loop (issue the next four diagonals, THEN consume the previous four) instead of leaving it to the scheduler's
// heuristics -- an unrelated edit once sank the loads behind the consumes and cost 0.8 ms per launch
__device__ __forceinline__ double2 ld_cs_v2(const double2* p) {
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ ushort2 ld_cs_us2(const ushort2* p) {
    unsigned w;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(w) : "l"(p));
    return make_ushort2((unsigned short) (w & 0xffffu), (unsigned short) (w >> 16));
}
template <bool INIT, int R, int CS>
__global__ void __launch_bounds__(R / 2) k_spmv_jds(int n, int n_blocks, const int* __restrict__ jbase,
                                                    const unsigned short* __restrict__ perm, const unsigned short* __restrict__ rlen,
                                                    const int* __restrict__ jdp, const int* __restrict__ jd,
                                                    const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                    const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                    const double* __restrict__ xin, const double* __restrict__ rhs,
                                                    const double* __restrict__ dinv, double* __restrict__ out,
                                                    double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                    double* __restrict__ alpha_out, int wcap, int jcap) {
    if (!INIT && cgs->done) return;
    constexpr int T = R / 2;
    extern __shared__ double s_dyn[];
    double* s_x = s_dyn;                      // wcap window entries
    int* s_jd = (int*) (s_dyn + wcap);        // jcap + 1 diagonal offsets (in 2-entry units)
    const int tid = threadIdx.x;
    double acc[2] = {0, 0};
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int r0 = b * R;
        const int w0 = __ldg(&win_off[b]), nw = __ldg(&win_off[b + 1]) - w0;
        const int j0 = __ldg(&jdp[b]), nj = __ldg(&jdp[b + 1]) - j0;          // maxlen + 1 offsets
        const long base = __ldg(&jbase[b]);
        const int sl0 = r0 + 2 * tid, sl1 = sl0 + 1;
        const int len0 = sl0 < n ? (int) __ldg(&rlen[sl0]) : 0, len1 = sl1 < n ? (int) __ldg(&rlen[sl1]) : 0;
        const int row0 = sl0 < n ? r0 + (int) __ldg(&perm[sl0]) : 0, row1 = sl1 < n ? r0 + (int) __ldg(&perm[sl1]) : 0;
        for (int i = tid; i < nw; i += T) s_x[i] = __ldg(&xin[__ldg(&win_list[w0 + i])]);
        for (int i = tid; i < nj; i += T) s_jd[i] = __ldg(&jd[j0 + i]) >> 1;
        __syncthreads();
        const double2* __restrict__ vb = reinterpret_cast<const double2*>(val + base) + tid;
        const ushort2* __restrict__ cb = reinterpret_cast<const ushort2*>(col16 + base) + tid;
        double sum0 = 0, sum1 = 0;
        // register double-buffered pipeline: the loads of the next 4 diagonals are issued before the current
        // 4 are consumed, so every thread keeps 4-8 x (16 B + 4 B) loads in flight without draining
        double2 va[4], vb2[4];
        ushort2 ca[4], cb2[4];
#define FB_ISSUE(JJ, V, C)                                                                     \
        _Pragma("unroll") for (int u = 0; u < 4; ++u)                                          \
            if ((JJ) + u < len0) { const int o = s_jd[(JJ) + u];                                     \
                if (CS == 2) { V[u] = ld_cs_v2(&vb[o]); C[u] = ld_cs_us2(&cb[o]); }                     \
                else { V[u] = CS ? __ldcs(&vb[o]) : __ldcg(&vb[o]); C[u] = CS ? __ldcs(&cb[o]) : __ldcg(&cb[o]); } }
#define FB_CONSUME(JJ, V, C)                                                                   \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                        \
            if ((JJ) + u < len0) sum0 += V[u].x * s_x[C[u].x];                                 \
            if ((JJ) + u < len1) sum1 += V[u].y * s_x[C[u].y];                                 \
        }
        FB_ISSUE(0, va, ca)
        for (int j = 0; j < len0; j += 8) {
            FB_ISSUE(j + 4, vb2, cb2)
            FB_CONSUME(j, va, ca)
            FB_ISSUE(j + 8, va, ca)
            FB_CONSUME(j + 4, vb2, cb2)
        }
#undef FB_ISSUE
#undef FB_CONSUME
        if (INIT) {
            if (sl0 < n) { const double di = dinv[row0]; const double g = di != 0.0 ? sum0 - rhs[row0] : 0.0; out[row0] = g; acc[0] += g * g * di; acc[1] += g * g; }
            if (sl1 < n) { const double di = dinv[row1]; const double g = di != 0.0 ? sum1 - rhs[row1] : 0.0; out[row1] = g; acc[0] += g * g * di; acc[1] += g * g; }
        } else {
            if (sl0 < n) { out[row0] = sum0; acc[0] += __ldg(&xin[row0]) * sum0; }
            if (sl1 < n) { out[row1] = sum1; acc[0] += __ldg(&xin[row1]) * sum1; }
        }
        __syncthreads();
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<INIT>(cgs, tot, alpha_out);
    }
}

// Block-JDS SpMV over the SEGMENTED layout (option "spmv_kernel" 306; tables: fb_host_jds_build with split > 0).  A row
// longer than the cap is stored as chained segments, each in a slot of its own, so every slot of a block walks at most
// `cap` diagonals: without it the one warp that owns the block's longest rows (high-valence vertices of the tetrahedral
// mesh: 57 entries on average against 27) runs twice as many trips as the other seven, which wait for it at the block's
// barrier (29 % of the warp samples of k_spmv_jds on the 3e6-DoF mesh).  A block = R slots covering the rows
// [rowbeg[b], rowbeg[b + 1]).  perm bit 15 marks a continuation segment: it leaves its sum in shared memory, and the head
// of the chain adds the segment sums in chain order after the barrier that ends the block (deterministic).  Two sets of
// sum / link buffers alternate between consecutive blocks of a CTA, so a head may still walk its chain while faster threads
// stage the next block.
template <bool INIT, int R>
__global__ void __launch_bounds__(R / 2) k_spmv_jdss(int n_blocks, const int* __restrict__ rowbeg, const int* __restrict__ jbase,
                                                     const unsigned short* __restrict__ perm, const unsigned short* __restrict__ rlen,
                                                     const unsigned short* __restrict__ link,
                                                     const int* __restrict__ jdp, const int* __restrict__ jd,
                                                     const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                     const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                     const double* __restrict__ xin, const double* __restrict__ rhs,
                                                     const double* __restrict__ dinv, double* __restrict__ out,
                                                     double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                     double* __restrict__ alpha_out, int wcap, int jcap) {
    if (!INIT && cgs->done) return;
    constexpr int T = R / 2;
    extern __shared__ double s_dyn[];
    double* s_x = s_dyn;                                          // wcap window entries
    double* s_sum = s_dyn + wcap;                                 // 2 x R segment sums
    int* s_jd = (int*) (s_sum + 2 * R);                           // jcap + 1 diagonal offsets (in 2-entry units), padded to even
    unsigned short* s_link = (unsigned short*) (s_jd + ((jcap + 2) & ~1));      // 2 x R chain links
    const int tid = threadIdx.x;
    double acc[2] = {0, 0};
    int par = 0;
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x, par ^= 1) {
        const int r0 = __ldg(&rowbeg[b]);
        const int w0 = __ldg(&win_off[b]), nw = __ldg(&win_off[b + 1]) - w0;
        const int j0 = __ldg(&jdp[b]), nj = __ldg(&jdp[b + 1]) - j0;
        const long base = __ldg(&jbase[b]);
        const size_t sl = (size_t) b * R + 2 * tid;
        const ushort2 ln = __ldg(reinterpret_cast<const ushort2*>(rlen + sl));
        const ushort2 pm = __ldg(reinterpret_cast<const ushort2*>(perm + sl));
        const int len0 = ln.x, len1 = ln.y;
        reinterpret_cast<ushort2*>(s_link + par * R)[tid] = __ldg(reinterpret_cast<const ushort2*>(link + sl));
        for (int i = tid; i < nw; i += T) s_x[i] = __ldg(&xin[__ldg(&win_list[w0 + i])]);
        for (int i = tid; i < nj; i += T) s_jd[i] = __ldg(&jd[j0 + i]) >> 1;
        __syncthreads();
        const double2* __restrict__ vb = reinterpret_cast<const double2*>(val + base) + tid;
        const ushort2* __restrict__ cb = reinterpret_cast<const ushort2*>(col16 + base) + tid;
        double sum0 = 0, sum1 = 0;
        double2 va[4], vb2[4];
        ushort2 ca[4], cb2[4];
#define FB_ISSUE(JJ, V, C)                                                                     \
        _Pragma("unroll") for (int u = 0; u < 4; ++u)                                          \
            if ((JJ) + u < len0) { const int o = s_jd[(JJ) + u]; V[u] = __ldcs(&vb[o]); C[u] = __ldcs(&cb[o]); }
#define FB_CONSUME(JJ, V, C)                                                                   \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                        \
            if ((JJ) + u < len0) sum0 += V[u].x * s_x[C[u].x];                                 \
            if ((JJ) + u < len1) sum1 += V[u].y * s_x[C[u].y];                                 \
        }
        FB_ISSUE(0, va, ca)
        for (int j = 0; j < len0; j += 8) {
            FB_ISSUE(j + 4, vb2, cb2)
            FB_CONSUME(j, va, ca)
            FB_ISSUE(j + 8, va, ca)
            FB_CONSUME(j + 4, vb2, cb2)
        }
#undef FB_ISSUE
#undef FB_CONSUME
        double* ss = s_sum + par * R;
        const unsigned short* lk = s_link + par * R;
        if (pm.x & 0x8000) ss[2 * tid] = sum0;                    // (a dead slot, perm 0xFFFF, parks a zero nobody reads)
        if (pm.y & 0x8000) ss[2 * tid + 1] = sum1;
        __syncthreads();                                          // ends the block: window and offsets may be overwritten
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned short p = h ? pm.y : pm.x;
            if (p & 0x8000) continue;
            double sum = h ? sum1 : sum0;
            for (unsigned s = lk[2 * tid + h]; s != 0xFFFFu; s = lk[s]) sum += ss[s];
            const int row = r0 + p;
            if (INIT) {
                const double di = dinv[row]; const double g = di != 0.0 ? sum - rhs[row] : 0.0;
                out[row] = g; acc[0] += g * g * di; acc[1] += g * g;
            } else {
                out[row] = sum; acc[0] += __ldg(&xin[row]) * sum;
            }
        }
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<INIT>(cgs, tot, alpha_out);
    }
}

// values of the CSR matrix into the segmented layout: entry o of a row with k segments of length sl sits in segment o / sl
__global__ void k_csr_to_jds_split(int n, int R, int n_blocks, int cap, const int* __restrict__ rowbeg, const int* __restrict__ rowptr,
                                   const unsigned short* __restrict__ slot, const unsigned short* __restrict__ link,
                                   const int* __restrict__ jbase, const int* __restrict__ jdp, const int* __restrict__ jd,
                                   const double* __restrict__ val, double* __restrict__ val_jds) {
    const int lane = threadIdx.x & 7;
    for (long r = ((long) blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n; r += ((long) gridDim.x * blockDim.x) >> 3) {
        int lo_b = 0, hi_b = n_blocks;                            // last block whose first row is <= r
        while (hi_b - lo_b > 1) { const int mid = (lo_b + hi_b) >> 1; if (rowbeg[mid] <= r) lo_b = mid; else hi_b = mid; }
        const int b = lo_b;
        const int base = jbase[b], lo = rowptr[r], len = rowptr[r + 1] - lo;
        const int k = len > cap ? (len + cap - 1) / cap : 1, sl = (len + k - 1) / k;
        const int* __restrict__ jdb = jd + jdp[b];
        const unsigned short* __restrict__ lk = link + (size_t) b * R;
        for (int o = lane; o < len; o += 8) {
            const int q = o / sl, j = o - q * sl;
            int t = slot[r];
            for (int i = 0; i < q; ++i) t = lk[t];
            val_jds[(long) base + jdb[j] + t] = val[lo + o];
        }
    }
}

// Block-JDS SpMV with the NEXT block's window prefetched (option "spmv_kernel" 303).  Same tables and arithmetic as
// k_spmv_jds<., 512>; the difference is the prologue: the window of the input vector and the diagonal offsets of the
// block a CTA will process next are gathered with cp.async (8-byte / 4-byte, straight into the second shared-memory
// buffer, no registers held) while the current block streams its matrix entries, so the two dependent L2 round
// trips (window list -> vector entries) and the barrier of the staging phase are off the critical path of every
// block but the first one of a CTA.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <bool INIT, int R>
__global__ void __launch_bounds__(R / 2) k_spmv_jdsp(int n, int n_blocks, const int* __restrict__ jbase,
                                                     const unsigned short* __restrict__ perm, const unsigned short* __restrict__ rlen,
                                                     const int* __restrict__ jdp, const int* __restrict__ jd,
                                                     const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                     const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                     const double* __restrict__ xin, const double* __restrict__ rhs,
                                                     const double* __restrict__ dinv, double* __restrict__ out,
                                                     double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                     double* __restrict__ alpha_out, int wcap, int jcap) {
    if (!INIT && cgs->done) return;
    constexpr int T = R / 2;
    extern __shared__ double s_dyn[];
    const int jstride = (jcap + 2 + 1) & ~1;                 // ints per offset buffer (even: keeps 8-byte alignment)
    int* const s_jall = (int*) (s_dyn + 2 * (size_t) wcap);
    const int tid = threadIdx.x;
    // gather the window and the diagonal offsets of block b into buffer `buf` (asynchronous)
    auto stage = [&](int b, int buf) {
        const int w0 = __ldg(&win_off[b]), nw = __ldg(&win_off[b + 1]) - w0;
        const int j0 = __ldg(&jdp[b]), nj = __ldg(&jdp[b + 1]) - j0;
        double* sx = s_dyn + (size_t) buf * wcap; int* sj = s_jall + (size_t) buf * jstride;
        for (int i0 = tid; i0 < nw; i0 += 4 * T) {           // four independent index loads, then four copies
            int idx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * T; idx[u] = (i < nw) ? __ldg(&win_list[w0 + i]) : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (idx[u] >= 0) cp_async8(&sx[i0 + u * T], &xin[idx[u]]);
        }
        for (int i = tid; i < nj; i += T) cp_async4(&sj[i], &jd[j0 + i]);
        cp_async_commit();
    };
    double acc[2] = {0, 0};
    int buf = 0;
    if ((int) blockIdx.x < n_blocks) stage(blockIdx.x, 0);
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x, buf ^= 1) {
        const int r0 = b * R;
        const long base = __ldg(&jbase[b]);
        const int sl0 = r0 + 2 * tid, sl1 = sl0 + 1;
        const int len0 = sl0 < n ? (int) __ldg(&rlen[sl0]) : 0, len1 = sl1 < n ? (int) __ldg(&rlen[sl1]) : 0;
        const int row0 = sl0 < n ? r0 + (int) __ldg(&perm[sl0]) : 0, row1 = sl1 < n ? r0 + (int) __ldg(&perm[sl1]) : 0;
        cp_async_wait_all();
        __syncthreads();             // buffer `buf` is complete; everybody has left the previous block (buffer buf ^ 1 is free)
        const double* __restrict__ s_x = s_dyn + (size_t) buf * wcap;
        const int* __restrict__ s_jd = s_jall + (size_t) buf * jstride;
        const double2* __restrict__ vb = reinterpret_cast<const double2*>(val + base) + tid;
        const ushort2* __restrict__ cb = reinterpret_cast<const ushort2*>(col16 + base) + tid;
        double sum0 = 0, sum1 = 0;
        double2 va[4], vb2[4];
        ushort2 ca[4], cb2[4];
#define FB_ISSUE(JJ, V, C)                                                                     \
        _Pragma("unroll") for (int u = 0; u < 4; ++u)                                          \
            if ((JJ) + u < len0) { const int o = s_jd[(JJ) + u] >> 1; V[u] = __ldcg(&vb[o]); C[u] = __ldcg(&cb[o]); }
#define FB_CONSUME(JJ, V, C)                                                                   \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                        \
            if ((JJ) + u < len0) sum0 += V[u].x * s_x[C[u].x];                                 \
            if ((JJ) + u < len1) sum1 += V[u].y * s_x[C[u].y];                                 \
        }
        FB_ISSUE(0, va, ca)
        if (b + (int) gridDim.x < n_blocks) stage(b + gridDim.x, buf ^ 1);      // prefetch behind the first matrix loads
        for (int j = 0; j < len0; j += 8) {
            FB_ISSUE(j + 4, vb2, cb2)
            FB_CONSUME(j, va, ca)
            FB_ISSUE(j + 8, va, ca)
            FB_CONSUME(j + 4, vb2, cb2)
        }
#undef FB_ISSUE
#undef FB_CONSUME
        if (INIT) {
            if (sl0 < n) { const double di = dinv[row0]; const double g = di != 0.0 ? sum0 - rhs[row0] : 0.0; out[row0] = g; acc[0] += g * g * di; acc[1] += g * g; }
            if (sl1 < n) { const double di = dinv[row1]; const double g = di != 0.0 ? sum1 - rhs[row1] : 0.0; out[row1] = g; acc[0] += g * g * di; acc[1] += g * g; }
        } else {
            if (sl0 < n) { out[row0] = sum0; acc[0] += __ldg(&xin[row0]) * sum0; }
            if (sl1 < n) { out[row1] = sum1; acc[0] += __ldg(&xin[row1]) * sum1; }
        }
    }
    cp_async_wait_all();
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<INIT>(cgs, tot, alpha_out);
    }
}

// k_spmv_jdss with the NEXT block's window INDEX LIST prefetched (option "spmv_kernel" 307; same tables as 306).  ncu of
// k_spmv_jdss on X: 28 % of the warp samples sit in the block prologue (+ 7.5 % at its barrier), on a chain of three
// dependent global loads -- win_off -> win_list -> vector entries -- that moves 5 % of the bytes.  Here the window
// offsets of the next block are loaded at the top of the current one and its index list is copied into the second of two
// shared-memory buffers with cp.async (4-byte, coalesced, no registers held) behind the first matrix loads of the main
// loop; the next prologue then gathers through indices that already sit in shared memory: one dependent level instead of
// three.  (The vector entries themselves cannot be prefetched: they change between launches but not within one -- they
// could, but the 8-byte scattered cp.async gathers of variant 303 were slower than LDG + STS.)
template <int R>
__global__ void __launch_bounds__(R / 2) k_spmv_jdsq(int n_blocks, const int* __restrict__ rowbeg, const int* __restrict__ jbase,
                                                     const unsigned short* __restrict__ perm, const unsigned short* __restrict__ rlen,
                                                     const unsigned short* __restrict__ link,
                                                     const int* __restrict__ jdp, const int* __restrict__ jd,
                                                     const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                     const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                     const double* __restrict__ xin, double* __restrict__ out,
                                                     double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                     double* __restrict__ alpha_out, int wcap, int jcap) {
    if (cgs->done) return;
    constexpr int T = R / 2;
    extern __shared__ double s_dyn[];
    double* s_x = s_dyn;                                          // wcap window entries
    double* s_sum = s_dyn + wcap;                                 // 2 x R segment sums
    int* s_jd = (int*) (s_sum + 2 * R);                           // jcap + 1 diagonal offsets (in 2-entry units), padded to even
    int* s_idx = s_jd + ((jcap + 2) & ~1);                        // 2 x wcap window indices (this block / next block)
    unsigned short* s_link = (unsigned short*) (s_idx + 2 * wcap);      // 2 x R chain links
    const int tid = threadIdx.x;
    double acc[2] = {0, 0};
    int par = 0;
    int nw = 0;
    if ((int) blockIdx.x < n_blocks) {                            // index list of the first block
        const int w0 = __ldg(&win_off[blockIdx.x]);
        nw = __ldg(&win_off[blockIdx.x + 1]) - w0;
        for (int i = tid; i < nw; i += T) cp_async4(&s_idx[i], &win_list[w0 + i]);
        cp_async_commit();
        cp_async_wait_all();
    }
    __syncthreads();
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x, par ^= 1) {
        const int bn = b + gridDim.x;
        int w0n = 0, nwn = 0;
        if (bn < n_blocks) { w0n = __ldg(&win_off[bn]); nwn = __ldg(&win_off[bn + 1]) - w0n; }
        const int r0 = __ldg(&rowbeg[b]);
        const int j0 = __ldg(&jdp[b]), nj = __ldg(&jdp[b + 1]) - j0;
        const long base = __ldg(&jbase[b]);
        const size_t sl = (size_t) b * R + 2 * tid;
        const ushort2 ln = __ldg(reinterpret_cast<const ushort2*>(rlen + sl));
        const ushort2 pm = __ldg(reinterpret_cast<const ushort2*>(perm + sl));
        const int len0 = ln.x, len1 = ln.y;
        reinterpret_cast<ushort2*>(s_link + par * R)[tid] = __ldg(reinterpret_cast<const ushort2*>(link + sl));
        const int* __restrict__ idx = s_idx + par * wcap;
        for (int i = tid; i < nw; i += T) s_x[i] = __ldg(&xin[idx[i]]);
        for (int i = tid; i < nj; i += T) s_jd[i] = __ldg(&jd[j0 + i]) >> 1;
        __syncthreads();
        const double2* __restrict__ vb = reinterpret_cast<const double2*>(val + base) + tid;
        const ushort2* __restrict__ cb = reinterpret_cast<const ushort2*>(col16 + base) + tid;
        double sum0 = 0, sum1 = 0;
        double2 va[4], vb2[4];
        ushort2 ca[4], cb2[4];
#define FB_ISSUE(JJ, V, C)                                                                     \
        _Pragma("unroll") for (int u = 0; u < 4; ++u)                                          \
            if ((JJ) + u < len0) { const int o = s_jd[(JJ) + u]; V[u] = __ldcs(&vb[o]); C[u] = __ldcs(&cb[o]); }
#define FB_CONSUME(JJ, V, C)                                                                   \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                        \
            if ((JJ) + u < len0) sum0 += V[u].x * s_x[C[u].x];                                 \
            if ((JJ) + u < len1) sum1 += V[u].y * s_x[C[u].y];                                 \
        }
        FB_ISSUE(0, va, ca)
        {   // index list of the next block into the other buffer (nobody reads that one before the barrier that ends this block)
            int* dst = s_idx + (par ^ 1) * wcap;
            for (int i = tid; i < nwn; i += T) cp_async4(&dst[i], &win_list[w0n + i]);
            cp_async_commit();
        }
        nw = nwn;
        for (int j = 0; j < len0; j += 8) {
            FB_ISSUE(j + 4, vb2, cb2)
            FB_CONSUME(j, va, ca)
            FB_ISSUE(j + 8, va, ca)
            FB_CONSUME(j + 4, vb2, cb2)
        }
#undef FB_ISSUE
#undef FB_CONSUME
        double* ss = s_sum + par * R;
        const unsigned short* lk = s_link + par * R;
        if (pm.x & 0x8000) ss[2 * tid] = sum0;
        if (pm.y & 0x8000) ss[2 * tid + 1] = sum1;
        cp_async_wait_all();
        __syncthreads();                                          // ends the block: window, offsets free; next index list complete
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned short p = h ? pm.y : pm.x;
            if (p & 0x8000) continue;
            double sum = h ? sum1 : sum0;
            for (unsigned s = lk[2 * tid + h]; s != 0xFFFFu; s = lk[s]) sum += ss[s];
            const int row = r0 + p;
            out[row] = sum; acc[0] += __ldg(&xin[row]) * sum;
        }
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<false>(cgs, tot, alpha_out);
    }
}

// ---------------------------------------------------------------------------------------
// Segmented block-JDS SpMV with the MATRIX STREAM fed by TMA bulk copies (option "spmv_kernel" 308; tables of 306).
// In k_spmv_jdss the loads in flight live in the registers of the warps that will consume them, so a CTA that is staging
// the window of its next block (28 % of the warp samples) streams nothing.  Here a producer warp walks the CTA's blocks
// ahead of the consumers and copies the value / column streams -- contiguous per block -- group by group (4 diagonals) into
// a ring of shared-memory stages with cp.async.bulk (1-D TMA, completion on an mbarrier, L2 evict-first policy); the 8
// consumer warps wait for a stage, read their entries from shared memory and hand the stage back (one arrival per warp).  The stream keeps
// flowing across block boundaries and through the prologues.  Consumers synchronise among themselves with a named
// barrier (bar.sync 1, 256); the producer warp joins only the final reduction.
// ---------------------------------------------------------------------------------------
namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) { } }
__device__ __forceinline__ unsigned long long evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace tma

template <int NS, int GD>
__global__ void __launch_bounds__(288) k_spmv_jdst(int n_blocks, const int* __restrict__ rowbeg, const int* __restrict__ jbase,
                                                   const unsigned short* __restrict__ perm, const unsigned short* __restrict__ rlen,
                                                   const unsigned short* __restrict__ link,
                                                   const int* __restrict__ jdp, const int* __restrict__ jd,
                                                   const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                   const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                   const double* __restrict__ xin, double* __restrict__ out,
                                                   double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                   double* __restrict__ alpha_out, int wcap, int jcap) {
    // A stage holds a GROUP of GD consecutive diagonals of one block (at most GD x 512 entries: values + window positions).
    if (cgs->done) return;
    constexpr int R = 512, T = 256, SC = GD * R;
    __shared__ __align__(8) unsigned long long s_full[NS], s_empty[NS];
    extern __shared__ __align__(16) unsigned char s_tma[];                  // (a named array of its own: the alignment, and the compiler keeps the shared state space)
    double* s_val = reinterpret_cast<double*>(s_tma);                       // NS x SC values (16-byte aligned stages)
    unsigned short* s_col = (unsigned short*) (s_val + NS * SC);            // NS x SC window positions
    double* s_x = (double*) (s_col + NS * SC);                              // wcap window entries
    double* s_sum = s_x + wcap;                                             // 2 x R segment sums
    int* s_jd = (int*) (s_sum + 2 * R);                                     // jcap + 1 diagonal offsets (entries), padded to even
    unsigned short* s_link = (unsigned short*) (s_jd + ((jcap + 2) & ~1));  // 2 x R chain links
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { tma::mbar_init(&s_full[s], 1); tma::mbar_init(&s_empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc[2] = {0, 0};
    if (warp == 8) {
        // ---- producer warp: lane g looks up the bounds of group g, lane 0 keeps the ring full ----
        const unsigned long long pol = tma::evict_first_policy();
        int stage = 0; unsigned use = 0;                              // ring position; use = times the ring has wrapped
        for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
            const int base = __ldg(&jbase[b]);
            const int j0 = __ldg(&jdp[b]), maxlen = __ldg(&jdp[b + 1]) - j0 - 1;
            const int ng = (maxlen + GD - 1) / GD;                    // <= 32 groups: one lane each
            int lo = 0, hi = 0;
            if (lane < ng) { lo = __ldg(&jd[j0 + lane * GD]); hi = __ldg(&jd[j0 + min(maxlen, (lane + 1) * GD)]); }
            for (int g = 0; g < ng; ++g) {
                const int glo = __shfl_sync(0xffffffffu, lo, g), ghi = __shfl_sync(0xffffffffu, hi, g);
                if (lane == 0) {
                    if (use > 0) tma::mbar_wait(&s_empty[stage], (use - 1) & 1);      // the consumers have left the previous tenant
                    const unsigned cnt = (unsigned) (ghi - glo);
                    tma::mbar_expect_tx(&s_full[stage], cnt * 10u);
                    if (cnt) {
                        tma::bulk_g2s(s_val + (size_t) stage * SC, val + (size_t) base + glo, cnt * 8u, &s_full[stage], pol);
                        tma::bulk_g2s(s_col + (size_t) stage * SC, col16 + (size_t) base + glo, cnt * 2u, &s_full[stage], pol);
                    }
                }
                if (++stage == NS) { stage = 0; ++use; }
            }
        }
    } else {
        // ---- consumers: 8 warps, thread t owns the slots 2t and 2t + 1 of every block ----
        int stage = 0; unsigned phase = 0;
        int par = 0;
        for (int b = blockIdx.x; b < n_blocks; b += gridDim.x, par ^= 1) {
            const int r0 = __ldg(&rowbeg[b]);
            const int w0 = __ldg(&win_off[b]), nw = __ldg(&win_off[b + 1]) - w0;
            const int j0 = __ldg(&jdp[b]), nj = __ldg(&jdp[b + 1]) - j0;
            const size_t sl = (size_t) b * R + 2 * tid;
            const ushort2 ln = __ldg(reinterpret_cast<const ushort2*>(rlen + sl));
            const ushort2 pm = __ldg(reinterpret_cast<const ushort2*>(perm + sl));
            const int len0 = ln.x, len1 = ln.y;
            reinterpret_cast<ushort2*>(s_link + par * R)[tid] = __ldg(reinterpret_cast<const ushort2*>(link + sl));
            for (int i = tid; i < nw; i += T) s_x[i] = __ldg(&xin[__ldg(&win_list[w0 + i])]);
            for (int i = tid; i < nj; i += T) s_jd[i] = __ldg(&jd[j0 + i]);
            tma::consumer_bar();
            const int ng = (nj - 1 + GD - 1) / GD;
            double sum0 = 0, sum1 = 0;
            for (int g = 0; g < ng; ++g) {
                tma::mbar_wait(&s_full[stage], phase);
                const int gbase = s_jd[g * GD];
                const double* sv = s_val + (size_t) stage * SC + 2 * tid - gbase;
                const unsigned short* sc = s_col + (size_t) stage * SC + 2 * tid - gbase;
#pragma unroll
                for (int u = 0; u < GD; ++u) {
                    const int jj = g * GD + u;
                    if (jj < len0) {
                        const int o = s_jd[jj];
                        const double2 v = *reinterpret_cast<const double2*>(sv + o);
                        const ushort2 cc = *reinterpret_cast<const ushort2*>(sc + o);
                        sum0 += v.x * s_x[cc.x];
                        if (jj < len1) sum1 += v.y * s_x[cc.y];
                    }
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&s_empty[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1; }
            }
            double* ss = s_sum + par * R;
            const unsigned short* lk = s_link + par * R;
            if (pm.x & 0x8000) ss[2 * tid] = sum0;
            if (pm.y & 0x8000) ss[2 * tid + 1] = sum1;
            tma::consumer_bar();                                      // ends the block: window and offsets may be overwritten
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned short p = h ? pm.y : pm.x;
                if (p & 0x8000) continue;
                double sum = h ? sum1 : sum0;
                for (unsigned s = lk[2 * tid + h]; s != 0xFFFFu; s = lk[s]) sum += ss[s];
                const int row = r0 + p;
                out[row] = sum; acc[0] += __ldg(&xin[row]) * sum;
            }
        }
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<false>(cgs, tot, alpha_out);
    }
}

// ---------------------------------------------------------------------------------------
// Symmetric block-JDS SpMV: the matrix after the symmetric Dirichlet elimination (apply_boundary_values with
// eliminate_columns, DealSolver.cpp:439) is symmetric, so only its strictly lower triangle is stored and
// streamed -- half the bytes of k_spmv_jds.  Thread t owns two rows i of the block; for every stored entry
// a_ij (j < i) it adds a_ij x_j to its own sum (gather from the shared-memory window, as before) AND adds
// a_ij x_i to the shared-memory accumulator of column j (the transposed entry).  At the end of the block the
// accumulators -- window columns below the block and the block's own rows, the latter including gather sum
// and diagonal -- are added to the (pre-zeroed) output vector with FP64 reductions in L2 (red.global.add.f64),
// because later blocks scatter into the same rows.
//   d.h = sum_i d_i (a_ii d_i + 2 sum_{j<i} a_ij d_j)   (d'L'd = d'Ld), so the dot product needs only the gather.
// Summation order differs from the sequential CSR order and, through the reductions, from run to run
// (~1e-16 relative), far below the 1e-8 parity bar.
// Algorithmic bytes per launch of THIS formulation: 10 B per stored entry ((nnz - n) / 2 of them) + 8 n (diag)
// + 16 n (read d, accumulate h) + 4 n (row permutation / lengths).
// STATUS (B200, X mesh, profiles/r01e_*): correct, DRAM traffic 4.2 GB instead of 7.6 GB per launch, but 2.09 ms
// against 1.43 ms for k_spmv_jds: shared-memory FP64 atomics are CAS loops (ATOMS.CAST.SPIN) that serialise in every
// thread (+0.7 ms), and the lower-triangle row lengths (0..26) unbalance the jagged diagonals (barrier stalls; the
// gather alone, without any scatter, still takes 1.39 ms).  Kept as a selectable variant (option "spmv_kernel" 310 /
// 311), not the default; DESIGN.md section 3.5 lists what a faster version needs.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f64(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <int R>
__global__ void __launch_bounds__(R / 2, 1536 / R) k_spmv_sym(int n, int n_blocks, const int* __restrict__ jbase,
                                                    const unsigned short* __restrict__ perm, const unsigned short* __restrict__ rlen,
                                                    const int* __restrict__ jdp, const int* __restrict__ jd,
                                                    const unsigned short* __restrict__ col16, const double* __restrict__ val,
                                                    const double* __restrict__ diag,
                                                    const int* __restrict__ win_off, const int* __restrict__ win_list,
                                                    const double* __restrict__ xin, double* __restrict__ out,
                                                    double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                    double* __restrict__ alpha_out, int wcap, int jcap) {
    if (cgs->done) return;
    constexpr int T = R / 2;
    extern __shared__ double s_dyn[];
    double* s_x = s_dyn;                          // wcap + R input entries (window below the block, then the block's rows)
    double* s_y = s_dyn + wcap + R;               // matching accumulators of the transposed entries
    int* s_jd = (int*) (s_y + wcap + R);          // jcap + 1 diagonal offsets (in 2-entry units)
    const int tid = threadIdx.x;
    double acc[2] = {0, 0};
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int r0 = b * R;
        const int w0 = __ldg(&win_off[b]), nw = __ldg(&win_off[b + 1]) - w0;
        const int j0 = __ldg(&jdp[b]), nj = __ldg(&jdp[b + 1]) - j0;
        const long base = __ldg(&jbase[b]);
        const int sl0 = r0 + 2 * tid, sl1 = sl0 + 1;
        const int len0 = sl0 < n ? (int) __ldg(&rlen[sl0]) : 0, len1 = sl1 < n ? (int) __ldg(&rlen[sl1]) : 0;
        const int p0 = sl0 < n ? (int) __ldg(&perm[sl0]) : 0, p1 = sl1 < n ? (int) __ldg(&perm[sl1]) : 0;
        const double dg0 = sl0 < n ? __ldg(&diag[r0 + p0]) : 0.0, dg1 = sl1 < n ? __ldg(&diag[r0 + p1]) : 0.0;
        for (int i = tid; i < nw; i += T) { s_x[i] = __ldg(&xin[__ldg(&win_list[w0 + i])]); s_y[i] = 0.0; }
        for (int i = tid; i < R; i += T) { s_x[nw + i] = (r0 + i < n) ? __ldg(&xin[r0 + i]) : 0.0; s_y[nw + i] = 0.0; }
        for (int i = tid; i < nj; i += T) s_jd[i] = __ldg(&jd[j0 + i]) >> 1;
        __syncthreads();
        const double x0 = s_x[nw + p0], x1 = s_x[nw + p1];
        const double2* __restrict__ vb = reinterpret_cast<const double2*>(val + base) + tid;
        const ushort2* __restrict__ cb = reinterpret_cast<const ushort2*>(col16 + base) + tid;
        double sum0 = 0, sum1 = 0;
        double2 va[4], vb2[4];
        ushort2 ca[4], cb2[4];
#define FB_ISSUE(JJ, V, C)                                                                     \
        _Pragma("unroll") for (int u = 0; u < 4; ++u)                                          \
            if ((JJ) + u < len0) { const int o = s_jd[(JJ) + u]; V[u] = __ldcg(&vb[o]); C[u] = __ldcg(&cb[o]); }
#define FB_CONSUME(JJ, V, C)                                                                   \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                        \
            if ((JJ) + u < len0) { sum0 += V[u].x * s_x[C[u].x]; atomicAdd(&s_y[C[u].x], V[u].x * x0); } \
            if ((JJ) + u < len1) { sum1 += V[u].y * s_x[C[u].y]; atomicAdd(&s_y[C[u].y], V[u].y * x1); } \
        }
        FB_ISSUE(0, va, ca)
        for (int j = 0; j < len0; j += 8) {
            FB_ISSUE(j + 4, vb2, cb2)
            FB_CONSUME(j, va, ca)
            FB_ISSUE(j + 8, va, ca)
            FB_CONSUME(j + 4, vb2, cb2)
        }
#undef FB_ISSUE
#undef FB_CONSUME
        if (sl0 < n) { atomicAdd(&s_y[nw + p0], dg0 * x0 + sum0); acc[0] += x0 * (dg0 * x0 + 2.0 * sum0); }
        if (sl1 < n) { atomicAdd(&s_y[nw + p1], dg1 * x1 + sum1); acc[0] += x1 * (dg1 * x1 + 2.0 * sum1); }
        __syncthreads();
        for (int i = tid; i < nw; i += T) { const double v = s_y[i]; if (v != 0.0) red_add_f64(&out[__ldg(&win_list[w0 + i])], v); }
        for (int i = tid; i < R; i += T) if (r0 + i < n) red_add_f64(&out[r0 + i], s_y[nw + i]);
        __syncthreads();
    }
    double tot[2];
    if (reduce_publish<2>(acc, partial, counter, tot)) {
        cg_finish_spmv<false>(cgs, tot, alpha_out);
    }
}

// ---------------------------------------------------------------------------------------
// Persistent cooperative CG for the native (L2-resident) meshes: the WHOLE solve is one launch.
// One CTA per SM owns a contiguous, nnz-balanced slice of rows for the entire solve: its matrix
// values live in shared memory, its column indices in registers, its slices of x, g, d, h, 1/diag
// in shared memory.  Per iteration only the search direction d crosses SMs (written to global,
// gathered through L2 with ld.global.cg) plus 3 doubles per CTA for the dot products; the three
// phases are separated by grid-wide barriers and every warp re-reduces the per-CTA partials in a
// fixed order, so all CTAs take bit-identical alpha/beta/convergence decisions (no divergence at
// the barriers) and the result is deterministic.  Same operation order as the multi-kernel path
// (deal.II SolverCG): h = A d; alpha = gh/(d.h); x += alpha d; g += alpha h; test |g|;
// beta = g.Dinv g / gh; d = beta d - Dinv g.
// ---------------------------------------------------------------------------------------
// Grid-wide all-reduce through L2: every CTA publishes its partial sums, then arrives on ONE monotonic counter
// with a release reduction (red.release.gpu: the partials and -- through the preceding bar.sync -- the d slice
// written by the other threads of the CTA are visible to whoever observes the count); one lane per CTA polls the
// counter with acquire loads until all G CTAs of this phase have arrived, then warp 0 adds the G partials in index
// order -> bit-identical totals in all CTAs.  (A first version polled one flag per CTA: G x G loads per phase on
// five cache lines of one L2 slice cost ~9000 cycles per barrier; the single counter costs 148 loads per poll round.)
// Partials are double-buffered by sequence parity: a CTA can run at most one phase ahead of the slowest one, so a
// slot is never overwritten while somebody still reads it.
__device__ __forceinline__ void red_release_gpu_inc(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

struct GridComm {
    double* part;        // [2][NVMAX][G]
    unsigned* counter;   // arrivals since the launch (zeroed by the host)
    int G; unsigned seq;
    long long* dbg;
};
constexpr int PERS_NV = 2;

template <int NV>
__device__ __forceinline__ void grid_allreduce(GridComm& gc, double (&v)[NV], double (*s_red)[32], double (&tot)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    ++gc.seq;
    double* slot = gc.part + (size_t) (gc.seq & 1) * PERS_NV * gc.G;
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (gc.dbg) t0 = clock64();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[k][warp] = x;
    }
    __syncthreads();            // also orders this CTA's global stores (d slice) before the arrival below
    if (warp == 0) {
        if (gc.dbg) t1 = clock64();
#pragma unroll
        for (int k = 0; k < NV; ++k) {       // warp 0 adds the per-warp partials (fixed order)
            double t = (lane < nwarp) ? s_red[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) __stcg(&slot[(size_t) k * gc.G + blockIdx.x], t);
        }
        if (lane == 0) {
            red_release_gpu_inc(gc.counter);
            if (gc.dbg) t2 = clock64();
            const unsigned target = gc.seq * (unsigned) gc.G;
            while (ld_acquire_gpu(gc.counter) < target) { }
        }
        __syncwarp();
        if (gc.dbg) t3 = clock64();
        // all partials of all NV sums are requested before the first add (G <= 160: five per lane and sum), so the
        // read costs one L2 round trip instead of one per 32 CTAs
        double pv[NV][5];
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int i = lane + 32 * j;
                pv[k][j] = (i < gc.G) ? __ldcg(&slot[(size_t) k * gc.G + i]) : 0.0;
            }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = ((pv[k][0] + pv[k][1]) + (pv[k][2] + pv[k][3])) + pv[k][4];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_red[k][0] = x;
        }
        if (gc.dbg) t4 = clock64();
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = s_red[k][0];
    __syncthreads();            // s_red is reused by the next call
    if (gc.dbg && threadIdx.x == 0 && blockIdx.x == 0) {
        const long long t5 = clock64();
        gc.dbg[24] += t1 - t0; gc.dbg[25] += t2 - t1; gc.dbg[26] += t3 - t2; gc.dbg[27] += t4 - t3; gc.dbg[28] += t5 - t4; gc.dbg[29] += 1;
    }
}

template <int THREADS, int PT>
__global__ void __launch_bounds__(THREADS, 1) k_cg_persistent(const int* __restrict__ cta_row, const int* __restrict__ rowptr,
                                                              const int* __restrict__ col, const double* __restrict__ val,
                                                              const double* __restrict__ rhs, const double* __restrict__ dinv_g,
                                                              double* x_g, double* z_g, double* partial, int* flags, CgScalars* cgs,
                                                              int cap, int rmax, long long* dbg) {
    // TWO grid-wide barriers per iteration.  What crosses SMs is z = Dinv g (published by the owner of a row right
    // after the update, i.e. BEFORE the all-reduce that yields beta); every CTA keeps the search direction at the
    // column of each of its non-zeros in shared memory (s_dc) and advances it itself,  d_j <- beta d_j - z_j,  with
    // the very same two instructions (__dmul_rn, __fma_rn) the owner uses, so all copies of d_j are bit-identical.
    // A third barrier ("every slice of the new d is visible") is therefore not needed.
    extern __shared__ double smem[];
    long long t_ph[6] = {0, 0, 0, 0, 0, 0}, t_last = clock64();
    auto lap = [&](int k) { if (dbg) { const long long t = clock64(); t_ph[k] += t - t_last; t_last = t; } };
    double* s_val = smem;
    double* s_dc = s_val + cap;          // d (at start: x) at the column of every non-zero of the slice
    double* s_x = s_dc + cap;
    double* s_g = s_x + rmax;
    double* s_d = s_g + rmax;
    double* s_h = s_d + rmax;
    double* s_dinv = s_h + rmax;
    int* s_rp = (int*) (s_dinv + rmax);
    __shared__ double s_red[PERS_NV][32];
    const int tid = threadIdx.x;
    GridComm gc = {partial, (unsigned*) flags, (int) gridDim.x, 0u, dbg};
    const int r0 = cta_row[blockIdx.x], nr = cta_row[blockIdx.x + 1] - r0;
    const int k0 = rowptr[r0], cnt = rowptr[r0 + nr] - k0;

    int cidx[PT];
#pragma unroll
    for (int u = 0; u < PT; ++u) {
        const int k = tid + u * THREADS;
        cidx[u] = (k < cnt) ? __ldg(&col[k0 + k]) : -1;
        if (k < cnt) { s_val[k] = __ldg(&val[k0 + k]); s_dc[k] = x_g[cidx[u]]; }
    }
    for (int r = tid; r <= nr; r += THREADS) s_rp[r] = rowptr[r0 + r] - k0;
    for (int r = tid; r < nr; r += THREADS) { s_x[r] = x_g[r0 + r]; s_dinv[r] = dinv_g[r0 + r]; s_d[r] = 0.0; }
    __syncthreads();

    // s_h = A * (vector held in s_dc) on the CTA's rows.  Four lanes add one row (fixed order: lane-strided partial
    // sums, then a 4-lane butterfly), so the ~150-entry rows of the tet vertices do not serialise the phase.
    auto rowsums = [&]() {
        const int sub = tid & 3;
        for (int rb = 0; rb < nr; rb += THREADS / 4) {          // block-uniform trip count
            const int r = rb + (tid >> 2);
            double sum = 0;
            if (r < nr)
                for (int j = s_rp[r] + sub; j < s_rp[r + 1]; j += 4) sum += s_val[j] * s_dc[j];
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (r < nr && sub == 0) s_h[r] = sum;
        }
        __syncthreads();
    };

    const double tol2 = cgs->tol2;
    const int max_iter = cgs->max_iter;
    // ---- start: g = A x - b (deal.II SolverCG), z = Dinv g published ----
    rowsums();
    double acc[2] = {0, 0}, tot[2];
    for (int r = tid; r < nr; r += THREADS) {
        const double g = s_dinv[r] != 0.0 ? s_h[r] - rhs[r0 + r] : 0.0;      // constrained rows: see k_bc_prepare
        s_g[r] = g;
        __stcg(&z_g[r0 + r], __dmul_rn(s_dinv[r], g));
        acc[0] += g * g * s_dinv[r];
        acc[1] += g * g;
    }
    grid_allreduce<2>(gc, acc, s_red, tot);       // also: every slice of z is visible
    double gh = tot[0], res2 = tot[1], beta = 0.0;    // first direction: d = -z  (beta = 0; x, the old content of s_dc, is finite)
    int it = 0;
    int done = (res2 <= tol2) ? 1 : ((max_iter <= 0 || res2 != res2) ? 2 : 0);
    while (!done) {
        // phase A: d = beta d - z (own rows and the column copies), h = A d, alpha = gh / (d.h)
        lap(5);
#pragma unroll
        for (int u = 0; u < PT; ++u)
            if (cidx[u] >= 0) { const int k = tid + u * THREADS; s_dc[k] = __fma_rn(beta, s_dc[k], -__ldcg(&z_g[cidx[u]])); }
        for (int r = tid; r < nr; r += THREADS) s_d[r] = __fma_rn(beta, s_d[r], -__dmul_rn(s_dinv[r], s_g[r]));
        __syncthreads();
        rowsums();
        double a[1] = {0}, at[1];
        for (int r = tid; r < nr; r += THREADS) a[0] += s_d[r] * s_h[r];
        lap(0);
        grid_allreduce<1>(gc, a, s_red, at);      // also: every CTA has finished reading z
        lap(1);
        // phase B: x += alpha d, g += alpha h, z = Dinv g published, |g|, g.Dinv g
        const double alpha = gh / at[0];
        acc[0] = 0; acc[1] = 0;
        for (int r = tid; r < nr; r += THREADS) {
            s_x[r] += alpha * s_d[r];
            const double g = s_dinv[r] != 0.0 ? s_g[r] + alpha * s_h[r] : 0.0;
            s_g[r] = g;
            __stcg(&z_g[r0 + r], __dmul_rn(s_dinv[r], g));
            acc[0] += g * g * s_dinv[r];
            acc[1] += g * g;
        }
        lap(2);
        grid_allreduce<2>(gc, acc, s_red, tot);
        lap(3);
        // convergence test (deal.II SolverControl: success first), beta
        res2 = tot[1];
        ++it;
        beta = tot[0] / gh;
        gh = tot[0];
        if (res2 <= tol2) done = 1;
        else if (it >= max_iter || res2 != res2) done = 2;
        lap(4);
    }
    if (dbg && tid == 0 && blockIdx.x < 4) for (int k = 0; k < 6; ++k) dbg[6 * blockIdx.x + k] = t_ph[k];
    for (int r = tid; r < nr; r += THREADS) x_g[r0 + r] = s_x[r];
    if (blockIdx.x == 0 && tid == 0) { cgs->gh = gh; cgs->res2 = res2; cgs->it = it; cgs->done = done; }
}

// DealSolver::export_solution_grad (DealSolver.cpp:280-301): for vertex v, MINUS the gradient of the solution at Gauss
// point number vertex2node[v] of cell vertex2cell[v] (FEValues::get_function_gradients at the QGauss<3>(2) points,
// lexicographic order) -- the reference indexes the quadrature points with the local vertex number.
__global__ void __launch_bounds__(128) k_solution_grad(int n_vert, const int* __restrict__ lastcell, const int* __restrict__ cells,
                                                       const double* __restrict__ vxyz, const double* __restrict__ x,
                                                       double* __restrict__ grad3) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vert) return;
    const int code = lastcell[v], c = code >> 3, q = code & 7;
    double X[8], Y[8], Z[8], phi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = cells[8 * (size_t) c + i];
        X[i] = vxyz[3 * (size_t) d]; Y[i] = vxyz[3 * (size_t) d + 1]; Z[i] = vxyz[3 * (size_t) d + 2]; phi[i] = x[d];
    }
    const double ga = 0.5 * (1.0 - 0.57735026918962576451), gb = 0.5 * (1.0 + 0.57735026918962576451);
    const double xi = (q & 1) ? gb : ga, eta = (q & 2) ? gb : ga, zeta = (q & 4) ? gb : ga;
    double dN[8][3], J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double fx = (i & 1) ? xi : 1.0 - xi, fy = (i & 2) ? eta : 1.0 - eta, fz = (i & 4) ? zeta : 1.0 - zeta;
        const double sx = (i & 1) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0, sz = (i & 4) ? 1.0 : -1.0;
        dN[i][0] = sx * fy * fz; dN[i][1] = fx * sy * fz; dN[i][2] = fx * fy * sz;
#pragma unroll
        for (int e = 0; e < 3; ++e) { J[0][e] += X[i] * dN[i][e]; J[1][e] += Y[i] * dN[i][e]; J[2][e] += Z[i] * dN[i][e]; }
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][0] * J[2][2] - J[1][2] * J[2][0], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double id = 1.0 / (J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02);
    double inv[3][3];
    inv[0][0] = c00 * id;  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    inv[1][0] = -c01 * id; inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    inv[2][0] = c02 * id;  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    double g[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int d = 0; d < 3; ++d) g[d] += phi[i] * (dN[i][0] * inv[0][d] + dN[i][1] * inv[1][d] + dN[i][2] * inv[2][d]);
    grad3[3 * (size_t) v] = -g[0]; grad3[3 * (size_t) v + 1] = -g[1]; grad3[3 * (size_t) v + 2] = -g[2];
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

// row-block parameters (non-zeros, rows) of a streaming-kernel variant
void stream_block_shape(int kernel, int& chunk, int& maxrows) {
    switch (kernel) {
        case 100: chunk = 1024; maxrows = 256; break;
        case 101: chunk = 4096; maxrows = 512; break;
        case 102: chunk = 1024; maxrows = 128; break;
        case 103: chunk = 4096; maxrows = 256; break;
        case 200: chunk = 2048; maxrows = 256; break;
        case 201: chunk = 4096; maxrows = 256; break;
        case 202: chunk = 4096; maxrows = 512; break;
        case 203: chunk = 1024; maxrows = 256; break;
        case 204: chunk = 8192; maxrows = 512; break;
        default: chunk = 2048; maxrows = 256; break;
    }
}

int choose_lanes(const fb_ctx* c) {
    // 0 selects the row-block streaming kernel (option "spmv_kernel": -1 auto, 0 stream, else lanes per row)
    if (c->spmv_kernel >= 0) return (c->world > 1 && c->spmv_kernel >= 310) ? 304 : c->spmv_kernel;
    if (c->nnz >= 4000000) return 306;        // segmented block-JDS (long rows split), 512 slots per block, evict-first matrix stream
    const double avg = c->n_dofs ? (double) c->nnz / c->n_dofs : 1.0;
    if (avg > 48) return 32;
    if (avg > 20) return 8;
    if (avg > 10) return 4;
    return 2;
}

// scatter map of the assembly kernels: once per mesh (kept across imports with unchanged topology)
static bool ensure_asm_map(fb_ctx* c) {
    if (c->asm_map_ready) return true;
    if (c->asm_map_opt == 0) return false;
    if (c->d_asm_map.alloc(64 * (size_t) c->n_cells) != cudaSuccess) { cudaGetLastError(); return false; }      // no memory: row-walk kernels
    k_build_asm_map<<<(unsigned) ((8L * c->n_cells + 255) / 256), 256, 0, c->stream>>>(c->n_cells, c->n_dofs, c->d_cells.p, c->d_rowptr.p, c->d_col.p,
                                                                                   c->d_asm_map.p);
    c->launches++;
    c->asm_map_ready = true;
    return true;
}

void launch_assemble_stiffness(fb_ctx* c) {
    if (c->imported_degree == 2) { launch_q2_stiffness(c); return; }
    if (ensure_asm_map(c))
        k_assemble_stiffness_mapped<<<(c->n_cells + ASM_BLOCK - 1) / ASM_BLOCK, ASM_BLOCK, 0, c->stream>>>(c->n_cells, c->d_cells.p, c->d_vxyz.p,
                                                                                                      c->d_asm_map.p, c->d_val_save.p);
    else
        k_assemble_stiffness<<<(c->n_cells + ASM_BLOCK - 1) / ASM_BLOCK, ASM_BLOCK, 0, c->stream>>>(
            c->n_cells, c->n_dofs, c->d_cells.p, c->d_vxyz.p, c->d_rowptr.p, c->d_col.p, c->d_val_save.p);
    c->launches++;
}

// charge_density = rhs (Neumann faces + space charge, before the Dirichlet conditions) / dof_volume, into c->d_rho
void launch_charge_density(fb_ctx* c, double* d_scratch) {
    cudaMemsetAsync(d_scratch, 0, (size_t) c->n_dofs * sizeof(double), c->stream);
    k_dof_volumes<<<(c->n_cells + ASM_BLOCK - 1) / ASM_BLOCK, ASM_BLOCK, 0, c->stream>>>(c->n_cells, c->n_dofs, c->d_cells.p, c->d_vxyz.p, d_scratch);
    k_charge_density<<<grid_for(c, c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_rhs.p, d_scratch, c->d_rho.p);
    c->launches += 2;
}

void launch_cell_volumes(fb_ctx* c, double* d_cell_vol) {
    k_cell_volumes<<<(c->n_cells + ASM_BLOCK - 1) / ASM_BLOCK, ASM_BLOCK, 0, c->stream>>>(c->n_cells, c->d_cells.p, c->d_vxyz.p, d_cell_vol);
    c->launches++;
}

void launch_neumann(fb_ctx* c) {
    if (c->n_top_faces == 0) return;
    if (c->imported_degree == 2) { launch_q2_neumann(c); return; }
    k_neumann_faces<<<(c->n_top_faces + 127) / 128, 128, 0, c->stream>>>(c->n_top_faces, c->n_dofs, c->d_topfaces.p, c->d_vxyz.p,
                                                                       c->applied_field, c->mesh_kind ? c->d_face_bc.p : nullptr, c->d_rhs.p);
    c->launches++;
}

void launch_assemble_heat(fb_ctx* c, double gamma, const double* d_T_prev, const double* d_phi) {
    k_assemble_heat<<<(c->n_cells + ASM_BLOCK - 1) / ASM_BLOCK, ASM_BLOCK, 0, c->stream>>>(
        c->n_cells, c->n_dofs, c->d_cells.p, c->d_vxyz.p, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, c->d_rhs.p, d_T_prev, d_phi, gamma,
        c->d_res_T.p, c->d_res_rho.p, c->ch_n_table, c->ch_lorentz, ensure_asm_map(c) ? c->d_asm_map.p : nullptr);
    c->launches++;
}

void launch_set_bc(fb_ctx* c, const int* d_dofs, int n, double value) {
    if (n == 0) return;
    k_set_bc<<<(n + 255) / 256, 256, 0, c->stream>>>(n, d_dofs, value, c->d_bcflag.p, c->d_bcval.p);
    c->launches++;
}

void launch_bc_prepare(fb_ctx* c) {          // dinv (0 on constrained rows), diagpos
    k_bc_prepare<<<grid_for(c, c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, c->d_bcflag.p,
                                                                     c->d_dinv.p, c->d_diagpos.p);
    c->launches++;
}

// test hook (fb_get_system): the system as the reference holds it after apply_boundary_values -- rows and columns of
// the constrained dofs eliminated, right-hand side lifted.  Not used by the solver.
void launch_materialize_eliminated(fb_ctx* c, double* d_val_out, double* d_rhs_out, double* d_lift, double* d_diag_inv, int* d_diagpos_tmp) {
    const int g = grid_for(c, (long) c->n_dofs * 8, 256);
    k_apply_bc_matrix<8><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, d_val_out,
                                                   c->d_bcflag.p, c->d_bcval.p, d_lift, d_diag_inv, d_diagpos_tmp);
    k_eliminated_rhs<<<grid_for(c, c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_bcflag.p, c->d_bcval.p, d_diag_inv, d_lift, c->d_rhs.p, d_rhs_out);
    c->launches += 2;
}

void launch_csr_to_jds(fb_ctx* c) {
    const int g = grid_for(c, (long) c->n_dofs * 8, 256);
    if (c->jds_split > 0) {
        k_csr_to_jds_split<<<g, 256, 0, c->stream>>>(c->n_dofs, c->jds_R, c->jds_nb, c->jds_split, c->d_jds_rowbeg.p, c->d_rowptr.p, c->d_jds_slot.p,
                                                     c->d_jds_link.p, c->d_jds_base.p, c->d_jds_jdp.p, c->d_jds_jd.p, c->d_val_save.p, c->d_val_jds.p);
        c->launches++;
        return;
    }
    k_csr_to_jds<<<g, 256, 0, c->stream>>>(c->n_dofs, c->jds_R, c->d_rowptr.p, c->d_jds_slot.p, c->d_jds_base.p, c->d_jds_jdp.p, c->d_jds_jd.p,
                                           c->d_val_save.p, c->d_val_jds.p, c->jds_sym ? c->d_diagpos.p : nullptr, c->d_diag.p);
    c->launches++;
}

void launch_bc_solution(fb_ctx* c) {
    k_bc_solution<<<grid_for(c, c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_bcflag.p, c->d_bcval.p, c->d_rhs.p, c->d_x.p);
    c->launches++;
}

template <bool INIT>
static void spmv_dispatch(fb_ctx* c, int lanes, const double* xin, double* out, double* alpha) {
    unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
    double* part = c->d_partial.p;
    if (lanes >= 310) {         // symmetric block-JDS kernel (310: 512 rows per block, 311: 256); accumulates into a zeroed `out`
        if (INIT) { spmv_dispatch<true>(c, 8, xin, out, alpha); return; }      // the initial residual (once per solve) streams the CSR copy
        const int nb = c->jds_nb;
        const size_t smem = 2 * sizeof(double) * ((size_t) c->win_cap + c->jds_R) + sizeof(int) * ((size_t) c->jds_maxlen + 2);
#define FB_SYM(RR, OCC) do {                                                                                                        \
        auto kern = k_spmv_sym<RR>;                                                                                                 \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);                                        \
        const int occ = std::max(1, std::min((OCC), (int) (200 * 1024 / (smem + 1024))));                                           \
        const int g = std::min(nb, c->n_sm * occ);                                                                                  \
        kern<<<g, (RR) / 2, smem, c->stream>>>(c->n_dofs, nb, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_jdp.p, c->d_jds_jd.p, \
                                         c->d_col16.p, c->d_val_jds.p, c->d_diag.p, c->d_win_off.p, c->d_win_list.p, xin,           \
                                         out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen); } while (0)
        if (c->jds_R == 256) FB_SYM(256, 12); else FB_SYM(512, 6);
#undef FB_SYM
        c->launches++;
        return;
    }
    if (lanes >= 300) {         // block-JDS kernel (300: 256 rows per block, 301: 128)
        const int nb = c->jds_nb;
        const size_t smem = sizeof(double) * (size_t) c->win_cap + sizeof(int) * ((size_t) c->jds_maxlen + 2);
#define FB_JDS(RR, OCC) do {                                                                                                        \
        auto kern = k_spmv_jds<INIT, RR, 0>;                                                                                           \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);                                        \
        const int occ = std::max(1, std::min((OCC), (int) (200 * 1024 / (smem + 1024))));                                           \
        const int g = std::min(nb, c->n_sm * occ);                                                                                  \
        kern<<<g, (RR) / 2, smem, c->stream>>>(c->n_dofs, nb, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_jdp.p, c->d_jds_jd.p, \
                                         c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin, c->d_rhs.p, c->d_dinv.p, \
                                         out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen); } while (0)
        if (lanes == 308 && !INIT && c->jds_pad == 8 && c->jds_maxlen <= 128) {   // segmented layout, matrix stream through a TMA-fed shared-memory ring
            constexpr int NS = 3, GD = 4;
            const size_t smem8 = 16 + (size_t) NS * GD * 512 * 10 + sizeof(double) * ((size_t) c->win_cap + 2 * 512)
                                 + sizeof(int) * (((size_t) c->jds_maxlen + 2 + 1) & ~(size_t) 1) + 2 * 512 * sizeof(unsigned short);
            auto kern = k_spmv_jdst<NS, GD>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem8);
            const int occ = std::max(1, std::min(c->spmv_occ, (int) (224 * 1024 / (smem8 + 2048))));
            const int g = std::min(nb, c->n_sm * occ);
            kern<<<g, 288, smem8, c->stream>>>(nb, c->d_jds_rowbeg.p, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_link.p, c->d_jds_jdp.p,
                                               c->d_jds_jd.p, c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin,
                                               out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen);
        } else
        if (lanes == 307 && !INIT) {   // segmented layout + index list of the next block prefetched
            const size_t smem7 = sizeof(double) * ((size_t) c->win_cap + 2 * 512) + sizeof(int) * ((((size_t) c->jds_maxlen + 2 + 1) & ~(size_t) 1) + 2 * (size_t) c->win_cap)
                                 + 2 * 512 * sizeof(unsigned short);
            auto kern = k_spmv_jdsq<512>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem7);
            const int occ = std::max(1, std::min(c->spmv_occ, (int) (200 * 1024 / (smem7 + 1024))));
            const int g = std::min(nb, c->n_sm * occ);
            kern<<<g, 256, smem7, c->stream>>>(nb, c->d_jds_rowbeg.p, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_link.p, c->d_jds_jdp.p,
                                               c->d_jds_jd.p, c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin,
                                               out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen);
        } else
        if (lanes >= 306 && lanes <= 308) {            // segmented layout (long rows split), evict-first matrix stream
            const size_t smem6 = sizeof(double) * ((size_t) c->win_cap + 2 * 512) + sizeof(int) * (((size_t) c->jds_maxlen + 2 + 1) & ~(size_t) 1) + 2 * 512 * sizeof(unsigned short);
            auto kern = k_spmv_jdss<INIT, 512>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem6);
            const int occ = std::max(1, std::min(c->spmv_occ, (int) (200 * 1024 / (smem6 + 1024))));
            const int g = std::min(nb, c->n_sm * occ);
            kern<<<g, 256, smem6, c->stream>>>(nb, c->d_jds_rowbeg.p, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_link.p, c->d_jds_jdp.p,
                                               c->d_jds_jd.p, c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin, c->d_rhs.p, c->d_dinv.p,
                                               out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen);
        } else
        if (lanes == 305) {            // evict-first matrix stream, loads pinned in program order (volatile asm)
            auto kern = k_spmv_jds<INIT, 512, 2>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            const int occ = std::max(1, std::min(c->spmv_occ, (int) (200 * 1024 / (smem + 1024))));
            const int g = std::min(nb, c->n_sm * occ);
            kern<<<g, 256, smem, c->stream>>>(c->n_dofs, nb, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_jdp.p, c->d_jds_jd.p,
                                              c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin, c->d_rhs.p, c->d_dinv.p,
                                              out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen);
        } else
        if (lanes == 304) {            // evict-first matrix stream
            auto kern = k_spmv_jds<INIT, 512, 1>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            const int occ = std::max(1, std::min(c->spmv_occ, (int) (200 * 1024 / (smem + 1024))));
            const int g = std::min(nb, c->n_sm * occ);
            kern<<<g, 256, smem, c->stream>>>(c->n_dofs, nb, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_jdp.p, c->d_jds_jd.p,
                                              c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin, c->d_rhs.p, c->d_dinv.p,
                                              out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen);
        } else
        if (lanes == 303) {            // prefetching variant: two window / offset buffers
            const size_t jstride = ((size_t) c->jds_maxlen + 2 + 1) & ~(size_t) 1;
            const size_t smem2 = 2 * sizeof(double) * (size_t) c->win_cap + 2 * sizeof(int) * jstride;
            auto kern = k_spmv_jdsp<INIT, 512>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
            const int occ = std::max(1, std::min(c->spmv_occ, (int) (200 * 1024 / (smem2 + 1024))));
            const int g = std::min(nb, c->n_sm * occ);
            kern<<<g, 256, smem2, c->stream>>>(c->n_dofs, nb, c->d_jds_base.p, c->d_jds_perm.p, c->d_jds_len.p, c->d_jds_jdp.p, c->d_jds_jd.p,
                                              c->d_col16.p, c->d_val_jds.p, c->d_win_off.p, c->d_win_list.p, xin, c->d_rhs.p, c->d_dinv.p,
                                              out, part, counter, c->d_cg.p, alpha, c->win_cap, c->jds_maxlen);
        } else
        if (c->jds_R == 128) FB_JDS(128, 24); else if (c->jds_R == 512) FB_JDS(512, 6); else FB_JDS(256, 12);
#undef FB_JDS
        c->launches++;
        return;
    }
    if (lanes >= 200) {         // windowed streaming kernel (variants 200.. are tuning points)
        const int nb = c->n_rowblk;
#define FB_WINDOW(T, P, OCC) do {                                                                                                   \
        auto kern = k_spmv_window<INIT, T, P>;                                                                                      \
        const size_t smem = sizeof(double) * ((size_t) (T) * (P) + (size_t) c->win_cap);                                            \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);                                        \
        const int occ = std::max(1, std::min((OCC), (int) (220 * 1024 / (smem + 4 * ((T) + 1) + 1024))));                           \
        const int g = std::min(nb, c->n_sm * occ);                                                                                  \
        kern<<<g, T, smem, c->stream>>>(nb, c->d_rowblk.p, c->d_rowptr.p, c->d_col16.p, c->d_val_save.p, c->d_win_off.p, c->d_win_list.p, \
                                        xin, c->d_rhs.p, c->d_dinv.p, out, part, counter, c->d_cg.p, alpha, c->win_cap); } while (0)
        switch (lanes) {
            case 201: FB_WINDOW(256, 16, 8); break;
            case 202: FB_WINDOW(512, 8, 4); break;
            case 203: FB_WINDOW(256, 4, 8); break;
            case 204: FB_WINDOW(512, 16, 4); break;
            default: FB_WINDOW(256, 8, 8); break;
        }
#undef FB_WINDOW
        c->launches++;
        return;
    }
    if (lanes == 0 || lanes >= 100) {           // row-block streaming kernel (variants 100.. are tuning points)
        const int nb = c->n_rowblk;
#define FB_STREAM(T, P, OCC) do { const int g = std::min(nb, c->n_sm * (OCC));                                                     \
        k_spmv_stream<INIT, T, P><<<g, T, 0, c->stream>>>(nb, c->d_rowblk.p, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, xin, c->d_rhs.p, \
                                                         c->d_dinv.p, out, part, counter, c->d_cg.p, alpha); } while (0)
        switch (lanes) {
            case 100: FB_STREAM(256, 4, 8); break;
            case 101: FB_STREAM(512, 8, 4); break;
            case 102: FB_STREAM(128, 8, 16); break;
            case 103: FB_STREAM(256, 16, 6); break;
            default: FB_STREAM(256, 8, 8); break;
        }
#undef FB_STREAM
        c->launches++;
        return;
    }
    const int g = grid_for(c, (long) c->n_dofs * lanes, 256);
#define FB_SPMV(L) k_spmv_dot<L, INIT><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, xin, \
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

// steps 1 .. k-1 of the Chebyshev recurrence on the current (p, z): SpMV w = A p into d_h (free once k_update has read
// it), then k_cheb_step; r_0 = g is read in place, later residuals live in d_cheb_r
static void launch_cheb_steps(fb_ctx* c, bool init) {
    unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
    const int g = grid_for(c, c->n_dofs, 256);
    const int k = c->cheb_k;
    for (int i = 1; i < k; ++i) {
        spmv_dispatch<false>(c, c->cheb_lanes, c->d_cheb_p.p, c->d_h.p, (double*) (c->d_cg.p + 1) + 3);     // alpha slot 3 = scratch
        const double* r_in = (i == 1) ? c->d_g.p : c->d_cheb_r.p;
        const double c1 = c->cheb_c1[i], c2 = c->cheb_c2[i];
#define FB_CHEB(LAST, INIT) k_cheb_step<LAST, INIT><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_h.p, r_in, c->d_cheb_r.p, c->d_dinv.p, c->d_cheb_p.p, \
                                c->d_z.p, c->d_g.p, c1, c2, c->d_partial.p, counter, c->d_cg.p, (double*) (c->d_cg.p + 1) + 1)
        if (i < k - 1) FB_CHEB(false, false); else if (init) FB_CHEB(true, true); else FB_CHEB(true, false);
#undef FB_CHEB
        c->launches++;
    }
}

// Gershgorin bound of Dinv A (once per assembly) and the recurrence coefficients
cudaError_t cheb_prepare(fb_ctx* c, int lanes) {
    const int k = std::max(2, std::min(16, c->cheb_degree));
    cudaError_t e = c->d_cheb_p.alloc(c->n_cols);
    if (e == cudaSuccess) e = c->d_cheb_r.alloc(c->n_cols);
    if (e != cudaSuccess) return e;
    if (c->cheb_lmax <= 0) {
        unsigned long long* bits = (unsigned long long*) c->d_minmax.p;
        cudaMemsetAsync(bits, 0, sizeof(unsigned long long), c->stream);
        k_gershgorin<<<grid_for(c, (long) c->n_dofs * 8, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_rowptr.p, c->d_val_save.p, c->d_dinv.p, bits);
        c->launches++;
        double lmax = 0;
        e = cudaMemcpyAsync(&lmax, bits, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return e;
        c->cheb_gershgorin = lmax;
        if (c->cheb_power_iters > 0) {
            // sharpen: power iteration (lower bound) x 1.2, never above the Gershgorin bound
            unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
            const int g = grid_for(c, c->n_dofs, 256);
            const int pl = (lanes >= 310) ? 304 : lanes;
            k_power_init<<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_dinv.p, c->d_cheb_p.p);
            for (int i = 0; i < c->cheb_power_iters; ++i) {
                spmv_dispatch<false>(c, pl, c->d_cheb_p.p, c->d_h.p, (double*) (c->d_cg.p + 1) + 3);
                k_power_step<<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_h.p, c->d_dinv.p, c->d_cheb_p.p, c->d_partial.p, counter, c->d_minmax.p);
            }
            c->launches += 1 + c->cheb_power_iters;
            double q[2] = {0, 0};
            e = cudaMemcpyAsync(q, c->d_minmax.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return e;
            if (q[1] > 0 && q[0] > 0) lmax = std::min(lmax, 1.2 * q[0] / q[1]);
        }
        c->cheb_lmax = lmax;
        if (getenv("FB_VERBOSE")) fprintf(stderr, "[fb] Chebyshev: lmax(Dinv A) = %.4f (Gershgorin %.4f)\n", c->cheb_lmax, c->cheb_gershgorin);
    }
    const double lmax = c->cheb_lmax, lmin = lmax / std::max(1.5, c->cheb_ratio);
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    c->cheb_inv_theta = 1.0 / theta;
    double rho = 1.0 / sigma;
    for (int i = 1; i < k; ++i) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        c->cheb_c1[i] = rho_new * rho; c->cheb_c2[i] = 2.0 * rho_new / delta;
        rho = rho_new;
    }
    c->cheb_k = k;
    c->cheb_lanes = (lanes >= 310) ? 304 : lanes;       // the symmetric layout accumulates into a zeroed vector: not used here
    return cudaSuccess;
}

void launch_cg_init(fb_ctx* c, int lanes) {
    if (c->cheb_active) {
        spmv_dispatch<true>(c, lanes, c->d_x.p, c->d_g.p, nullptr);
        const int g = grid_for(c, c->n_dofs, 256);
        k_cheb_start<<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_cheb_p.p, c->d_z.p, c->cheb_inv_theta, c->d_cg.p);
        c->launches++;
        launch_cheb_steps(c, true);
        k_direction_z<true><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_z.p, c->d_d.p, c->d_cg.p, nullptr);
        c->launches++;
        return;
    }
    if (c->tl_active) {          // two-level preconditioner: the INIT SpMV leaves g, gh = g.Dinv g, |g|; twolevel.cu adds the coarse part
        c->h_needs_zero = false;
        spmv_dispatch<true>(c, lanes, c->d_x.p, c->d_g.p, nullptr);
        launch_tl_init_tail(c);
        return;
    }
    c->h_needs_zero = lanes >= 310;
    if (c->h_needs_zero) cudaMemsetAsync(c->d_h.p, 0, (size_t) c->n_dofs * sizeof(double), c->stream);
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
    const int gu = std::min(g, c->n_sm * 6);          // k_update: 6 resident blocks per SM, one wave
    if (c->tl_active) { launch_tl_vectors(c); return; }
    if (c->cheb_active) {       // Chebyshev-preconditioned iteration: update (+ first step), k-1 x (SpMV + step), direction
        k_update<false, true><<<gu, 256, 0, c->stream>>>(c->n_dofs, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter,
                                                         c->d_cg.p, alpha_ptr(c), beta_ptr(c), c->d_cheb_p.p, c->d_z.p, c->cheb_inv_theta);
        c->launches++;
        launch_cheb_steps(c, false);
        k_direction_z<false><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_z.p, c->d_d.p, c->d_cg.p, beta_ptr(c));
        c->launches++;
        return;
    }
    if (c->h_needs_zero)
        k_update<true, false><<<gu, 256, 0, c->stream>>>(c->n_dofs, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter,
                                                        c->d_cg.p, alpha_ptr(c), beta_ptr(c), nullptr, nullptr, 0.0);
    else
        k_update<false, false><<<gu, 256, 0, c->stream>>>(c->n_dofs, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter,
                                                          c->d_cg.p, alpha_ptr(c), beta_ptr(c), nullptr, nullptr, 0.0);
    k_direction<false><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_d.p, c->d_cg.p, beta_ptr(c));
    c->launches += 2;
}

void launch_cg_iteration(fb_ctx* c, int lanes) {
    if (c->world > 1) {         // peer-mapped mode (the NCCL mode issues its iteration from api.cu): 6 kernels, no host involvement
        launch_pack_p2p(c, c->d_d.p);
        launch_cg_spmv(c, lanes);
        launch_cg_scalars(c, 1);
        if (c->tl_active) { launch_tl_vectors(c); return; }     // update | restrict | exchange | coarse solve | direction (twolevel.cu)
        launch_cg_update_only(c);
        launch_cg_scalars(c, 2);
        launch_cg_direction_only(c);
        return;
    }
    launch_cg_spmv(c, lanes);
    launch_cg_vectors(c);
}

// Persistent solve: returns false when the system does not fit the per-SM shared memory / register budget.
template <int PT>
static cudaError_t launch_persistent_pt(fb_ctx* c, size_t smem) {
    auto kern = k_cg_persistent<512, PT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    const int* a0 = c->d_cta_row.p; const int* a1 = c->d_rowptr.p; const int* a2 = c->d_col.p; const double* a3 = c->d_val_save.p;
    const double* a4 = c->d_rhs.p; const double* a5 = c->d_dinv.p; double* a6 = c->d_x.p; double* a7 = c->d_d.p;
    double* a8 = c->d_partial.p; int* a8b = c->d_pers_flags.p; CgScalars* a9 = c->d_cg.p; int a10 = c->pers_cap, a11 = c->pers_rmax;
    long long* a12 = c->cg_debug ? c->d_dbg.p : nullptr;
    void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a8b, &a9, &a10, &a11, &a12};
    return cudaLaunchCooperativeKernel((void*) kern, dim3(c->pers_grid), dim3(512), args, smem, c->stream);
}

bool persistent_eligible(fb_ctx* c) {
    if (c->cg_persistent == 0 || c->n_dofs <= 0) return false;
    if (c->pers_grid == 0) {
        // nnz-balanced contiguous row slices, one per SM
        const int G = std::min(std::min(c->pers_ctas > 0 ? std::min(c->pers_ctas, c->n_sm) : c->n_sm, 160), std::max(1, c->n_dofs / 8));   // grid_allreduce reads <= 160 partials
        std::vector<int> cta_row(G + 1, c->n_dofs);
        cta_row[0] = 0;
        int r = 0, cap = 0, rmax = 0;
        for (int b = 0; b < G; ++b) {
            const long target = (long) c->nnz * (b + 1) / G;
            const int start = r;
            while (r < c->n_dofs && (c->rowptr[r + 1] <= target || r == start) && (c->n_dofs - r) > (G - 1 - b)) ++r;
            if (b == G - 1) r = c->n_dofs;
            cta_row[b + 1] = r;
            cap = std::max(cap, c->rowptr[r] - c->rowptr[start]);
            rmax = std::max(rmax, r - start);
        }
        c->pers_grid = G; c->pers_cap = (cap + 1) & ~1; c->pers_rmax = (rmax + 1) & ~1;
        c->pers_cta_row = cta_row;
        c->pers_uploaded = false;
    }
    const size_t smem = 16 * (size_t) c->pers_cap + 40 * (size_t) c->pers_rmax + 4 * ((size_t) c->pers_rmax + 2);
    return c->pers_cap <= 24 * 512 && smem <= 220 * 1024;
}

cudaError_t launch_cg_persistent(fb_ctx* c) {
    if (!c->pers_uploaded) {
        cudaError_t e = c->d_cta_row.upload(c->pers_cta_row, c->stream);
        if (e == cudaSuccess) e = c->d_pers_flags.alloc(c->n_sm);
        if (e == cudaSuccess) e = c->d_dbg.alloc(64);
        if (e != cudaSuccess) return e;
        c->pers_uploaded = true;
    }
    cudaError_t ez = cudaMemsetAsync(c->d_pers_flags.p, 0, c->n_sm * sizeof(int), c->stream);    // sequence flags restart at 0
    if (ez == cudaSuccess) ez = cudaMemsetAsync(c->d_dbg.p, 0, 64 * sizeof(long long), c->stream);
    if (ez != cudaSuccess) return ez;
    const size_t smem = 16 * (size_t) c->pers_cap + 40 * (size_t) c->pers_rmax + 4 * ((size_t) c->pers_rmax + 2);
    const int pt = (c->pers_cap + 511) / 512;
    c->launches++;
    if (pt <= 4) return launch_persistent_pt<4>(c, smem);
    if (pt <= 8) return launch_persistent_pt<8>(c, smem);
    if (pt <= 12) return launch_persistent_pt<12>(c, smem);
    if (pt <= 16) return launch_persistent_pt<16>(c, smem);
    return launch_persistent_pt<24>(c, smem);
}

// ---- multi-GPU helpers -------------------------------------------------------------------
__global__ void k_pack(int n, const int* __restrict__ idx, const double* __restrict__ v, double* __restrict__ out) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) out[i] = v[idx[i]];
}
// Halo exchange of the search direction WITHOUT NCCL: every owned value a peer needs is stored directly into that peer's
// ghost segment over NVLink; the last block to finish publishes "exchange #seq has landed" in the peers' slot records
// (release at system scope, after every block has fenced its stores) and then waits for the same announcement from the
// ranks it receives from, so that the kernel boundary orders the SpMV behind the complete ghost segment.  The peers'
// previous SpMV has finished reading the old ghost values: this kernel runs behind an all-reduce that needed their sums.
__global__ void __launch_bounds__(256) k_pack_p2p(int n, const int* __restrict__ idx, const double* __restrict__ v, P2pDesc* D,
                                                  unsigned* counter, CgScalars* cgs) {
    if (cgs->done) return;
    __shared__ bool is_last;
    const int W = D->world;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        int p = 0;
        while (p + 1 < W && i >= D->send_off[p + 1]) ++p;
        D->peer_vec[p][D->dst_base[p] + (int) (i - D->send_off[p])] = v[idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last || threadIdx.x >= 32) return;
    // the first warp of the last block: lane p announces to / waits for rank p (all peers at once, not one after the other)
    const int lane = threadIdx.x, me = D->rank;
    const long long seq = D->halo_count + 1;
    __syncwarp();
    if (lane == 0) { *counter = 0; D->halo_count = seq; }
    __threadfence_system();
    if (lane < W && lane != me) {
        if (D->send_off[lane + 1] > D->send_off[lane]) st_release_sys(&D->slots[lane]->halo_flag[me], seq);
        if (D->n_recv[lane] > 0) {
            const long long t0 = clock64();
            while (ld_acquire_sys(&D->slots[me]->halo_flag[lane]) < seq)
                if (clock64() - t0 > P2P_SPIN_LIMIT) { cgs->done = 3; break; }
        }
    }
}

__global__ void k_flags_to_double(int n, const int* __restrict__ flag, double* __restrict__ out) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) out[i] = flag[i] ? 1.0 : 0.0;
}
__global__ void k_double_to_flags(int n0, int n1, const double* __restrict__ in, int* __restrict__ flag) {
    for (long i = n0 + (long) blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (long) gridDim.x * blockDim.x) flag[i] = in[i] != 0.0;
}
void launch_pack(fb_ctx* c, const double* v) {
    const int n = (int) c->send_idx.size();
    if (n == 0) return;
    k_pack<<<grid_for(c, n, 256), 256, 0, c->stream>>>(n, c->d_send_idx.p, v, c->d_sendbuf.p);
    c->launches++;
}
void launch_pack_p2p(fb_ctx* c, const double* v) {
    const int n = (int) c->send_idx.size();
    // (a rank without halo still takes part: the kernel is what advances halo_count on every rank alike)
    const int g = std::max(1, std::min(grid_for(c, std::max(1, n), 256), c->n_sm * 2));
    k_pack_p2p<<<g, 256, 0, c->stream>>>(n, c->d_send_idx.p, v, c->d_p2p.p, c->d_p2p_counter.p, c->d_cg.p);
    c->launches++;
}
void launch_flags_to_double(fb_ctx* c, double* out) {
    k_flags_to_double<<<grid_for(c, c->n_cols, 256), 256, 0, c->stream>>>(c->n_cols, c->d_bcflag.p, out);
    c->launches++;
}
void launch_double_to_ghost_flags(fb_ctx* c, const double* in) {
    if (c->n_cols == c->n_dofs) return;
    k_double_to_flags<<<grid_for(c, c->n_cols - c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->n_cols, in, c->d_bcflag.p);
    c->launches++;
}
void launch_cg_scalars(fb_ctx* c, int which) {      // 0: after the initial residual, 1: after SpMV, 2: after the update
    if (which == 0) k_cg_scalars_spmv<true><<<1, 32, 0, c->stream>>>(c->d_cg.p, nullptr);
    else if (which == 1) k_cg_scalars_spmv<false><<<1, 32, 0, c->stream>>>(c->d_cg.p, alpha_ptr(c));
    else k_cg_scalars_update<<<1, 32, 0, c->stream>>>(c->d_cg.p, beta_ptr(c));
    c->launches++;
}
void launch_cg_init_spmv(fb_ctx* c, int lanes) { spmv_dispatch<true>(c, lanes, c->d_x.p, c->d_g.p, nullptr); }
void launch_cg_init_direction(fb_ctx* c) {
    const int g = grid_for(c, c->n_dofs, 256);
    k_direction<true><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_d.p, c->d_cg.p, nullptr);
    c->launches++;
}
void launch_cg_update_only(fb_ctx* c) {
    unsigned* counter = (unsigned*) (c->d_partial.p + c->d_partial.n - 8);
    const int g = std::min(grid_for(c, c->n_dofs, 256), c->n_sm * 6);
    k_update<false, false><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter,
                                                     c->d_cg.p, alpha_ptr(c), beta_ptr(c), nullptr, nullptr, 0.0);
    c->launches++;
}
void launch_cg_direction_only(fb_ctx* c) {
    const int g = grid_for(c, c->n_dofs, 256);
    k_direction<false><<<g, 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_d.p, c->d_cg.p, beta_ptr(c));
    c->launches++;
}

void launch_solution_grad(fb_ctx* c, double* d_grad3) {
    k_solution_grad<<<(c->n_vert + 127) / 128, 128, 0, c->stream>>>(c->n_vert, c->d_vert_lastcell.p, c->d_cells.p, c->d_vxyz.p, c->d_x.p, d_grad3);
    c->launches++;
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
