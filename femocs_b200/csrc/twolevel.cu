// Two-level preconditioner of the Jacobi-PCG for HBM-sized systems (FB_PRECOND_TWOLEVEL):
//
//     M^-1 = D^-1 + P (P^T K P)^-1 P^T ,      P = piecewise constants over aggregates of free dofs.
//
// Jacobi-PCG needs O(L / h) iterations because nothing in D^-1 sees the smooth error modes (2 363 iterations on X); the
// additive coarse correction removes exactly those.  Aggregates are runs of `tl_agg` dofs along the Morton curve through
// the dof coordinates (compact boxes whatever the dof numbering is); the Galerkin matrix P^T K P (a few thousand rows) is
// accumulated on the device, inverted ONCE per matrix (dense Cholesky; cuSOLVER potrf / potri bound at run time -- set-up
// only, nothing of it runs inside the iteration) and kept as a dense symmetric matrix, so that the coarse solve inside
// the iteration is one dense matrix-vector product from HBM (8 n_c^2 bytes) by a hand-written kernel.
// M^-1 is a fixed symmetric positive definite operator: plain CG stays valid, the solution of K phi = b is unchanged,
// only the iteration count drops (scripts/two_level_prototype.py: 1095 -> 352 on X refined once, n_c = 5808).
//
// The reference preconditions with SSOR (DealSolver.cpp:447-449), which is sequential; north_star asks for Jacobi /
// Chebyshev and parity at equal residual tolerance -- this is a third preconditioner under the same contract.
//
// Iteration (one GPU): k_spmv_* (h = K d, alpha) | k_tl_update (x, g, |g|, g.Dinv g) | k_tl_restrict (r_c = P^T g) |
// k_tl_apply (e_c = A_c^-1 r_c, g.z = g.Dinv g + r_c.e_c, beta) | k_tl_direction (d = beta d - Dinv g - P e_c).
// Extra traffic per iteration: 8 n_c^2 + 32 n bytes on top of 12 nnz + 108 n.
#include <cub/device/device_radix_sort.cuh>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <vector>

#include "kernels.h"

namespace fb {

namespace {

// ---- cuSOLVER, bound at run time (the library must load without it) ----
struct Solver {
    cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
    cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*PotrfBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*Potrf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*PotriBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*Potri)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
    bool ok = false;
    static Solver& get() {
        static Solver s; static bool tried = false;
        if (tried) return s;
        tried = true;
        void* h = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libcusolver.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return s;
#define FB_SYM(field, name) *(void**) (&s.field) = dlsym(h, name); if (!s.field) return s;
        FB_SYM(Create, "cusolverDnCreate") FB_SYM(Destroy, "cusolverDnDestroy") FB_SYM(SetStream, "cusolverDnSetStream")
        FB_SYM(PotrfBuf, "cusolverDnDpotrf_bufferSize") FB_SYM(Potrf, "cusolverDnDpotrf")
        FB_SYM(PotriBuf, "cusolverDnDpotri_bufferSize") FB_SYM(Potri, "cusolverDnDpotri")
#undef FB_SYM
        s.ok = true;
        return s;
    }
};

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

__global__ void k_tl_keys(int n, const double* __restrict__ vxyz, double lx, double ly, double lz, double sx, double sy, double sz,
                          unsigned long long* __restrict__ key, int* __restrict__ idx) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const double qx = fmin(fmax((vxyz[3 * i] - lx) * sx, 0.0), 2097151.0), qy = fmin(fmax((vxyz[3 * i + 1] - ly) * sy, 0.0), 2097151.0),
                     qz = fmin(fmax((vxyz[3 * i + 2] - lz) * sz, 0.0), 2097151.0);
        key[i] = spread21((unsigned long long) qx) | (spread21((unsigned long long) qy) << 1) | (spread21((unsigned long long) qz) << 2);
        idx[i] = (int) i;
    }
}

__global__ void k_tl_agg(int n, int agg_size, const int* __restrict__ perm, int* __restrict__ agg) {
    for (long k = (long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long) gridDim.x * blockDim.x) agg[perm[k]] = (int) (k / agg_size);
}

// Galerkin matrix of the piecewise-constant prolongator restricted to the free dofs: A_c[I][J] = sum_{i in I, j in J} K_ij.
// One thread per fine row; the entries that stay inside the row's own aggregate (most of them) are summed in a register.
__global__ void __launch_bounds__(256) k_tl_coarse_matrix(int n, int nc, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                          const double* __restrict__ val, const double* __restrict__ dinv,
                                                          const int* __restrict__ agg, double* __restrict__ Ac) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        if (dinv[i] == 0.0) continue;                     // constrained row: not part of any aggregate
        const int I = agg[i];
        double own = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = col[k];
            if (dinv[j] == 0.0) continue;
            const int J = agg[j];
            if (J == I) own += val[k]; else atomicAdd(&Ac[(size_t) I * nc + J], val[k]);
        }
        atomicAdd(&Ac[(size_t) I * nc + I], own);
    }
}
__global__ void k_tl_fix_diag(int nc, double* __restrict__ Ac) {      // an aggregate of constrained dofs only: identity row
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I < nc && Ac[(size_t) I * nc + I] == 0.0) Ac[(size_t) I * nc + I] = 1.0;
}
// cuSOLVER leaves the inverse in one triangle (column-major LOWER = the upper triangle of the row-major view)
__global__ void k_tl_symmetrize(int nc, double* __restrict__ M) {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long) nc * nc) return;
    const int i = (int) (t / nc), j = (int) (t % nc);
    if (j > i) M[(size_t) j * nc + i] = M[(size_t) i * nc + j];
}

// block sum of two accumulators + "last block reduces the partials" (same scheme as poisson_kernels.cu)
__device__ __forceinline__ bool reduce2(double (&v)[2], double* __restrict__ partial, unsigned* counter, double (&total)[2]) {
    __shared__ double sm[2][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            double s = 0;
            for (int w = 0; w < nwarp; ++w) s += sm[k][w];
            partial[(size_t) k * gridDim.x + blockIdx.x] = s;
        }
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
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
        for (int k = 0; k < 2; ++k) { double s = 0; for (int w = 0; w < nwarp; ++w) s += sm[k][w]; total[k] = s; }
        *counter = 0;
    }
    return threadIdx.x == 0;
}

// x += alpha d ; g += alpha h (0 on constrained rows) ; |g|^2 -> iteration count / convergence ; g.Dinv g -> slot[0]
__global__ void __launch_bounds__(256, 6) k_tl_update(int n, const double* __restrict__ d, const double* __restrict__ h, const double* __restrict__ dinv,
                                                      double* __restrict__ x, double* __restrict__ g, double* __restrict__ partial, unsigned* counter,
                                                      CgScalars* __restrict__ cgs, const double* __restrict__ alpha_in, double* __restrict__ slot) {
    if (cgs->done) return;
    const double alpha = *alpha_in;
    double acc[2] = {0, 0};
    const long n2 = n >> 1, stride = (long) gridDim.x * blockDim.x;
    const double2* __restrict__ d2 = reinterpret_cast<const double2*>(d);
    const double2* __restrict__ i2 = reinterpret_cast<const double2*>(dinv);
    const double2* __restrict__ h2 = reinterpret_cast<const double2*>(h);
    double2* __restrict__ x2 = reinterpret_cast<double2*>(x);
    double2* __restrict__ g2 = reinterpret_cast<double2*>(g);
#pragma unroll 1
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 dd = __ldg(&d2[i]), di = __ldg(&i2[i]), hh = __ldg(&h2[i]);
        double2 xx = x2[i], gg = g2[i];
        xx.x += alpha * dd.x; xx.y += alpha * dd.y;
        gg.x = di.x != 0.0 ? gg.x + alpha * hh.x : 0.0; gg.y = di.y != 0.0 ? gg.y + alpha * hh.y : 0.0;
        x2[i] = xx; g2[i] = gg;
        acc[0] += gg.x * gg.x * di.x + gg.y * gg.y * di.y;
        acc[1] += gg.x * gg.x + gg.y * gg.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int i = n - 1;
        x[i] += alpha * d[i];
        const double gi = dinv[i] != 0.0 ? g[i] + alpha * h[i] : 0.0;
        g[i] = gi;
        acc[0] += gi * gi * dinv[i]; acc[1] += gi * gi;
    }
    double tot[2];
    if (reduce2(acc, partial, counter, tot)) {
        const int it = cgs->it + 1;
        cgs->it = it; cgs->res2 = tot[1];
        if (tot[1] <= cgs->tol2) cgs->done = 1;
        else if (it >= cgs->max_iter || tot[1] != tot[1]) cgs->done = 2;
        slot[0] = tot[0];
    }
}

// r_c[I] = sum of g over aggregate I (its dofs are perm[I agg_size ...]); one block per aggregate, fixed summation order
__global__ void __launch_bounds__(256) k_tl_restrict(int n, int agg_size, const int* __restrict__ perm, const double* __restrict__ g,
                                                     double* __restrict__ rc, const CgScalars* __restrict__ cgs) {
    if (cgs->done) return;
    __shared__ double sm[8];
    const long a = (long) blockIdx.x * agg_size, e = min((long) n, a + agg_size);
    double s = 0;
    for (long k = a + threadIdx.x; k < e; k += blockDim.x) s += g[perm[k]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; rc[blockIdx.x] = t; }
}

// e_c = A_c^-1 r_c (dense, one warp per row, coalesced 8 n_c bytes per row) ; g.z = g.Dinv g + r_c.e_c ; beta = g.z / gh
template <bool INIT>
__global__ void __launch_bounds__(256) k_tl_apply(int nc, const double* __restrict__ Ainv, const double* __restrict__ rc, double* __restrict__ ec,
                                                  double* __restrict__ partial, unsigned* counter, CgScalars* __restrict__ cgs,
                                                  const double* __restrict__ slot, double* __restrict__ beta_out) {
    if (cgs->done) return;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    double acc[2] = {0, 0};
    for (int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < nc; I += warps) {
        const double* __restrict__ row = Ainv + (size_t) I * nc;
        double s = 0;
        for (int j = lane; j < nc; j += 32) s += row[j] * __ldg(&rc[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { ec[I] = s; acc[0] += s * rc[I]; }
    }
    double tot[2];
    if (reduce2(acc, partial, counter, tot)) {
        if (INIT) cgs->gh += tot[0];                       // the initial residual kernel left gh = g.Dinv g
        else { const double gz = slot[0] + tot[0]; *beta_out = gz / cgs->gh; cgs->gh = gz; }
    }
}

// d = beta d - (Dinv g + P e_c) on free dofs, 0 on constrained ones
template <bool INIT>
__global__ void __launch_bounds__(256) k_tl_direction(int n, const double* __restrict__ g, const double* __restrict__ dinv, const int* __restrict__ agg,
                                                      const double* __restrict__ ec, double* __restrict__ d, const CgScalars* __restrict__ cgs,
                                                      const double* __restrict__ beta_in) {
    if (cgs->done) return;
    const double beta = INIT ? 0.0 : *beta_in;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const double di = dinv[i];
        const double z = di != 0.0 ? di * g[i] + __ldg(&ec[agg[i]]) : 0.0;
        d[i] = (INIT ? 0.0 : beta * d[i]) - z;
    }
}

inline int grid_for(const fb_ctx* c, long work_items, int block) {
    long g = (work_items + block - 1) / block;
    const long cap = (long) c->n_sm * (2048 / block);
    if (g > cap) g = cap;
    return (int) (g < 1 ? 1 : g);
}
inline double* slots(fb_ctx* c) { return (double*) (c->d_cg.p + 1); }      // [0] alpha, [1] beta, [4] g.Dinv g of the two-level iteration
inline unsigned* counter_of(fb_ctx* c) { return (unsigned*) (c->d_partial.p + c->d_partial.n - 8); }

}  // namespace

// Set-up for the matrix currently assembled: Morton aggregates (once per mesh), Galerkin matrix and its dense inverse
// (once per matrix).  Returns FB_ERR_ARG with a reason when the mode cannot be used.
int tl_prepare(fb_ctx* c) {
    if (c->tl_ready) return FB_OK;
    Solver& S = Solver::get();
    if (!S.ok) return c->fail(FB_ERR_ARG, "FB_PRECOND_TWOLEVEL: libcusolver (dense Cholesky of the coarse matrix, set-up only) could not be loaded");
    cudaStream_t s = c->stream;
    const int n = c->n_dofs;
    if (!c->tl_agg_ready) {
        int agg = c->tl_agg_opt > 0 ? c->tl_agg_opt : std::max(256, (int) ((n + 4095L) / 4096));
        agg = (agg + 63) & ~63;
        c->tl_agg = agg; c->tl_nc = (int) ((n + (long) agg - 1) / agg);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
#pragma omp parallel for schedule(static) reduction(min : lo[:3]) reduction(max : hi[:3])
        for (int dof = 0; dof < n; ++dof) {
            const double* p = &c->xyz[3 * (size_t) c->vert2node[c->dof2vertex[dof]]];
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
        }
        double sc[3];
        for (int k = 0; k < 3; ++k) sc[k] = 2097151.0 / std::max(hi[k] - lo[k], 1e-300);
        DevBuf<unsigned long long> key_in, key_out; DevBuf<int> idx_in; DevBuf<unsigned char> tmp;
        FB_CUDA(c, key_in.alloc(n)); FB_CUDA(c, key_out.alloc(n)); FB_CUDA(c, idx_in.alloc(n));
        FB_CUDA(c, c->d_tl_perm.alloc(n)); FB_CUDA(c, c->d_tl_agg.alloc(n));
        k_tl_keys<<<grid_for(c, n, 256), 256, 0, s>>>(n, c->d_vxyz.p, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], key_in.p, idx_in.p);
        size_t bytes = 0;
        FB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.p, key_out.p, idx_in.p, c->d_tl_perm.p, n, 0, 63, s));
        FB_CUDA(c, tmp.alloc(bytes));
        FB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_in.p, key_out.p, idx_in.p, c->d_tl_perm.p, n, 0, 63, s));
        k_tl_agg<<<grid_for(c, n, 256), 256, 0, s>>>(n, agg, c->d_tl_perm.p, c->d_tl_agg.p);
        c->launches += 2;
        FB_CUDA(c, cudaStreamSynchronize(s));              // the temporaries go out of scope
        c->tl_agg_ready = true;
    }
    const int nc = c->tl_nc;
    FB_CUDA(c, c->d_tl_inv.alloc((size_t) nc * nc)); FB_CUDA(c, c->d_tl_rc.alloc(nc)); FB_CUDA(c, c->d_tl_ec.alloc(nc));
    FB_CUDA(c, cudaMemsetAsync(c->d_tl_inv.p, 0, sizeof(double) * (size_t) nc * nc, s));
    k_tl_coarse_matrix<<<grid_for(c, n, 256), 256, 0, s>>>(n, nc, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, c->d_dinv.p, c->d_tl_agg.p, c->d_tl_inv.p);
    k_tl_fix_diag<<<(nc + 255) / 256, 256, 0, s>>>(nc, c->d_tl_inv.p);
    c->launches += 2;
    if (!c->tl_solver) {                                   // created once per context (cusolverDnCreate takes ~0.1 s)
        cusolverDnHandle_t hn = nullptr;
        if (S.Create(&hn) != CUSOLVER_STATUS_SUCCESS) return c->fail(FB_ERR_CUDA, "cusolverDnCreate failed");
        c->tl_solver = hn;
    }
    cusolverDnHandle_t h = (cusolverDnHandle_t) c->tl_solver;
    S.SetStream(h, s);
    int lw1 = 0, lw2 = 0, info = 0;
    DevBuf<double> work; DevBuf<int> d_info;
    bool ok = S.PotrfBuf(h, CUBLAS_FILL_MODE_LOWER, nc, c->d_tl_inv.p, nc, &lw1) == CUSOLVER_STATUS_SUCCESS &&
              S.PotriBuf(h, CUBLAS_FILL_MODE_LOWER, nc, c->d_tl_inv.p, nc, &lw2) == CUSOLVER_STATUS_SUCCESS;
    ok = ok && work.alloc(std::max(1, std::max(lw1, lw2))) == cudaSuccess && d_info.alloc(1) == cudaSuccess;
    if (ok) ok = S.Potrf(h, CUBLAS_FILL_MODE_LOWER, nc, c->d_tl_inv.p, nc, work.p, lw1, d_info.p) == CUSOLVER_STATUS_SUCCESS;
    if (ok) { cudaMemcpyAsync(&info, d_info.p, sizeof(int), cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); ok = (info == 0); }
    if (ok) ok = S.Potri(h, CUBLAS_FILL_MODE_LOWER, nc, c->d_tl_inv.p, nc, work.p, lw2, d_info.p) == CUSOLVER_STATUS_SUCCESS;
    if (ok) { cudaMemcpyAsync(&info, d_info.p, sizeof(int), cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); ok = (info == 0); }
    if (!ok) return c->fail(FB_ERR_CUDA, "FB_PRECOND_TWOLEVEL: Cholesky inverse of the %d x %d coarse matrix failed (info %d)", nc, nc, info);
    k_tl_symmetrize<<<(unsigned) (((long) nc * nc + 255) / 256), 256, 0, s>>>(nc, c->d_tl_inv.p);
    c->launches++;
    FB_CUDA(c, cudaStreamSynchronize(s));
    c->tl_ready = true;
    return FB_OK;
}

void tl_release(fb_ctx* c) {
    if (c->tl_solver && Solver::get().ok) Solver::get().Destroy((cusolverDnHandle_t) c->tl_solver);
    c->tl_solver = nullptr;
}

// tail of the initial step (after the INIT SpMV has left g, gh = g.Dinv g, |g|): coarse part of z, first direction
void launch_tl_init_tail(fb_ctx* c) {
    const int nc = c->tl_nc;
    k_tl_restrict<<<nc, 256, 0, c->stream>>>(c->n_dofs, c->tl_agg, c->d_tl_perm.p, c->d_g.p, c->d_tl_rc.p, c->d_cg.p);
    k_tl_apply<true><<<std::min(c->n_sm * 4, (nc + 7) / 8), 256, 0, c->stream>>>(nc, c->d_tl_inv.p, c->d_tl_rc.p, c->d_tl_ec.p, c->d_partial.p, counter_of(c),
                                                                              c->d_cg.p, slots(c) + 4, nullptr);
    k_tl_direction<true><<<grid_for(c, c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_tl_agg.p, c->d_tl_ec.p, c->d_d.p,
                                                                           c->d_cg.p, nullptr);
    c->launches += 3;
}

// the vector part of one iteration (after the SpMV)
void launch_tl_vectors(fb_ctx* c) {
    const int nc = c->tl_nc, n = c->n_dofs;
    const int gu = std::min(grid_for(c, n, 256), c->n_sm * 6);
    k_tl_update<<<gu, 256, 0, c->stream>>>(n, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter_of(c), c->d_cg.p,
                                           slots(c), slots(c) + 4);
    k_tl_restrict<<<nc, 256, 0, c->stream>>>(n, c->tl_agg, c->d_tl_perm.p, c->d_g.p, c->d_tl_rc.p, c->d_cg.p);
    k_tl_apply<false><<<std::min(c->n_sm * 4, (nc + 7) / 8), 256, 0, c->stream>>>(nc, c->d_tl_inv.p, c->d_tl_rc.p, c->d_tl_ec.p, c->d_partial.p, counter_of(c),
                                                                               c->d_cg.p, slots(c) + 4, slots(c) + 1);
    k_tl_direction<false><<<grid_for(c, n, 256), 256, 0, c->stream>>>(n, c->d_g.p, c->d_dinv.p, c->d_tl_agg.p, c->d_tl_ec.p, c->d_d.p, c->d_cg.p, slots(c) + 1);
    c->launches += 4;
}

}  // namespace fb
