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
#include "nccl_dyn.h"

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
                                                          const double* __restrict__ val, const int* __restrict__ bcflag,
                                                          const int* __restrict__ agg, double* __restrict__ Ac) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        if (bcflag[i]) continue;                          // constrained row: not part of any aggregate
        const int I = agg[i];
        double own = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = col[k];                         // (partitioned: ghost columns carry the owner's flag and aggregate)
            if (bcflag[j]) continue;
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
        if (cgs->red) { cgs->red[0] = tot[0]; cgs->red[1] = tot[1]; return; }     // partitioned: k_tl_allreduce_p2p finishes the step
        const int it = cgs->it + 1;
        cgs->it = it; cgs->res2 = tot[1];
        if (tot[1] <= cgs->tol2) cgs->done = 1;
        else if (it >= cgs->max_iter || tot[1] != tot[1]) cgs->done = 2;
        slot[0] = tot[0];
    }
}

// r_c[I] = sum of g over aggregate I (its dofs are perm[I agg_size ...]); one block per aggregate, fixed summation order
__global__ void __launch_bounds__(256) k_tl_restrict(int n, int agg_size, const int* __restrict__ aoff, const int* __restrict__ perm,
                                                     const double* __restrict__ g, double* __restrict__ rc, const CgScalars* __restrict__ cgs) {
    if (cgs->done) return;
    __shared__ double sm[8];
    // one GPU: aggregate I = perm[I agg_size, (I + 1) agg_size); partitioned: this rank's rows of aggregate I = perm[aoff[I], aoff[I + 1])
    const long a = aoff ? aoff[blockIdx.x] : (long) blockIdx.x * agg_size, e = aoff ? aoff[blockIdx.x + 1] : min((long) n, a + agg_size);
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


// ---- partitioned two-level (peer-mapped mode) --------------------------------------------------------------------
__device__ __forceinline__ void tl_st_release_sys(long long* p, long long v) { asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ long long tl_ld_acquire_sys(const long long* p) {
    long long v; asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
// All-reduce over the ranks of this rank's P^T g (n_c values) and, outside the initial step, of its g.Dinv g and g.g:
// every block stores its share of the payload into EVERY rank's slot record over NVLink and fences; the last block
// publishes the sequence number, waits for the other ranks', and adds the contributions in rank order -- the same
// totals, bit for bit, on every rank.  It then does the bookkeeping k_tl_update does on one GPU.
template <bool INIT>
__global__ void __launch_bounds__(256) k_tl_allreduce_p2p(P2pDesc* D, int nc, const double* __restrict__ rc_part, double* __restrict__ rc,
                                                          CgScalars* cgs, double* __restrict__ slot, unsigned* counter) {
    if (cgs->done) return;
    __shared__ bool is_last;
    __shared__ double s_tot[2];
    const int W = D->world, me = D->rank, len = nc + (INIT ? 0 : 2);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += gridDim.x * blockDim.x) {
        const double v = k < nc ? rc_part[k] : cgs->red[k - nc];
        for (int p = 0; p < W; ++p) D->slots[p]->tl[me][k] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x < 32) {                 // lane p announces to / waits for rank p, all peers at once
        const int lane = threadIdx.x;
        const long long seq = D->tl_count + 1;
        __syncwarp();
        if (lane == 0) { *counter = 0; D->tl_count = seq; }
        __threadfence_system();
        if (lane < W) {
            tl_st_release_sys(&D->slots[lane]->tl_seq[me], seq);
            const long long t0 = clock64();
            while (tl_ld_acquire_sys(&D->slots[me]->tl_seq[lane]) < seq)
                if (clock64() - t0 > 8000000000LL) { cgs->done = 3; break; }
        }
    }
    __syncthreads();
    if (cgs->done == 3) return;
    const P2pSlots* mine = D->slots[me];
    for (int k = threadIdx.x; k < len; k += blockDim.x) {
        double s = 0;
        for (int p = 0; p < W; ++p) s += ((const volatile double*) mine->tl[p])[k];
        if (k < nc) rc[k] = s; else s_tot[k - nc] = s;
    }
    __syncthreads();
    if (!INIT && threadIdx.x == 0) {
        const int it = cgs->it + 1;
        cgs->it = it; cgs->res2 = s_tot[1];
        if (s_tot[1] <= cgs->tol2) cgs->done = 1;
        else if (it >= cgs->max_iter || s_tot[1] != s_tot[1]) cgs->done = 2;
        slot[0] = s_tot[0];
    }
}
__global__ void k_tl_gather_agg(int n, const int* __restrict__ l2g, const int* __restrict__ gagg, int* __restrict__ agg, int* __restrict__ idx) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) { agg[i] = gagg[l2g[i]]; idx[i] = (int) i; }
}
__global__ void k_tl_offsets(int nc, int n, const int* __restrict__ sorted_agg, int* __restrict__ aoff) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I > nc) return;
    int lo = 0, hi = n;                                   // first position whose aggregate is >= I
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_agg[mid] < I) lo = mid + 1; else hi = mid; }
    aoff[I] = lo;
}
__global__ void k_tl_zero_unless(int keep, size_t n, double* __restrict__ a) {
    if (keep) return;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) a[i] = 0.0;
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
    const bool part = c->world > 1;
    if (part && !c->p2p_ready) return c->fail(FB_ERR_ARG, "FB_PRECOND_TWOLEVEL on a partitioned mesh needs the peer-mapped iteration (fb_comm_mode 2)");
    auto nccl_sum = [&](double* buf, size_t count) -> int {
        ncclResult_t r = Nccl::get().AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t) c->nccl_comm, s);
        return r == ncclSuccess ? FB_OK : c->fail(FB_ERR_CUDA, "ncclAllReduce failed in the two-level set-up: %s", Nccl::get().GetErrorString(r));
    };
    if (!c->tl_agg_ready) {
        // aggregates = runs of `agg` dofs along the Morton curve through ALL dofs of the mesh (partitioned: every rank sorts the
        // global vertex list itself -- 5 ms -- and keeps the aggregate of its owned rows and ghost columns)
        const long ng = part ? c->n_vert_global : n;
        int agg = c->tl_agg_opt > 0 ? c->tl_agg_opt : std::max(256, (int) ((ng + 4095L) / 4096));
        agg = (agg + 63) & ~63;
        c->tl_agg = agg; c->tl_nc = (int) ((ng + (long) agg - 1) / agg);
        if (part && c->tl_nc + 2 > TL_SLOT) return c->fail(FB_ERR_ARG, "FB_PRECOND_TWOLEVEL: %d aggregates exceed the exchange slot (%d); raise tl_agg", c->tl_nc, TL_SLOT - 2);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
#pragma omp parallel for schedule(static) reduction(min : lo[:3]) reduction(max : hi[:3])
        for (long v = 0; v < ng; ++v) {
            if (!part && c->dof2vertex[v] < 0) continue;          // FE_Q(2): the other support points lie inside the hull of the vertices
            const double* p = part ? &c->part_gxyz[3 * (size_t) v] : &c->xyz[3 * (size_t) c->vert2node[c->dof2vertex[v]]];
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
        }
        double sc[3];
        for (int k = 0; k < 3; ++k) sc[k] = 2097151.0 / std::max(hi[k] - lo[k], 1e-300);
        DevBuf<unsigned long long> key_in, key_out; DevBuf<int> idx_in, perm_g, gagg; DevBuf<unsigned char> tmp; DevBuf<double> gxyz;
        FB_CUDA(c, key_in.alloc(ng)); FB_CUDA(c, key_out.alloc(ng)); FB_CUDA(c, idx_in.alloc(ng));
        FB_CUDA(c, c->d_tl_perm.alloc(n)); FB_CUDA(c, c->d_tl_agg.alloc(c->n_cols));
        if (part) { FB_CUDA(c, gxyz.upload(c->part_gxyz, s)); FB_CUDA(c, perm_g.alloc(ng)); FB_CUDA(c, gagg.alloc(ng)); }
        int* perm_out = part ? perm_g.p : c->d_tl_perm.p;
        k_tl_keys<<<grid_for(c, ng, 256), 256, 0, s>>>((int) ng, part ? gxyz.p : c->d_vxyz.p, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], key_in.p, idx_in.p);
        size_t bytes = 0;
        FB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.p, key_out.p, idx_in.p, perm_out, (int) ng, 0, 63, s));
        FB_CUDA(c, tmp.alloc(bytes));
        FB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_in.p, key_out.p, idx_in.p, perm_out, (int) ng, 0, 63, s));
        k_tl_agg<<<grid_for(c, ng, 256), 256, 0, s>>>((int) ng, agg, perm_out, part ? gagg.p : c->d_tl_agg.p);
        c->launches += 2;
        if (part) {
            // aggregate of every local column; this rank's rows grouped by aggregate (perm + offsets) for the restriction
            DevBuf<int> a_sorted, idx2;
            FB_CUDA(c, a_sorted.alloc(n)); FB_CUDA(c, idx2.alloc(c->n_cols)); FB_CUDA(c, c->d_tl_aoff.alloc(c->tl_nc + 1));
            k_tl_gather_agg<<<grid_for(c, c->n_cols, 256), 256, 0, s>>>(c->n_cols, c->d_l2g.p, gagg.p, c->d_tl_agg.p, idx2.p);
            size_t b2 = 0;
            FB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, b2, c->d_tl_agg.p, a_sorted.p, idx2.p, c->d_tl_perm.p, n, 0, 32, s));
            if (b2 > tmp.n) FB_CUDA(c, tmp.alloc(b2));
            FB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, b2, c->d_tl_agg.p, a_sorted.p, idx2.p, c->d_tl_perm.p, n, 0, 32, s));
            k_tl_offsets<<<(c->tl_nc + 256) / 256, 256, 0, s>>>(c->tl_nc, n, a_sorted.p, c->d_tl_aoff.p);
            c->launches += 2;
            FB_CUDA(c, cudaStreamSynchronize(s));
        }
        FB_CUDA(c, cudaStreamSynchronize(s));              // the temporaries go out of scope
        c->tl_agg_ready = true;
    }
    const int nc = c->tl_nc;
    FB_CUDA(c, c->d_tl_inv.alloc((size_t) nc * nc)); FB_CUDA(c, c->d_tl_rc.alloc(nc)); FB_CUDA(c, c->d_tl_ec.alloc(nc));
    if (part) FB_CUDA(c, c->d_tl_rc_part.alloc(nc));
    FB_CUDA(c, cudaMemsetAsync(c->d_tl_inv.p, 0, sizeof(double) * (size_t) nc * nc, s));
    k_tl_coarse_matrix<<<grid_for(c, n, 256), 256, 0, s>>>(n, nc, c->d_rowptr.p, c->d_col.p, c->d_val_save.p, c->d_bcflag.p, c->d_tl_agg.p, c->d_tl_inv.p);
    c->launches++;
    if (part) { const int rc = nccl_sum(c->d_tl_inv.p, (size_t) nc * nc); if (rc) return rc; }       // every rank adds the rows it owns
    k_tl_fix_diag<<<(nc + 255) / 256, 256, 0, s>>>(nc, c->d_tl_inv.p);
    c->launches++;
    if (part && c->rank != 0) {
        // the inverse has to be bit-identical on every rank (identical beta / convergence decisions): rank 0 inverts, the
        // others contribute zeros to a sum
        k_tl_zero_unless<<<grid_for(c, (long) nc * nc, 256), 256, 0, s>>>(0, (size_t) nc * nc, c->d_tl_inv.p);
        const int rc = nccl_sum(c->d_tl_inv.p, (size_t) nc * nc); if (rc) return rc;
        double first = 0;
        FB_CUDA(c, cudaMemcpyAsync(&first, c->d_tl_inv.p, sizeof(double), cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
        if (first != first) return c->fail(FB_ERR_CUDA, "FB_PRECOND_TWOLEVEL: the Cholesky inverse of the coarse matrix failed on rank 0");
        c->tl_ready = true;
        return FB_OK;
    }
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
    if (!ok && part) {      // the other ranks are waiting in the broadcast below: hand them NaNs, every rank then fails alike
        cudaMemsetAsync(c->d_tl_inv.p, 0xff, sizeof(double) * (size_t) nc * nc, s);
        nccl_sum(c->d_tl_inv.p, (size_t) nc * nc);
        cudaStreamSynchronize(s);
    }
    if (!ok) return c->fail(FB_ERR_CUDA, "FB_PRECOND_TWOLEVEL: Cholesky inverse of the %d x %d coarse matrix failed (info %d)", nc, nc, info);
    k_tl_symmetrize<<<(unsigned) (((long) nc * nc + 255) / 256), 256, 0, s>>>(nc, c->d_tl_inv.p);
    c->launches++;
    if (part) { const int rc = nccl_sum(c->d_tl_inv.p, (size_t) nc * nc); if (rc) return rc; }       // rank 0's inverse reaches everybody
    FB_CUDA(c, cudaStreamSynchronize(s));
    c->tl_ready = true;
    return FB_OK;
}

void tl_release(fb_ctx* c) {
    if (c->tl_solver && Solver::get().ok) Solver::get().Destroy((cusolverDnHandle_t) c->tl_solver);
    c->tl_solver = nullptr;
}

// tail of the initial step (after the INIT SpMV has left g, gh = g.Dinv g, |g|): coarse part of z, first direction
static void launch_tl_exchange(fb_ctx* c, bool init) {      // partitioned: P^T g (+ the two scalars) over the peer mappings
    const int nc = c->tl_nc;
    const int g = std::max(1, std::min(16, (nc + 2 + 255) / 256));
    if (init) k_tl_allreduce_p2p<true><<<g, 256, 0, c->stream>>>(c->d_p2p.p, nc, c->d_tl_rc_part.p, c->d_tl_rc.p, c->d_cg.p, slots(c) + 4, c->d_p2p_counter.p);
    else k_tl_allreduce_p2p<false><<<g, 256, 0, c->stream>>>(c->d_p2p.p, nc, c->d_tl_rc_part.p, c->d_tl_rc.p, c->d_cg.p, slots(c) + 4, c->d_p2p_counter.p);
    c->launches++;
}

void launch_tl_init_tail(fb_ctx* c) {
    const int nc = c->tl_nc;
    const bool part = c->world > 1;
    k_tl_restrict<<<nc, 256, 0, c->stream>>>(c->n_dofs, c->tl_agg, part ? c->d_tl_aoff.p : nullptr, c->d_tl_perm.p, c->d_g.p,
                                             part ? c->d_tl_rc_part.p : c->d_tl_rc.p, c->d_cg.p);
    if (part) launch_tl_exchange(c, true);
    k_tl_apply<true><<<std::min(c->n_sm * 4, (nc + 7) / 8), 256, 0, c->stream>>>(nc, c->d_tl_inv.p, c->d_tl_rc.p, c->d_tl_ec.p, c->d_partial.p, counter_of(c),
                                                                              c->d_cg.p, slots(c) + 4, nullptr);
    k_tl_direction<true><<<grid_for(c, c->n_dofs, 256), 256, 0, c->stream>>>(c->n_dofs, c->d_g.p, c->d_dinv.p, c->d_tl_agg.p, c->d_tl_ec.p, c->d_d.p,
                                                                           c->d_cg.p, nullptr);
    c->launches += 3;
}

// the vector part of one iteration (after the SpMV)
void launch_tl_vectors(fb_ctx* c) {
    const int nc = c->tl_nc, n = c->n_dofs;
    const bool part = c->world > 1;
    const int gu = std::min(grid_for(c, n, 256), c->n_sm * 6);
    k_tl_update<<<gu, 256, 0, c->stream>>>(n, c->d_d.p, c->d_h.p, c->d_dinv.p, c->d_x.p, c->d_g.p, c->d_partial.p, counter_of(c), c->d_cg.p,
                                           slots(c), slots(c) + 4);
    k_tl_restrict<<<nc, 256, 0, c->stream>>>(n, c->tl_agg, part ? c->d_tl_aoff.p : nullptr, c->d_tl_perm.p, c->d_g.p,
                                             part ? c->d_tl_rc_part.p : c->d_tl_rc.p, c->d_cg.p);
    if (part) launch_tl_exchange(c, false);
    k_tl_apply<false><<<std::min(c->n_sm * 4, (nc + 7) / 8), 256, 0, c->stream>>>(nc, c->d_tl_inv.p, c->d_tl_rc.p, c->d_tl_ec.p, c->d_partial.p, counter_of(c),
                                                                               c->d_cg.p, slots(c) + 4, slots(c) + 1);
    k_tl_direction<false><<<grid_for(c, n, 256), 256, 0, c->stream>>>(n, c->d_g.p, c->d_dinv.p, c->d_tl_agg.p, c->d_tl_ec.p, c->d_d.p, c->d_cg.p, slots(c) + 1);
    c->launches += 4;
}

}  // namespace fb
