// extern "C" entry points of libfemocs_b200 (see include/femocs_b200.h for the contract and the
// reference interface each one replaces).  Host orchestration only: uploads, kernel launches,
// CUDA-graph capture of the CG loop, pinned staging for host<->device copies.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include <omp.h>

#include "kernels.h"
#include "nccl_dyn.h"

static std::string g_create_error;

#define FB_REQUIRE(ctx, cond, msg) \
    do { if (!(cond)) return (ctx)->fail(FB_ERR_ARG, "%s", msg); } while (0)

static int sync_check(fb_ctx* c, const char* what) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return c->fail(FB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return FB_OK;
}

static void drop_graph(fb_ctx* c) {
    if (c->cg_graph) { cudaGraphExecDestroy(c->cg_graph); c->cg_graph = nullptr; }
    c->cg_graph_n = 0;
}

#define FB_NCCL(ctx, call)                                                                      \
    do {                                                                                        \
        ncclResult_t r__ = (call);                                                              \
        if (r__ != ncclSuccess)                                                                 \
            return (ctx)->fail(FB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, fb::Nccl::get().GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

// owner -> ghost copy of a dof vector (n_cols entries, ghosts behind the n_dofs owned rows): pack the owned
// entries the peers need, then grouped point-to-point transfers over NVLink; receives land directly in the
// contiguous ghost segment of each peer.  All on the context's stream.
static int halo_exchange(fb_ctx* c, double* vec) {
    if (c->world == 1) return FB_OK;
    fb::Nccl& N = fb::Nccl::get();
    ncclComm_t comm = (ncclComm_t) c->nccl_comm;
    fb::launch_pack(c, vec);
    FB_NCCL(c, N.GroupStart());
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        const int ns = c->send_off[p + 1] - c->send_off[p], nr = c->recv_off[p + 1] - c->recv_off[p];
        if (ns > 0) FB_NCCL(c, N.Send(c->d_sendbuf.p + c->send_off[p], ns, ncclDouble, p, comm, c->stream));
        if (nr > 0) FB_NCCL(c, N.Recv(vec + c->n_dofs + c->recv_off[p], nr, ncclDouble, p, comm, c->stream));
    }
    FB_NCCL(c, N.GroupEnd());
    return FB_OK;
}

static int allreduce(fb_ctx* c, double* buf, int n, ncclRedOp_t op) {
    if (c->world == 1) return FB_OK;
    FB_NCCL(c, fb::Nccl::get().AllReduce(buf, buf, n, ncclDouble, op, (ncclComm_t) c->nccl_comm, c->stream));
    return FB_OK;
}

// ---- peer-mapped iteration: CUDA IPC mappings of every peer's search-direction vector and slot
// record, exchanged once per partitioned import with ncclAllGather.  Collective: every rank leaves with the same verdict
// (p2p_ready on all ranks or on none), otherwise the ranks would run different iteration protocols.
static void p2p_release(fb_ctx* c) {
    for (void* m : c->p2p_mapped) cudaIpcCloseMemHandle(m);
    c->p2p_mapped.clear();
    c->p2p_ready = false;
}

static int p2p_setup(fb_ctx* c) {
    using fb::P2P_MAX;
    c->p2p_ready = false;
    if (c->world <= 1 || c->world > P2P_MAX) return FB_OK;
    struct Rec { cudaIpcMemHandle_t vec, slots; int n_dofs, ok; int recv_off[P2P_MAX + 1]; };
    cudaStream_t s = c->stream;
    const int W = c->world, me = c->rank;
    Rec mine; memset(&mine, 0, sizeof mine);
    mine.ok = c->cg_p2p ? 1 : 0;
    if (mine.ok && (c->d_p2p_slots.alloc(1) != cudaSuccess || c->d_p2p.alloc(1) != cudaSuccess || c->d_p2p_counter.alloc(1) != cudaSuccess)) mine.ok = 0;
    if (mine.ok) { c->d_p2p_slots.zero(s); c->d_p2p_counter.zero(s); }
    if (mine.ok && cudaIpcGetMemHandle(&mine.vec, c->d_d.p) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    if (mine.ok && cudaIpcGetMemHandle(&mine.slots, c->d_p2p_slots.p) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    mine.n_dofs = c->n_dofs;
    for (int p = 0; p <= W; ++p) mine.recv_off[p] = c->recv_off[p];
    fb::DevBuf<unsigned char> d_all;
    FB_CUDA(c, d_all.alloc(sizeof(Rec) * (size_t) W));
    FB_CUDA(c, cudaMemcpyAsync(d_all.p + sizeof(Rec) * (size_t) me, &mine, sizeof(Rec), cudaMemcpyHostToDevice, s));
    FB_NCCL(c, fb::Nccl::get().AllGather(d_all.p + sizeof(Rec) * (size_t) me, d_all.p, sizeof(Rec), ncclChar, (ncclComm_t) c->nccl_comm, s));
    std::vector<Rec> all(W);
    FB_CUDA(c, cudaMemcpyAsync(all.data(), d_all.p, sizeof(Rec) * (size_t) W, cudaMemcpyDeviceToHost, s));
    FB_CUDA(c, cudaStreamSynchronize(s));
    bool ok = true;
    for (int p = 0; p < W; ++p) ok = ok && all[p].ok;
    fb::P2pDesc D; memset(&D, 0, sizeof D);
    D.rank = me; D.world = W;
    if (ok) {
        for (int p = 0; p < W && ok; ++p) {
            if (p == me) { D.slots[p] = c->d_p2p_slots.p; D.peer_vec[p] = c->d_d.p; continue; }
            void* pv = nullptr; void* ps = nullptr;
            if (cudaIpcOpenMemHandle(&pv, all[p].vec, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
            c->p2p_mapped.push_back(pv);
            if (cudaIpcOpenMemHandle(&ps, all[p].slots, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
            c->p2p_mapped.push_back(ps);
            D.peer_vec[p] = (double*) pv; D.slots[p] = (fb::P2pSlots*) ps;
            D.dst_base[p] = all[p].n_dofs + all[p].recv_off[me];
        }
    }
    // second round: a rank that could not map a peer vetoes the mode for everybody
    double* hb = (double*) c->pin_out.p;
    hb[0] = ok ? 1.0 : 0.0;
    FB_CUDA(c, cudaMemcpyAsync(c->d_red.p, hb, sizeof(double), cudaMemcpyHostToDevice, s));
    int rc = allreduce(c, c->d_red.p, 1, ncclMin);
    if (rc) return rc;
    FB_CUDA(c, cudaMemcpyAsync(hb, c->d_red.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    FB_CUDA(c, cudaStreamSynchronize(s));
    if (hb[0] != 1.0) { p2p_release(c); return FB_OK; }
    for (int p = 0; p <= W; ++p) D.send_off[p] = c->send_off[p];
    for (int p = 0; p < W; ++p) D.n_recv[p] = c->recv_off[p + 1] - c->recv_off[p];
    FB_CUDA(c, cudaMemcpyAsync(c->d_p2p.p, &D, sizeof D, cudaMemcpyHostToDevice, s));
    FB_CUDA(c, cudaStreamSynchronize(s));
    c->p2p_ready = true;
    return FB_OK;
}

extern "C" {

const char* fb_create_error(void) { return g_create_error.c_str(); }

fb_ctx* fb_create(int device) {
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libfemocs_b200 has no CPU fallback";
        return nullptr;
    }
    if (device < 0 || device >= n_dev) { g_create_error = "invalid CUDA device ordinal"; return nullptr; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return nullptr; }
    fb_ctx* c = new fb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->n_sm = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); delete c; return nullptr;
    }
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    // partials: 2 values x max grid (n_sm * 32 blocks) + counter slot, zero-initialised once
    const size_t np = 2 * (size_t) c->n_sm * 32 + 8;
    if (c->d_partial.alloc(np) != cudaSuccess || c->d_cg.alloc(2) != cudaSuccess || c->d_minmax.alloc(2) != cudaSuccess ||
        c->d_flag.alloc(4) != cudaSuccess || c->pin_out.reserve(4096) != cudaSuccess) {
        g_create_error = "device allocation failed"; fb_destroy(c); return nullptr;
    }
    c->d_partial.zero(c->stream); c->d_cg.zero(c->stream);
    cudaStreamSynchronize(c->stream);
    return c;
}

void fb_destroy(fb_ctx* c) {
    if (!c) return;
    if (c->host_only) { delete c; return; }
    cudaSetDevice(c->device);
    p2p_release(c);
    fb::tl_release(c);
    if (c->nccl_comm) { cudaStreamSynchronize(c->stream); fb::Nccl::get().CommDestroy((ncclComm_t) c->nccl_comm); c->nccl_comm = nullptr; }
    drop_graph(c);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    if (c->stream) { cudaStreamSynchronize(c->stream); }
    cudaStream_t s = c->stream;
    delete c;                       // frees device buffers
    if (s) cudaStreamDestroy(s);
}

const char* fb_last_error(const fb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
long fb_kernel_launches(const fb_ctx* c) { return c->launches; }

int fb_set_option(fb_ctx* c, const char* key, double value) {
    const std::string k(key);
    if (k == "cg_graph_iters") { c->cg_graph_iters = std::max(1, (int) value); drop_graph(c); }
    else if (k == "cheb_degree") { c->cheb_degree = (int) value; drop_graph(c); }
    else if (k == "cheb_eig_ratio") { c->cheb_ratio = value; drop_graph(c); }
    else if (k == "cheb_power_iters") { c->cheb_power_iters = std::max(0, (int) value); c->cheb_lmax = 0; }
    else if (k == "dof_order") c->dof_order = (int) value;
    else if (k == "spmv_kernel") { c->spmv_kernel = (int) value; drop_graph(c); }
    else if (k == "spmv_split") { c->spmv_split = std::max(8, (int) value); c->jds_ready = false; drop_graph(c); }   // segment cap of spmv_kernel 306
    else if (k == "spmv_occ") { c->spmv_occ = std::max(1, std::min(16, (int) value)); drop_graph(c); }
    else if (k == "fe_degree") {            // element of the NEXT fb_import_mesh: 1 = FE_Q(1) (DealSolver.h:130 as shipped), 2 = FE_Q(2)
        if ((int) value != 1 && (int) value != 2) return c->fail(FB_ERR_ARG, "fe_degree must be 1 or 2");
        c->fe_degree = (int) value;
    }
    else if (k == "cg_persistent") c->cg_persistent = (int) value;
    else if (k == "cg_debug") c->cg_debug = (int) value;
    else if (k == "mesh_reuse") c->mesh_reuse = (int) value;
    else if (k == "tl_agg") { c->tl_agg_opt = (int) value; c->tl_ready = c->tl_agg_ready = false; drop_graph(c); }   // dofs per aggregate of FB_PRECOND_TWOLEVEL (0 = auto)
    else if (k == "asm_map") { c->asm_map_opt = (int) value; c->asm_map_ready = false; }   // 0: assemble by walking the rows (no 256 B / hexahedron map)
    else if (k == "charge_density") c->want_charge_density = (int) value;   // the reference's write_time(): the next assemble keeps rhs / dof_volume
    else if (k == "cell_grid") c->cell_grid = (int) value;    // 0: brute-force tetrahedron scan (read by the next fb_interp_initialize)  // 0: fb_import_mesh never takes the unchanged-topology path
    else if (k == "cg_p2p") c->cg_p2p = (int) value;          // read by the next partitioned fb_import_mesh
    else if (k == "cg_persistent_ctas") { c->pers_ctas = (int) value; c->pers_grid = 0; }
    else if (k == "cg_profile") c->cg_profile = std::max(0, std::min(4096, (int) value));
    else return c->fail(FB_ERR_ARG, "unknown option %s", key);
    return FB_OK;
}

int fb_synchronize(fb_ctx* c) { return sync_check(c, "fb_synchronize"); }

int fb_comm_unique_id(char* id128) {
    fb::Nccl& N = fb::Nccl::get();
    if (!N.ok) { g_create_error = std::string("NCCL unavailable: ") + N.why; return FB_ERR_CUDA; }
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return FB_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return FB_OK;
}

int fb_comm_init(fb_ctx* c, int rank, int world, const char* id128) {
    FB_REQUIRE(c, world >= 1 && rank >= 0 && rank < world && id128, "fb_comm_init: invalid rank / world");
    FB_REQUIRE(c, !c->mesh_ok, "fb_comm_init: must precede fb_import_mesh");
    c->rank = rank; c->world = world;
    if (world == 1) return FB_OK;
    fb::Nccl& N = fb::Nccl::get();
    if (!N.ok) return c->fail(FB_ERR_CUDA, "NCCL unavailable: %s", N.why);
    cudaSetDevice(c->device);
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclComm_t comm;
    FB_NCCL(c, N.CommInitRank(&comm, world, id, rank));
    c->nccl_comm = comm;
    FB_CUDA(c, c->d_red.alloc(8));
    return FB_OK;
}

// ---- host-only view of the partition (CPU tests of the N > 1 logic; no CUDA involved) ----
fb_ctx* fb_plan_create(int rank, int world) {
    fb_ctx* c = new fb_ctx();
    c->host_only = true; c->rank = rank; c->world = world;
    return c;
}
// host-only: mesh kind of the next fb_plan_import (0 = vacuum hexahedra, 1 = bulk hexahedra of the current / heat solvers)
int fb_plan_set_kind(fb_ctx* c, int kind) {
    FB_REQUIRE(c, c->host_only && (kind == 0 || kind == 1), "fb_plan_set_kind: needs a plan context and kind in {0, 1}");
    c->mesh_kind = kind;
    return FB_OK;
}
int fb_plan_phase1(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex, double* bbox6) {
    FB_REQUIRE(c, c->host_only, "fb_plan_phase1: not a plan context");
    const int rc = fb_host_partition_phase1(c, xyz, n_nodes, hex8, hex_marker, n_hex);
    if (rc) return rc;
    for (int d = 0; d < 3; ++d) { bbox6[d] = c->bb_mn[d]; bbox6[3 + d] = c->bb_mx[d]; }
    return FB_OK;
}
// host-only: the UN-partitioned import of fb_import_mesh (vertex compaction, orientation, boundary ids, first-touch
// numbering, sparsity) on a plan context; fb_plan_sizes / fb_plan_get / fb_plan_jds then describe the complete system
int fb_plan_import(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    FB_REQUIRE(c, c->host_only, "fb_plan_import: not a plan context");
    c->last_import_reused = fb_host_try_reuse(c, xyz, n_nodes, hex8, hex_marker, n_hex, c->mesh_kind);
    if (c->last_import_reused) return FB_OK;
    const int rc = fb_host_import_mesh(c, xyz, n_nodes, hex8, hex_marker, n_hex);
    if (rc) return rc;
    c->n_vert_global = c->n_vert; c->n_cells_global = c->n_cells; c->n_dofs_global = c->n_dofs;
    c->part_l2g.resize(c->n_vert); for (int v = 0; v < c->n_vert; ++v) c->part_l2g[v] = v;
    c->part_owner.assign(c->n_vert, 0);
    c->part_cell_g.resize(c->n_cells); for (int k = 0; k < c->n_cells; ++k) c->part_cell_g[k] = k;
    c->send_off.assign(c->world + 1, 0); c->recv_off.assign(c->world + 1, 0); c->send_idx.clear();
    return FB_OK;
}
int fb_plan_phase2(fb_ctx* c, const double* bbox6_global) {
    FB_REQUIRE(c, c->host_only, "fb_plan_phase2: not a plan context");
    for (int d = 0; d < 3; ++d) { c->bb_mn[d] = bbox6_global[d]; c->bb_mx[d] = bbox6_global[3 + d]; }
    return fb_host_import_phase2(c);
}
// sizes: n_rows (owned), n_cols, nnz, n_cells_local, n_send, n_ghost, n_vert_global, n_cells_global
int fb_plan_sizes(const fb_ctx* c, long* out8) {
    out8[0] = c->n_dofs; out8[1] = c->n_cols; out8[2] = c->nnz; out8[3] = c->n_cells; out8[4] = (long) c->send_idx.size();
    out8[5] = c->n_cols - c->n_dofs; out8[6] = c->n_vert_global; out8[7] = c->n_cells_global;
    return FB_OK;
}
int fb_plan_get(const fb_ctx* c, int* local2global, int* owner, int* send_off, int* send_idx, int* recv_off, int* rowptr, int* col,
                int* cells_dof, int* local_cell2global, int* copper_flag, int* top_flag) {
    if (local2global) std::copy(c->part_l2g.begin(), c->part_l2g.end(), local2global);
    if (owner) std::copy(c->part_owner.begin(), c->part_owner.end(), owner);
    if (send_off) std::copy(c->send_off.begin(), c->send_off.end(), send_off);
    if (send_idx) std::copy(c->send_idx.begin(), c->send_idx.end(), send_idx);
    if (recv_off) std::copy(c->recv_off.begin(), c->recv_off.end(), recv_off);
    if (rowptr) std::copy(c->rowptr.begin(), c->rowptr.end(), rowptr);
    if (col) std::copy(c->col.begin(), c->col.end(), col);
    if (cells_dof) std::copy(c->cells_dof.begin(), c->cells_dof.end(), cells_dof);
    if (local_cell2global) std::copy(c->part_cell_g.begin(), c->part_cell_g.end(), local_cell2global);
    if (copper_flag) { std::fill(copper_flag, copper_flag + c->n_cols, 0); for (int d : c->copper_dofs) copper_flag[d] = 1; }
    if (top_flag) { std::fill(top_flag, top_flag + c->n_cols, 0); for (int d : c->top_dofs) top_flag[d] = 1; }
    return FB_OK;
}
// host-only: the interpolator tables of fb_interp_initialize (bit-exact precompute of the reference's cell classes),
// for CPU parity tests.  The plan context imports the COMPLETE mesh here (no partition): call on a fresh fb_plan_create.
// Layouts: tet17 = {det0, d[4][4]} per tetrahedron, tri16 = {vert0, edge1, edge2, pvec, norm, maxd} per triangle,
// hex24 = f0..f7 per hexahedron.
int fb_plan_interp_tables(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex,
                          const int* node_marker, const int* tet4, const int* tet_nbr4, const int* tet_marker, int n_tet,
                          const int* tri3, const double* tri_norm3, int n_tri, const int* quad4, int n_quad,
                          double* tet17, double* tet_cent3, int* tet_mark, double* hex24, double* tri16, double* tri_cent3,
                          int* qtet10, int* qtri6) {
    FB_REQUIRE(c, c->host_only, "fb_plan_interp_tables: not a plan context");
    FB_REQUIRE(c, 4L * n_tet == n_hex, "fb_plan_interp_tables: expected 4 hexahedra per tetrahedron");
    const int rc = fb_host_import_mesh(c, xyz, n_nodes, hex8, hex_marker, n_hex);
    if (rc) return rc;
    fb_interp_tables T;
    fb_host_interp_tables(c, node_marker, tet4, tet_nbr4, tet_marker, n_tet, tri3, tri_norm3, n_tri, quad4, n_quad, T);
    static_assert(sizeof(fb::TetRec) == 17 * 8 && sizeof(fb::TriRec) == 16 * 8 && sizeof(fb::HexRec) == 24 * 8, "record layouts");
    if (tet17) memcpy(tet17, T.tet.data(), T.tet.size() * sizeof(fb::TetRec));
    if (tet_cent3) std::copy(T.tet_cent.begin(), T.tet_cent.end(), tet_cent3);
    if (tet_mark) std::copy(T.tet_mark.begin(), T.tet_mark.end(), tet_mark);
    if (hex24) memcpy(hex24, T.hex.data(), T.hex.size() * sizeof(fb::HexRec));
    if (tri16 && n_tri > 0) memcpy(tri16, T.tri.data(), T.tri.size() * sizeof(fb::TriRec));
    if (tri_cent3 && n_tri > 0) std::copy(T.tri_cent.begin(), T.tri_cent.end(), tri_cent3);
    if (qtet10) std::copy(T.qtet.begin(), T.qtet.end(), qtet10);
    if (qtri6 && n_tri > 0) std::copy(T.qtri.begin(), T.qtri.end(), qtri6);
    return FB_OK;
}

// host-only: the block-JDS tables of the HBM SpMV for the plan's sparsity (CPU tests of fb_host_jds_build)
int fb_plan_jds(fb_ctx* c, int R, int max_window, int sym, long* sizes6) {
    FB_REQUIRE(c, c->host_only && c->mesh_ok, "fb_plan_jds: needs a plan context after fb_plan_phase2");
    FB_REQUIRE(c, R == 128 || R == 256 || R == 512, "fb_plan_jds: R must be 128, 256 or 512");
    if (!fb_host_jds_build(c, R, max_window, sym != 0)) return c->fail(FB_ERR_ARG, "fb_plan_jds: a window exceeds max_window or a row is too long");
    sizes6[0] = c->jds_nb; sizes6[1] = c->jds_size; sizes6[2] = c->win_off[c->jds_nb]; sizes6[3] = c->jds_maxlen; sizes6[4] = c->win_max;
    sizes6[5] = (long) c->jds_jd.size();
    return FB_OK;
}
// host-only: the same tables with rows longer than `split` entries stored as chained segments (spmv_kernel 306)
int fb_plan_jds_split(fb_ctx* c, int R, int max_window, int split, long* sizes6) {
    FB_REQUIRE(c, c->host_only && c->mesh_ok, "fb_plan_jds_split: needs a plan context after fb_plan_phase2");
    FB_REQUIRE(c, R == 512 && split >= 8, "fb_plan_jds_split: R must be 512 and split >= 8");
    if (!fb_host_jds_build(c, R, max_window, false, split)) return c->fail(FB_ERR_ARG, "fb_plan_jds_split: a window exceeds max_window or a row is too long");
    sizes6[0] = c->jds_nb; sizes6[1] = c->jds_size; sizes6[2] = c->win_off[c->jds_nb]; sizes6[3] = c->jds_maxlen; sizes6[4] = c->win_max;
    sizes6[5] = (long) c->jds_jd.size();
    return FB_OK;
}
int fb_plan_jds_get_split(const fb_ctx* c, int* rowbeg, unsigned short* link) {
    if (rowbeg) std::copy(c->jds_rowbeg.begin(), c->jds_rowbeg.end(), rowbeg);
    if (link) std::copy(c->jds_link.begin(), c->jds_link.end(), link);
    return FB_OK;
}
int fb_plan_jds_get(const fb_ctx* c, unsigned short* perm, unsigned short* len, unsigned short* slot, int* jdp, int* jd, int* base,
                    unsigned short* col16, int* win_off, int* win_list) {
    if (perm) std::copy(c->jds_perm.begin(), c->jds_perm.end(), perm);
    if (len) std::copy(c->jds_len.begin(), c->jds_len.end(), len);
    if (slot) std::copy(c->jds_slot.begin(), c->jds_slot.end(), slot);
    if (jdp) std::copy(c->jds_jdp.begin(), c->jds_jdp.end(), jdp);
    if (jd) std::copy(c->jds_jd.begin(), c->jds_jd.end(), jd);
    if (base) std::copy(c->jds_base.begin(), c->jds_base.end(), base);
    if (col16) std::copy(c->col16.begin(), c->col16.begin() + c->jds_size, col16);
    if (win_off) std::copy(c->win_off.begin(), c->win_off.end(), win_off);
    if (win_list) std::copy(c->win_list.begin(), c->win_list.end(), win_list);
    return FB_OK;
}
void* fb_get_stream(fb_ctx* c) { return (void*) c->stream; }

// ------------------------------------------------------------------------------------------
static int import_mesh_impl(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex);

int fb_import_mesh(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    c->mesh_kind = 0;
    return import_mesh_impl(c, xyz, n_nodes, hex8, hex_marker, n_hex);
}

int fb_import_bulk_mesh(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    FB_REQUIRE(c, c->world == 1, "fb_import_bulk_mesh: the current / heat solvers run on un-partitioned meshes (native sizes)");
    c->mesh_kind = 1;
    const int rc = import_mesh_impl(c, xyz, n_nodes, hex8, hex_marker, n_hex);
    if (rc) return rc;
    // second system (the heat equation shares the sparsity of the current equation) and the per-face Neumann data
    FB_CUDA(c, c->d_val_other.alloc(c->nnz)); FB_CUDA(c, c->d_x_other.alloc(c->n_cols));
    FB_CUDA(c, c->d_face_bc.alloc(std::max(1, c->n_top_faces)));
    c->ch_active = 0; c->ch_setup_ok = false;
    c->ch_matrix_ok[0] = c->ch_matrix_ok[1] = c->ch_assembled[0] = c->ch_assembled[1] = false;
    return FB_OK;
}

static int import_mesh_impl(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex) {
    FB_REQUIRE(c, xyz && hex8 && hex_marker && n_nodes > 0 && n_hex > 0, "fb_import_mesh: empty mesh");
    FB_REQUIRE(c, c->fe_degree == 1 || (c->world == 1 && c->mesh_kind == 0), "fb_import_mesh: fe_degree 2 is provided for the un-partitioned field solver only");
    cudaSetDevice(c->device);
    c->last_import_reused = false;
    if (c->world == 1 && fb_host_try_reuse(c, xyz, n_nodes, hex8, hex_marker, n_hex, c->mesh_kind)) {
        // same connectivity, new coordinates: numbering, sparsity, block-JDS tables, persistent-CG slices, the
        // captured CG graph and every integer device array stay; the matrices and the interpolator tables depend on the
        // geometry and are rebuilt by the next assemble(true) / fb_interp_initialize
        c->setup_ok = c->assembled = c->matrix_ok = c->interp_ok = false;
        c->jds_val_dirty = true; c->cheb_lmax = 0; c->tl_ready = c->tl_agg_ready = false;
        const int n = c->n_cols;
        double* vx = (double*) c->pin_in.p;
        if (c->pin_in.bytes < 3 * (size_t) n * sizeof(double)) { FB_CUDA(c, c->pin_in.reserve(3 * (size_t) n * sizeof(double))); vx = (double*) c->pin_in.p; }
#pragma omp parallel for schedule(static)
        for (int d = 0; d < n; ++d) {
            const double* p = &c->xyz[3 * (size_t) c->vert2node[c->dof2vertex[d]]];
            vx[3 * (size_t) d] = p[0]; vx[3 * (size_t) d + 1] = p[1]; vx[3 * (size_t) d + 2] = p[2];
        }
        FB_CUDA(c, cudaMemcpyAsync(c->d_vxyz.p, vx, 3 * (size_t) n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaMemsetAsync(c->d_x.p, 0, n * sizeof(double), c->stream));
        c->last_import_reused = true;
        return sync_check(c, "fb_import_mesh");
    }
    c->setup_ok = c->assembled = c->matrix_ok = c->interp_ok = false;
    c->n_mesh_faces = c->n_mesh_edges = -1; c->d_vert_lastcell.release();
    drop_graph(c);
    cudaStream_t s = c->stream;
    int rc;
    if (c->world > 1) {
        // partitioned import: local sub-mesh, extremes of the boundary-face centres reduced over the ranks
        FB_REQUIRE(c, c->nccl_comm, "fb_import_mesh: fb_comm_init has not been called");
        p2p_release(c);     // peers' mappings of the previous mesh go first (the collective below orders this before any re-allocation)
        // a rank-local failure (e.g. "rank owns no vertices") must not leave the other ranks waiting in the collective:
        // the error flag travels with the extremes and every rank leaves with the same verdict
        const int rc_local = fb_host_partition_phase1(c, xyz, n_nodes, hex8, hex_marker, n_hex);
        const std::string err_local = c->err;
        double* hb = (double*) c->pin_out.p;                 // [max(mx), max(-mn), failed]
        for (int d = 0; d < 3; ++d) { hb[d] = rc_local ? -1e300 : c->bb_mx[d]; hb[3 + d] = rc_local ? -1e300 : -c->bb_mn[d]; }
        hb[6] = rc_local ? 1.0 : 0.0;
        FB_CUDA(c, cudaMemcpyAsync(c->d_red.p, hb, 7 * sizeof(double), cudaMemcpyHostToDevice, s));
        if ((rc = allreduce(c, c->d_red.p, 7, ncclMax))) return rc;
        FB_CUDA(c, cudaMemcpyAsync(hb, c->d_red.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
        if (hb[6] != 0.0)
            return rc_local ? c->fail(rc_local, "%s", err_local.c_str()) : c->fail(FB_ERR_MESH, "fb_import_mesh: the mesh import failed on another rank");
        for (int d = 0; d < 3; ++d) { c->bb_mx[d] = hb[d]; c->bb_mn[d] = -hb[3 + d]; }
        if ((rc = fb_host_import_phase2(c))) return rc;
    } else {
        if ((rc = fb_host_import_mesh(c, xyz, n_nodes, hex8, hex_marker, n_hex))) return rc;
        c->n_vert_global = c->n_vert; c->n_cells_global = c->n_cells; c->n_dofs_global = c->n_dofs;
    }
    const int n = c->n_cols;              // vectors indexed by column (owned rows + ghosts)
    // coordinates in DoF order
    const bool q2 = c->imported_degree == 2;
    std::vector<double> vxyz(3 * (size_t) n);
    if (q2) vxyz.swap(c->q2_xyz);         // support points of FE_Q(2), computed with the numbering
    else
    for (int d = 0; d < n; ++d) {
        const double* p = &c->xyz[3 * (size_t) c->vert2node[c->dof2vertex[d]]];
        vxyz[3 * (size_t) d] = p[0]; vxyz[3 * (size_t) d + 1] = p[1]; vxyz[3 * (size_t) d + 2] = p[2];
    }
    static const int FV[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
    std::vector<int> top;
    for (const auto& bf : c->bfaces)
        if (bf.id == (c->mesh_kind ? 2 : 8)) for (int k = 0; k < 4; ++k) top.push_back(c->cells_dof[8 * (size_t) bf.cell + FV[bf.face][k]]);
    FB_CUDA(c, c->d_vxyz.upload(vxyz, s));
    FB_CUDA(c, c->d_cells.upload(c->cells_dof, s));
    if (q2) { FB_CUDA(c, c->d_cells27.upload(c->cells27, s)); FB_CUDA(c, c->d_topfaces9.upload(c->topfaces9, s)); }
    FB_CUDA(c, c->d_rowptr.upload(c->rowptr, s));
    FB_CUDA(c, c->d_col.upload(c->col, s));
    c->n_rowblk = 0; c->rowblk_chunk = 0; c->win_cap = 0; c->jds_ready = false; c->jds_val_dirty = true;
    c->tl_ready = c->tl_agg_ready = false; c->asm_map_ready = false;
    c->pers_grid = 0;
    FB_CUDA(c, c->d_topfaces.upload(top, s));
    FB_CUDA(c, c->d_vertex2dof.upload(c->vertex2dof, s));
    FB_CUDA(c, c->d_cell2hex.upload(c->cell2hex, s));
    FB_CUDA(c, c->d_hex2cell.upload(c->hex2cell, s));
    {   // Dirichlet candidates (copper, then top) and the number of owned constrained rows in either anode mode
        std::vector<int> all(c->copper_dofs);
        all.insert(all.end(), c->top_dofs.begin(), c->top_dofs.end());
        if (all.empty()) all.push_back(0);
        FB_CUDA(c, c->d_bc_dofs.upload(all, s));
        std::vector<unsigned char> mark(c->n_cols, 0);
        for (int d : c->copper_dofs) mark[d] = 1;
        c->n_dirichlet_cu = (int) std::count(mark.begin(), mark.begin() + c->n_dofs, (unsigned char) 1);
        for (int d : c->top_dofs) mark[d] = 1;
        c->n_dirichlet_cu_top = (int) std::count(mark.begin(), mark.begin() + c->n_dofs, (unsigned char) 1);
    }
    FB_CUDA(c, c->d_val_save.alloc(c->nnz));
    FB_CUDA(c, c->d_diagpos.alloc(n));
    for (fb::DevBuf<double>* b : {&c->d_rhs, &c->d_x, &c->d_g, &c->d_d, &c->d_h, &c->d_dinv, &c->d_z, &c->d_w, &c->d_bcval})
        FB_CUDA(c, b->alloc(n));
    FB_CUDA(c, c->d_bcflag.alloc(n));
    FB_CUDA(c, cudaMemsetAsync(c->d_x.p, 0, n * sizeof(double), s));
    if (c->world > 1) {
        FB_CUDA(c, c->d_send_idx.upload(c->send_idx, s));
        FB_CUDA(c, c->d_sendbuf.alloc(std::max<size_t>(1, c->send_idx.size())));
        FB_CUDA(c, c->d_l2g.upload(c->part_l2g, s));
        std::vector<int> g2l(c->n_cells_global, -1);
        for (int lc = 0; lc < c->n_cells; ++lc) g2l[c->part_cell_g[lc]] = lc;
        FB_CUDA(c, c->d_gcell2local.upload(g2l, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
        if ((rc = p2p_setup(c))) return rc;
    }
    return sync_check(c, "fb_import_mesh");
}

int fb_get_sizes(const fb_ctx* c, long* out) {
    out[0] = c->n_dofs; out[1] = c->n_cells; out[2] = c->nnz; out[3] = c->n_vert;
    out[4] = (long) c->bfaces.size(); out[5] = c->n_top_faces; out[6] = c->n_dirichlet;
    return FB_OK;
}

int fb_get_cells27(const fb_ctx* c, int* cells27) {
    if (!c->mesh_ok || c->imported_degree != 2 || !cells27) return FB_ERR_ARG;
    std::copy(c->cells27.begin(), c->cells27.end(), cells27);
    return FB_OK;
}

int fb_poisson_setup(fb_ctx* c, double field, double potential, int anode_is_dirichlet) {
    FB_REQUIRE(c, c->mesh_ok, "fb_poisson_setup: no mesh imported");
    FB_REQUIRE(c, c->mesh_kind == 0, "fb_poisson_setup: the context holds a bulk mesh (fb_import_bulk_mesh)");
    cudaSetDevice(c->device);
    c->applied_field = field; c->applied_potential = potential; c->anode_dirichlet = anode_is_dirichlet;
    c->setup_ok = true; c->assembled = false; c->matrix_ok = false;
    FB_CUDA(c, cudaMemsetAsync(c->d_x.p, 0, c->n_cols * sizeof(double), c->stream));     // solution = dirichlet_bc_value (0)
    FB_CUDA(c, cudaMemsetAsync(c->d_rhs.p, 0, c->n_dofs * sizeof(double), c->stream));
    return FB_OK;
}

static int assemble_impl(fb_ctx* c, int first_time, const double* d_pxyz, const int* d_pcell, long n_parts, double charge_factor) {
    cudaStream_t s = c->stream;
    const int n = c->n_dofs;
    if (first_time || !c->matrix_ok) {
        // stiffness matrix -> val_save (the reference's system_matrix_save).  It is never modified afterwards: the
        // Dirichlet conditions are a mask (k_bc_prepare), so a later assemble(false) has no matrix work at all
        // (the reference copies the saved matrix back and eliminates again, PoissonSolver.cpp:157-159,178-192).
        FB_CUDA(c, cudaMemsetAsync(c->d_val_save.p, 0, c->nnz * sizeof(double), s));
        fb::launch_assemble_stiffness(c);
        // boundary values: copper = 0 (+ anode = V0 in Dirichlet mode); map semantics: later wins
        FB_CUDA(c, cudaMemsetAsync(c->d_bcflag.p, 0, c->n_cols * sizeof(int), s));
        FB_CUDA(c, cudaMemsetAsync(c->d_bcval.p, 0, c->n_cols * sizeof(double), s));
        fb::launch_set_bc(c, c->d_bc_dofs.p, (int) c->copper_dofs.size(), 0.0);
        if (c->anode_dirichlet) fb::launch_set_bc(c, c->d_bc_dofs.p + c->copper_dofs.size(), (int) c->top_dofs.size(), c->applied_potential);
        c->n_dirichlet = c->anode_dirichlet ? c->n_dirichlet_cu_top : c->n_dirichlet_cu;
        if (c->world > 1) {
            // a ghost dof may be constrained through a face this rank does not hold: take flag and value from the owner
            fb::launch_flags_to_double(c, c->d_z.p);
            int rc = halo_exchange(c, c->d_z.p);
            if (rc) return rc;
            fb::launch_double_to_ghost_flags(c, c->d_z.p);
            if ((rc = halo_exchange(c, c->d_bcval.p))) return rc;
        }
        fb::launch_bc_prepare(c);           // dinv (0 on constrained rows), diagpos
        c->jds_val_dirty = true;
        c->cheb_lmax = 0;                   // Gershgorin bound of the new matrix is computed by the next Chebyshev solve
        c->tl_ready = false;                // ... and the coarse matrix of the two-level preconditioner by the next such solve
        c->matrix_ok = true;
    }
    // right-hand side: Neumann faces (or nothing), space charge; constrained dofs take their value in the solution
    FB_CUDA(c, cudaMemsetAsync(c->d_rhs.p, 0, n * sizeof(double), s));
    if (!c->anode_dirichlet) fb::launch_neumann(c);
    if (n_parts > 0) {
        fb::launch_space_charge(c, n_parts, d_pxyz, d_pcell, charge_factor);
    }
    // PoissonSolver.cpp:196-207: when a file is about to be written (option "charge_density", the reference's write_time())
    // charge_density = rhs / dof_volume is kept, BEFORE the Dirichlet conditions touch the right-hand side; zeros otherwise
    c->rho_valid = false;
    if (c->want_charge_density && c->world == 1 && c->imported_degree == 1) {
        FB_CUDA(c, c->d_rho.alloc(n));
        fb::launch_charge_density(c, c->d_h.p);          // d_h is scratch until the solve
        c->rho_valid = true;
    }
    fb::launch_bc_solution(c);          // (partitioned: the solve exchanges the halo of x before the initial residual)
    c->assembled = true;
    return FB_OK;
}

int fb_poisson_assemble_dev(fb_ctx* c, int first_time, const double* pxyz_dev, const int* pcell_dev, long n, double cf) {
    FB_REQUIRE(c, c->setup_ok, "fb_poisson_assemble: call fb_poisson_setup first");
    cudaSetDevice(c->device);
    return assemble_impl(c, first_time, pxyz_dev, pcell_dev, n, cf);
}

int fb_poisson_assemble(fb_ctx* c, int first_time, const double* pxyz, const int* pcell, long n, double cf) {
    FB_REQUIRE(c, c->setup_ok, "fb_poisson_assemble: call fb_poisson_setup first");
    cudaSetDevice(c->device);
    if (n > 0) {
        FB_REQUIRE(c, pxyz && pcell, "fb_poisson_assemble: particle arrays missing");
        FB_CUDA(c, c->d_pts.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_cellsA.alloc(n));
        FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, pxyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(c, cudaMemcpyAsync(c->d_cellsA.p, pcell, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    int rc = assemble_impl(c, first_time, c->d_pts.p, c->d_cellsA.p, n, cf);
    if (rc) return rc;
    return sync_check(c, "fb_poisson_assemble");
}

int fb_poisson_solve(fb_ctx* c, int max_iter, double abs_tol, int precond, int* n_iter, double* final_residual) {
    FB_REQUIRE(c, c->assembled, "fb_poisson_solve: system not assembled");
    FB_REQUIRE(c, precond == FB_PRECOND_JACOBI || precond == FB_PRECOND_CHEBYSHEV || precond == FB_PRECOND_TWOLEVEL,
               "fb_poisson_solve: preconditioner must be FB_PRECOND_JACOBI, FB_PRECOND_CHEBYSHEV or FB_PRECOND_TWOLEVEL (SSOR is sequential and not provided)");
    FB_REQUIRE(c, precond != FB_PRECOND_CHEBYSHEV || c->world == 1, "fb_poisson_solve: FB_PRECOND_CHEBYSHEV runs on un-partitioned meshes only");
    FB_REQUIRE(c, precond != FB_PRECOND_TWOLEVEL || c->world == 1 || c->p2p_ready,
               "fb_poisson_solve: FB_PRECOND_TWOLEVEL on a partitioned mesh needs the peer-mapped iteration (fb_comm_mode 2)");
    const bool tl = precond == FB_PRECOND_TWOLEVEL;
    c->tl_active = false;
    const bool cheb = precond == FB_PRECOND_CHEBYSHEV && c->cheb_degree >= 2;        // degree 1 is Jacobi up to a scale factor
    c->cheb_active = false;
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    fb::CgScalars init; memset(&init, 0, sizeof init);
    init.tol2 = abs_tol * abs_tol; init.max_iter = max_iter;
    fb::CgScalars* h = (fb::CgScalars*) c->pin_out.p;
    *h = init;
    long spmv = 1;
    c->prof_samples = 0; c->prof_spmv_ms = c->prof_vec_ms = 0;
    const bool persistent = c->world == 1 && c->cg_profile == 0 && !cheb && !tl && fb::persistent_eligible(c);
    if (c->world > 1) { h->red = c->d_red.p; if (c->p2p_ready) h->p2p = c->d_p2p.p; }
    FB_CUDA(c, cudaEventRecord(c->ev0, s));
    FB_CUDA(c, cudaMemcpyAsync(c->d_cg.p, h, sizeof(fb::CgScalars), cudaMemcpyHostToDevice, s));
    c->last_kernel = -2;
    if (persistent) {
        // native meshes: the whole solve is ONE cooperative launch (matrix slice resident in shared memory)
        FB_CUDA(c, fb::launch_cg_persistent(c));
        FB_CUDA(c, cudaMemcpyAsync(h, c->d_cg.p, sizeof(fb::CgScalars), cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
        spmv += h->it;
        if (c->cg_debug) {          // per-phase SM cycles of CTAs 0..3 (diagnostics)
            long long t[32];
            cudaMemcpy(t, c->d_dbg.p, sizeof t, cudaMemcpyDeviceToHost);
            const long long na = std::max(1LL, t[29]);
            fprintf(stderr, "[fb cg_debug] allreduce on cta 0 (avg cycles over %lld calls): block-reduce+sync %lld | publish+release %lld | poll %lld | fence+read %lld | exit syncs %lld\n",
                    t[29], t[24] / na, t[25] / na, t[26] / na, t[27] / na, t[28] / na);
            for (int b = 0; b < 4; ++b)
                fprintf(stderr, "[fb cg_debug] cta %d it %d cycles/it: spmv %lld | reduce(d.h) %lld | update %lld | reduce(g.g) %lld | direction %lld | barrier(d) %lld\n",
                        b, h->it, t[6 * b] / std::max(1, h->it), t[6 * b + 1] / std::max(1, h->it), t[6 * b + 2] / std::max(1, h->it),
                        t[6 * b + 3] / std::max(1, h->it), t[6 * b + 4] / std::max(1, h->it), t[6 * b + 5] / std::max(1, h->it));
        }
    } else {
        int lanes = fb::choose_lanes(c);
        if ((cheb || tl) && lanes >= 310) lanes = 304;      // the Chebyshev steps multiply by the FULL matrix: the symmetric (lower-triangle) layout cannot serve them
        while (lanes >= 300) {                     // block-JDS SpMV: tables once per mesh, values once per assemble
            const bool sym = lanes >= 310;         // 310/311: symmetric layout (lower triangle only)
            const int R = sym ? (lanes == 311 ? 256 : 512) : ((lanes == 301) ? 128 : ((lanes >= 302 && lanes <= 308) ? 512 : 256));
            const int split = (lanes >= 306 && lanes <= 308) ? c->spmv_split : 0;
            const int pad = lanes == 308 ? 8 : 2;
            if (!c->jds_ready || c->jds_R != R || c->jds_sym != sym || c->jds_split != split || c->jds_pad != pad) {
                drop_graph(c);
                // window capacity: shared memory holds the input window (and, symmetric layout, its accumulators)
                if (fb_host_jds_build(c, R, sym ? 6144 : 8192, sym, split, pad)) {
                    c->win_cap = (c->win_max + 15) & ~15;
                    if (sym) FB_CUDA(c, c->d_diag.alloc(c->n_dofs));
                    FB_CUDA(c, c->d_col16.upload(c->col16, s));
                    FB_CUDA(c, c->d_win_off.upload(c->win_off, s)); FB_CUDA(c, c->d_win_list.upload(c->win_list, s));
                    FB_CUDA(c, c->d_jds_perm.upload(c->jds_perm, s)); FB_CUDA(c, c->d_jds_len.upload(c->jds_len, s));
                    FB_CUDA(c, c->d_jds_slot.upload(c->jds_slot, s));
                    FB_CUDA(c, c->d_jds_jdp.upload(c->jds_jdp, s)); FB_CUDA(c, c->d_jds_jd.upload(c->jds_jd, s));
                    FB_CUDA(c, c->d_jds_base.upload(c->jds_base, s));
                    if (split > 0) { FB_CUDA(c, c->d_jds_rowbeg.upload(c->jds_rowbeg, s)); FB_CUDA(c, c->d_jds_link.upload(c->jds_link, s)); }
                    FB_CUDA(c, c->d_val_jds.alloc((size_t) c->jds_size + 8));
                    FB_CUDA(c, c->d_val_jds.zero(s));                  // padding entries stay 0
                    FB_CUDA(c, cudaStreamSynchronize(s));
                    std::vector<unsigned short>().swap(c->col16);      // host copies no longer needed
                    std::vector<int>().swap(c->win_list);
                    c->jds_ready = true; c->jds_val_dirty = true;
                    c->rowblk_chunk = 0; c->n_rowblk = 0;              // col16 / windows now belong to the JDS layout
                } else if (sym) {
                    lanes = 304; continue;         // a window + its accumulators exceed shared memory: full block-JDS
                } else {
                    lanes = 100; break;            // window too large for shared memory: plain streaming kernel
                }
            }
            if (c->jds_val_dirty) { fb::launch_csr_to_jds(c); c->jds_val_dirty = false; }
            break;
        }
        if (lanes == 0 || (lanes >= 100 && lanes < 300)) {          // streaming SpMV: (re)build its row blocks for the chosen variant
            if (lanes >= 200) c->jds_ready = false;                  // the windowed variants reuse the col16 / window buffers
            int chunk, maxrows;
            fb::stream_block_shape(lanes, chunk, maxrows);
            if (c->rowblk_chunk != chunk || c->rowblk_maxrows != maxrows || c->n_rowblk == 0) {
                drop_graph(c);
                c->win_cap = 0;
                if (fb_host_row_blocks(c, chunk, maxrows)) FB_CUDA(c, c->d_rowblk.upload(c->rowblk, s));
                else lanes = 32;                   // a row longer than the chunk: per-row kernel
            }
            if (lanes >= 200 && c->win_cap == 0) { // windowed variant: column windows + 16-bit window positions
                if (fb_host_col_windows(c, 6144)) {
                    c->win_cap = (c->win_max + 15) & ~15;
                    FB_CUDA(c, c->d_col16.upload(c->col16, s));
                    FB_CUDA(c, c->d_win_off.upload(c->win_off, s));
                    FB_CUDA(c, c->d_win_list.upload(c->win_list, s));
                    FB_CUDA(c, cudaStreamSynchronize(s));
                    std::vector<unsigned short>().swap(c->col16);      // host copy no longer needed
                } else {
                    lanes = (lanes == 201 || lanes == 204) ? 103 : (lanes == 202 ? 101 : (lanes == 203 ? 100 : 0));   // same block shape, plain gather
                }
            }
        }
        c->last_kernel = lanes;
        if (c->world > 1 && !c->p2p_ready) {
            // ---- partitioned CG, NCCL mode (fallback when the peers' memory cannot be mapped): halo exchange of the SpMV input over NVLink (NCCL point-to-point) and one
            //      all-reduce per dot-product pair; same operation order as on one GPU ----
            int rc;
            if ((rc = halo_exchange(c, c->d_x.p))) return rc;
            fb::launch_cg_init_spmv(c, lanes);
            if ((rc = allreduce(c, c->d_red.p, 2, ncclSum))) return rc;
            fb::launch_cg_scalars(c, 0);
            fb::launch_cg_init_direction(c);
            const int check_every = 16;
            const bool prof = c->cg_profile > 0;
            if (prof) while ((int) c->prof_ev.size() < 3 * c->cg_profile) { cudaEvent_t e; FB_CUDA(c, cudaEventCreate(&e)); c->prof_ev.push_back(e); }
            auto iteration = [&](int sample) -> int {          // one CG iteration on the stream (sample >= 0: bracketed by events)
                int rc2;
                if (sample >= 0) FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * sample], s));
                if ((rc2 = halo_exchange(c, c->d_d.p))) return rc2;
                fb::launch_cg_spmv(c, lanes);
                if ((rc2 = allreduce(c, c->d_red.p, 2, ncclSum))) return rc2;
                fb::launch_cg_scalars(c, 1);
                if (sample >= 0) FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * sample + 1], s));
                fb::launch_cg_update_only(c);
                if ((rc2 = allreduce(c, c->d_red.p, 2, ncclSum))) return rc2;
                fb::launch_cg_scalars(c, 2);
                fb::launch_cg_direction_only(c);
                if (sample >= 0) FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * sample + 2], s));
                return FB_OK;
            };
            // Host-issued loop, check_every iterations between two looks at the device-resident `done` flag.  (Capturing
            // the iteration -- kernels + grouped ncclSend/ncclRecv + all-reduces -- into a CUDA graph was tried to remove
            // the ~0.1 ms of launch gaps per iteration seen at 4 GPUs; the capture dead-locked on 2 x B200 with NCCL
            // 2.28.9 and was backed out: DESIGN.md section 4.)
            int n_prof = 0, launched = 0;
            while (true) {
                FB_CUDA(c, cudaMemcpyAsync(h, c->d_cg.p, sizeof(fb::CgScalars), cudaMemcpyDeviceToHost, s));
                FB_CUDA(c, cudaStreamSynchronize(s));
                if (h->done) break;
                for (int i = 0; i < check_every; ++i, ++launched) {
                    const bool sample = prof && launched < c->cg_profile;
                    if ((rc = iteration(sample ? launched : -1))) return rc;
                    if (sample) n_prof = launched + 1;
                }
                spmv += check_every;
            }
            if (n_prof > 0) {
                const int live = std::min(n_prof, h->it);
                for (int i = 0; i < live; ++i) {
                    float a = 0, b = 0;
                    cudaEventElapsedTime(&a, c->prof_ev[3 * i], c->prof_ev[3 * i + 1]);
                    cudaEventElapsedTime(&b, c->prof_ev[3 * i + 1], c->prof_ev[3 * i + 2]);
                    c->prof_spmv_ms += a; c->prof_vec_ms += b;
                }
                c->prof_samples = live;
                if (live > 0) { c->prof_spmv_ms /= live; c->prof_vec_ms /= live; }
            }
        } else {
            int graph_key = lanes;
            if (tl) {
                const int rc = fb::tl_prepare(c);
                if (rc) return rc;
                c->tl_active = true;
                graph_key = lanes + 100000;
                FB_CUDA(c, cudaMemcpyAsync(c->d_cg.p, h, sizeof(fb::CgScalars), cudaMemcpyHostToDevice, s));   // tl_prepare used the stream
            }
            if (cheb) {
                FB_CUDA(c, fb::cheb_prepare(c, lanes));
                c->cheb_active = true;
                graph_key = lanes + 1000 * c->cheb_k;
                FB_CUDA(c, cudaMemcpyAsync(c->d_cg.p, h, sizeof(fb::CgScalars), cudaMemcpyHostToDevice, s));   // cheb_prepare may have used the stream
            }
            // partitioned, peer-mapped mode: the initial residual needs the ghosts of x (NCCL, once per solve); inside the
            // iteration the halo of d and the two all-reduces travel over the peer mappings, so the loop below -- graph
            // capture included -- is the single-GPU one with k_pack_p2p in front of every SpMV
            if (c->world > 1) {
                const int rc = halo_exchange(c, c->d_x.p); if (rc) return rc;
                fb::launch_cg_init_spmv(c, lanes); fb::launch_cg_scalars(c, 0);
                if (tl) fb::launch_tl_init_tail(c); else fb::launch_cg_init_direction(c);
            } else
                fb::launch_cg_init(c, lanes);
            // CUDA graph of cg_graph_iters iterations; kernels become no-ops once cgs->done is set
            if (!c->cg_graph || c->cg_graph_n != c->cg_graph_iters || c->cg_graph_precond != graph_key) {
                drop_graph(c);
                cudaGraph_t graph;
                const long before = c->launches;
                FB_CUDA(c, cudaStreamSynchronize(s));
                FB_CUDA(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
                for (int i = 0; i < c->cg_graph_iters; ++i) fb::launch_cg_iteration(c, lanes);
                FB_CUDA(c, cudaStreamEndCapture(s, &graph));
                c->launches = before;     // captured, not launched
                FB_CUDA(c, cudaGraphInstantiate(&c->cg_graph, graph, 0));
                cudaGraphDestroy(graph);
                c->cg_graph_n = c->cg_graph_iters; c->cg_graph_precond = graph_key;
            }
            // optional profile: the first cg_profile iterations run un-graphed, each bracketed by CUDA
            // events on this stream (SpMV+dot | vector updates), for the live roofline figures of bench.py
            int n_prof = 0;
            if (c->cg_profile > 0) {
                while ((int) c->prof_ev.size() < 3 * c->cg_profile) {
                    cudaEvent_t e; FB_CUDA(c, cudaEventCreate(&e)); c->prof_ev.push_back(e);
                }
                n_prof = std::min(c->cg_profile, std::max(0, max_iter));
                for (int i = 0; i < n_prof; ++i) {
                    FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * i], s));
                    if (c->world > 1) {
                        fb::launch_pack_p2p(c, c->d_d.p); fb::launch_cg_spmv(c, lanes); fb::launch_cg_scalars(c, 1);
                        FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * i + 1], s));
                        if (tl) fb::launch_tl_vectors(c);
                        else { fb::launch_cg_update_only(c); fb::launch_cg_scalars(c, 2); fb::launch_cg_direction_only(c); }
                        FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * i + 2], s));
                        continue;
                    }
                    fb::launch_cg_spmv(c, lanes);
                    FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * i + 1], s));
                    fb::launch_cg_vectors(c);
                    FB_CUDA(c, cudaEventRecord(c->prof_ev[3 * i + 2], s));
                }
                spmv += n_prof;
            }
            FB_CUDA(c, cudaMemcpyAsync(h, c->d_cg.p, sizeof(fb::CgScalars), cudaMemcpyDeviceToHost, s));
            FB_CUDA(c, cudaStreamSynchronize(s));
            if (n_prof > 0) {
                const int live = std::min(n_prof, h->it);      // launches after convergence are no-ops: not sampled
                for (int i = 0; i < live; ++i) {
                    float a = 0, b = 0;
                    cudaEventElapsedTime(&a, c->prof_ev[3 * i], c->prof_ev[3 * i + 1]);
                    cudaEventElapsedTime(&b, c->prof_ev[3 * i + 1], c->prof_ev[3 * i + 2]);
                    c->prof_spmv_ms += a; c->prof_vec_ms += b;
                }
                c->prof_samples = live;
                if (live > 0) { c->prof_spmv_ms /= live; c->prof_vec_ms /= live; }
            }
            while (!h->done) {
                FB_CUDA(c, cudaGraphLaunch(c->cg_graph, s));
                c->launches += (cheb ? 1L + 2L * c->cheb_k : (tl ? (c->world > 1 ? 8L : 5L) : (c->world > 1 ? 6L : 3L))) * c->cg_graph_n;
                spmv += (long) (cheb ? c->cheb_k : 1) * c->cg_graph_n;
                FB_CUDA(c, cudaMemcpyAsync(h, c->d_cg.p, sizeof(fb::CgScalars), cudaMemcpyDeviceToHost, s));
                FB_CUDA(c, cudaStreamSynchronize(s));
            }
        }
    }
    FB_CUDA(c, cudaEventRecord(c->ev1, s));
    FB_CUDA(c, cudaEventSynchronize(c->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_solve_ms = ms; c->last_iters = h->it; c->last_spmv = spmv;
    if (h->done == 3) return c->fail(FB_ERR_CUDA, "fb_poisson_solve: a peer GPU did not answer within the spin limit (iteration %d)", h->it);
    if (n_iter) *n_iter = (h->done == 1) ? h->it : -h->it;
    if (final_residual) *final_residual = std::sqrt(h->res2);
    return sync_check(c, "fb_poisson_solve");
}

int fb_last_solve_stats(const fb_ctx* c, double* ms, int* it, long* spmv) {
    if (ms) *ms = c->last_solve_ms; if (it) *it = c->last_iters; if (spmv) *spmv = c->last_spmv;
    return FB_OK;
}

int fb_last_solve_kernel(const fb_ctx* c) { return c->last_kernel; }

int fb_last_solve_profile(const fb_ctx* c, double* spmv_ms, double* vec_ms, int* n_samples) {
    if (spmv_ms) *spmv_ms = c->prof_spmv_ms; if (vec_ms) *vec_ms = c->prof_vec_ms; if (n_samples) *n_samples = c->prof_samples;
    return FB_OK;
}

static int export_by_vertex(fb_ctx* c, const double* d_src, double* out) {
    FB_CUDA(c, c->d_sol.alloc(c->n_vert));
    fb::launch_gather(c, c->n_vert, c->d_vertex2dof.p, d_src, c->d_sol.p);
    FB_CUDA(c, cudaMemcpyAsync(out, c->d_sol.p, c->n_vert * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "export");
}

int fb_export_solution(fb_ctx* c, double* phi_vertex) {
    FB_REQUIRE(c, c->mesh_ok && phi_vertex, "fb_export_solution: no mesh");
    cudaSetDevice(c->device);
    if (c->world > 1) {
        // partitioned: every rank scatters its owned values into a zeroed global vertex array, summed over the ranks
        const size_t ng = (size_t) c->n_vert_global;
        FB_CUDA(c, c->d_sol.alloc(ng));
        FB_CUDA(c, cudaMemsetAsync(c->d_sol.p, 0, ng * sizeof(double), c->stream));
        fb::launch_scatter(c, c->n_dofs, c->d_l2g.p, c->d_x.p, c->d_sol.p);
        int rc = allreduce(c, c->d_sol.p, (int) ng, ncclSum);
        if (rc) return rc;
        FB_CUDA(c, cudaMemcpyAsync(phi_vertex, c->d_sol.p, ng * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        return sync_check(c, "fb_export_solution");
    }
    return export_by_vertex(c, c->d_x.p, phi_vertex);
}

// rank, world, owned rows, local columns, local nnz, local cells, halo send count, ghost count, global vertices, global cells
int fb_get_partition(const fb_ctx* c, long* out10) {
    out10[0] = c->rank; out10[1] = c->world; out10[2] = c->n_dofs; out10[3] = c->n_cols; out10[4] = c->nnz; out10[5] = c->n_cells;
    out10[6] = (long) c->send_idx.size(); out10[7] = c->n_cols - c->n_dofs; out10[8] = c->n_vert_global; out10[9] = c->n_cells_global;
    return FB_OK;
}

// 1 when the last fb_import_mesh / fb_import_bulk_mesh found the connectivity unchanged and only refreshed the geometry
int fb_last_import_reused(const fb_ctx* c) { return c->last_import_reused ? 1 : 0; }

// 0 = one GPU, 1 = partitioned with NCCL inside the iteration, 2 = partitioned, peer-mapped iteration (CUDA IPC over NVLink)
int fb_comm_mode(const fb_ctx* c) { return c->world <= 1 ? 0 : (c->p2p_ready ? 2 : 1); }

int fb_export_charge_dens(fb_ctx* c, double* rho_vertex) {
    // PoissonSolver.cpp:198-207: charge_density is only filled when a file is being written
    // (outside the hot path); otherwise it is reinit'ed to zeros, which is what export returns.
    FB_REQUIRE(c, c->mesh_ok && rho_vertex, "fb_export_charge_dens: no mesh");
    if (c->rho_valid) { cudaSetDevice(c->device); return export_by_vertex(c, c->d_rho.p, rho_vertex); }
    std::fill(rho_vertex, rho_vertex + c->n_vert, 0.0);
    return FB_OK;
}

int fb_export_solution_grad(fb_ctx* c, double* grad3) {
    FB_REQUIRE(c, c->mesh_ok && grad3, "fb_export_solution_grad: no mesh");
    FB_REQUIRE(c, c->world == 1, "fb_export_solution_grad: un-partitioned meshes only");
    FB_REQUIRE(c, c->imported_degree == 1, "fb_export_solution_grad: fe_degree 1 only");
    cudaSetDevice(c->device);
    if (c->d_vert_lastcell.n < (size_t) c->n_vert || c->d_vert_lastcell.p == nullptr) {
        std::vector<int> lc;
        fb_host_vertex_lastcell(c, lc);
        FB_CUDA(c, c->d_vert_lastcell.upload(lc, c->stream));
        FB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    FB_CUDA(c, c->d_sol.alloc(3 * (size_t) c->n_vert));
    fb::launch_solution_grad(c, c->d_sol.p);
    FB_CUDA(c, cudaMemcpyAsync(grad3, c->d_sol.p, 3 * (size_t) c->n_vert * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_export_solution_grad");
}

int fb_get_mesh_counts(fb_ctx* c, long* n_faces, long* n_edges) {
    FB_REQUIRE(c, c->mesh_ok, "fb_get_mesh_counts: no mesh");
    if (c->n_mesh_edges < 0) {
        c->n_mesh_faces = (6L * c->n_cells + (long) c->bfaces.size()) / 2;      // every interior face is shared by two cells
        c->n_mesh_edges = fb_host_count_edges(c);
    }
    if (n_faces) *n_faces = c->n_mesh_faces;
    if (n_edges) *n_edges = c->n_mesh_edges;
    return FB_OK;
}

int fb_get_solver_mesh(fb_ctx* c, double* xyz_vertex, int* cells_ucd8) {
    FB_REQUIRE(c, c->mesh_ok, "fb_get_solver_mesh: no mesh");
    if (xyz_vertex)
        for (int v = 0; v < c->n_vert; ++v)
            for (int d = 0; d < 3; ++d) xyz_vertex[3 * (size_t) v + d] = c->xyz[3 * (size_t) c->vert2node[v] + d];
    if (cells_ucd8)
        for (int ce = 0; ce < c->n_cells; ++ce)
            for (int k = 0; k < 8; ++k) cells_ucd8[8 * (size_t) ce + k] = c->node2vert[c->hex8[8 * (size_t) c->cell2hex[ce] + k]];
    return FB_OK;
}

int fb_import_solution(fb_ctx* c, const double* phi_vertex) {
    FB_REQUIRE(c, c->mesh_ok && phi_vertex, "fb_import_solution: no mesh");
    cudaSetDevice(c->device);
    FB_CUDA(c, c->d_sol.alloc(c->n_vert));
    FB_CUDA(c, cudaMemcpyAsync(c->d_sol.p, phi_vertex, c->n_vert * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    fb::launch_scatter(c, c->n_vert, c->d_vertex2dof.p, c->d_sol.p, c->d_x.p);
    return sync_check(c, "fb_import_solution");
}

int fb_check_limits(fb_ctx* c, double lo, double hi, int* bad, double* mn, double* mx) {
    FB_REQUIRE(c, c->mesh_ok, "fb_check_limits: no mesh");
    cudaSetDevice(c->device);
    fb::launch_minmax(c);
    if (c->world > 1) {
        FB_NCCL(c, fb::Nccl::get().AllReduce(c->d_minmax.p, c->d_minmax.p, 1, ncclDouble, ncclMin, (ncclComm_t) c->nccl_comm, c->stream));
        FB_NCCL(c, fb::Nccl::get().AllReduce(c->d_minmax.p + 1, c->d_minmax.p + 1, 1, ncclDouble, ncclMax, (ncclComm_t) c->nccl_comm, c->stream));
    }
    double* h = (double*) c->pin_out.p;
    FB_CUDA(c, cudaMemcpyAsync(h, c->d_minmax.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    int rc = sync_check(c, "fb_check_limits");
    if (rc) return rc;
    if (mn) *mn = h[0]; if (mx) *mx = h[1];
    if (bad) *bad = (h[0] < lo || h[1] > hi);
    return FB_OK;
}

int fb_get_cell_volumes(fb_ctx* c, double* vol) {
    FB_REQUIRE(c, c->mesh_ok && vol, "fb_get_cell_volumes: no mesh");
    cudaSetDevice(c->device);
    fb::DevBuf<double> v;
    FB_CUDA(c, v.alloc(c->n_cells));
    fb::launch_cell_volumes(c, v.p);
    FB_CUDA(c, cudaMemcpyAsync(vol, v.p, c->n_cells * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_get_cell_volumes");
}

int fb_get_system(fb_ctx* c, int* rowptr, int* col, double* val, double* val_save, double* rhs, double* sol,
                  int* vertex2dof, int* vertex2node) {
    FB_REQUIRE(c, c->mesh_ok, "fb_get_system: no mesh");
    cudaSetDevice(c->device);
    if (rowptr) std::copy(c->rowptr.begin(), c->rowptr.end(), rowptr);
    if (col) std::copy(c->col.begin(), c->col.end(), col);
    if (vertex2dof) std::copy(c->vertex2dof.begin(), c->vertex2dof.end(), vertex2dof);
    if (vertex2node) std::copy(c->vert2node.begin(), c->vert2node.end(), vertex2node);
    cudaStream_t s = c->stream;
    if (val_save) FB_CUDA(c, cudaMemcpyAsync(val_save, c->d_val_save.p, c->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (val || rhs) {
        // the solver keeps K and the raw right-hand side and applies the Dirichlet conditions as a mask; the ELIMINATED
        // matrix and the lifted right-hand side of the reference (MatrixTools::apply_boundary_values) are produced here,
        // on demand, for the parity tests that read them back
        FB_REQUIRE(c, c->assembled, "fb_get_system: val / rhs need an assembled system");
        fb::DevBuf<double> tv, tr, tl, td; fb::DevBuf<int> tp;
        FB_CUDA(c, tv.alloc(c->nnz)); FB_CUDA(c, tr.alloc(c->n_dofs)); FB_CUDA(c, tl.alloc(c->n_dofs)); FB_CUDA(c, td.alloc(c->n_dofs));
        FB_CUDA(c, tp.alloc(c->n_dofs));
        fb::launch_materialize_eliminated(c, tv.p, tr.p, tl.p, td.p, tp.p);
        if (val) FB_CUDA(c, cudaMemcpyAsync(val, tv.p, c->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (rhs) FB_CUDA(c, cudaMemcpyAsync(rhs, tr.p, c->n_dofs * sizeof(double), cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
    }
    if (sol) FB_CUDA(c, cudaMemcpyAsync(sol, c->d_x.p, c->n_dofs * sizeof(double), cudaMemcpyDeviceToHost, s));
    return sync_check(c, "fb_get_system");
}

// ------------------------------------------------------------------------------------------
// CurrentHeatSolver<3> on the bulk mesh (src/CurrentHeatSolver.cpp).  Two systems share one sparsity
// pattern and the CG engine of fb_poisson_solve: `which` 0 = CurrentSolver (Laplace, sigma = 1), 1 = HeatSolver.
// The engine reads d_val_save / d_x; the system that is not being worked on is parked in d_val_other / d_x_other.
// ------------------------------------------------------------------------------------------
static void ch_activate(fb_ctx* c, int which) {
    if (c->ch_active == which) return;
    std::swap(c->d_val_save.p, c->d_val_other.p); std::swap(c->d_val_save.n, c->d_val_other.n);
    std::swap(c->d_x.p, c->d_x_other.p); std::swap(c->d_x.n, c->d_x_other.n);
    c->ch_active = which;
    drop_graph(c);                      // captured launches hold the old pointers
    c->jds_val_dirty = true; c->cheb_lmax = 0; c->tl_ready = false;
    c->assembled = false;               // the Dirichlet mask / dinv / rhs belong to the system assembled last
}
static double* ch_solution(fb_ctx* c, int which) { return which == c->ch_active ? c->d_x.p : c->d_x_other.p; }

int fb_ch_set_physics(fb_ctx* c, const double* T, const double* rho, int n, double lorentz) {
    FB_REQUIRE(c, T && rho && n >= 2, "fb_ch_set_physics: the resistivity table needs at least two rows");
    for (int i = 1; i < n; ++i) FB_REQUIRE(c, T[i] > T[i - 1], "fb_ch_set_physics: temperatures must increase");
    cudaSetDevice(c->device);
    FB_CUDA(c, c->d_res_T.upload(T, n, c->stream)); FB_CUDA(c, c->d_res_rho.upload(rho, n, c->stream));
    c->ch_n_table = n; c->ch_lorentz = lorentz;
    return sync_check(c, "fb_ch_set_physics");
}

int fb_ch_setup(fb_ctx* c, double T_ambient) {
    FB_REQUIRE(c, c->mesh_ok && c->mesh_kind == 1, "fb_ch_setup: no bulk mesh imported (fb_import_bulk_mesh)");
    cudaSetDevice(c->device);
    ch_activate(c, 0);
    c->ch_T_ambient = T_ambient;
    // DealSolver::setup_system: solution = dirichlet_bc_value (0 for the current, T_ambient for the heat equation)
    FB_CUDA(c, cudaMemsetAsync(c->d_x.p, 0, c->n_cols * sizeof(double), c->stream));
    FB_CUDA(c, c->d_sol.alloc(c->n_cols));
    std::vector<double> t(c->n_cols, T_ambient);
    FB_CUDA(c, cudaMemcpyAsync(c->d_x_other.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->ch_setup_ok = true; c->setup_ok = true; c->assembled = false;
    c->ch_matrix_ok[0] = c->ch_matrix_ok[1] = c->ch_assembled[0] = c->ch_assembled[1] = false;
    return FB_OK;
}

static int ch_face_data(fb_ctx* c, const double* face_bc, int n_faces, const char* who) {
    if (n_faces != c->n_top_faces) return c->fail(FB_ERR_ARG, "%s: %d face values for %d copper_surface faces", who, n_faces, c->n_top_faces);
    if (n_faces > 0) {
        if (!face_bc) return c->fail(FB_ERR_ARG, "%s: face data missing", who);
        FB_CUDA(c, cudaMemcpyAsync(c->d_face_bc.p, face_bc, n_faces * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    return FB_OK;
}

// boundary part shared by the two assemblies: Neumann faces, Dirichlet mask on copper_bottom, constrained dofs of the solution
static int ch_finish_assembly(fb_ctx* c, int which, double dirichlet_value) {
    cudaStream_t s = c->stream;
    fb::launch_neumann(c);                                              // assemble_rhs(copper_surface) with per-face data
    FB_CUDA(c, cudaMemsetAsync(c->d_bcflag.p, 0, c->n_cols * sizeof(int), s));
    FB_CUDA(c, cudaMemsetAsync(c->d_bcval.p, 0, c->n_cols * sizeof(double), s));
    fb::launch_set_bc(c, c->d_bc_dofs.p, (int) c->copper_dofs.size(), dirichlet_value);      // append_dirichlet(copper_bottom, value)
    c->n_dirichlet = c->n_dirichlet_cu;
    fb::launch_bc_prepare(c);
    fb::launch_bc_solution(c);
    c->jds_val_dirty = true; c->cheb_lmax = 0; c->tl_ready = false;
    c->matrix_ok = true; c->assembled = true;
    c->ch_assembled[which] = true; c->ch_assembled[1 - which] = false;
    return sync_check(c, which ? "fb_heat_assemble" : "fb_current_assemble");
}

int fb_current_assemble(fb_ctx* c, const double* face_current_density, int n_faces) {
    FB_REQUIRE(c, c->ch_setup_ok, "fb_current_assemble: call fb_ch_setup first");
    cudaSetDevice(c->device);
    ch_activate(c, 0);
    int rc = ch_face_data(c, face_current_density, n_faces, "fb_current_assemble");
    if (rc) return rc;
    if (!c->ch_matrix_ok[0]) {
        // sigma = 1 (CurrentHeatSolver.cpp:466): the matrix is the plain stiffness matrix of the mesh, assembled once
        FB_CUDA(c, cudaMemsetAsync(c->d_val_save.p, 0, c->nnz * sizeof(double), c->stream));
        fb::launch_assemble_stiffness(c);
        c->ch_matrix_ok[0] = true;
    }
    FB_CUDA(c, cudaMemsetAsync(c->d_rhs.p, 0, c->n_dofs * sizeof(double), c->stream));
    return ch_finish_assembly(c, 0, 0.0);
}

int fb_heat_assemble(fb_ctx* c, double delta_time, const double* face_nottingham, int n_faces) {
    FB_REQUIRE(c, c->ch_setup_ok, "fb_heat_assemble: call fb_ch_setup first");
    FB_REQUIRE(c, c->ch_n_table >= 2, "fb_heat_assemble: no resistivity table (fb_ch_set_physics)");
    FB_REQUIRE(c, delta_time > 0, "fb_heat_assemble: invalid delta time");
    cudaSetDevice(c->device);
    ch_activate(c, 1);
    int rc = ch_face_data(c, face_nottingham, n_faces, "fb_heat_assemble");
    if (rc) return rc;
    const double cu_rho_cp = 3.4496e-24;                                // CurrentHeatSolver.h:118 [J/(K*Ang^3)]
    FB_CUDA(c, cudaMemsetAsync(c->d_val_save.p, 0, c->nnz * sizeof(double), c->stream));
    FB_CUDA(c, cudaMemsetAsync(c->d_rhs.p, 0, c->n_dofs * sizeof(double), c->stream));
    fb::launch_assemble_heat(c, cu_rho_cp * (1.0 / delta_time), c->d_x.p, c->d_x_other.p);
    c->ch_matrix_ok[1] = true;
    return ch_finish_assembly(c, 1, c->ch_T_ambient);
}

int fb_ch_solve(fb_ctx* c, int which, int max_iter, double abs_tol, int precond, int* n_iter, double* final_residual) {
    FB_REQUIRE(c, which == 0 || which == 1, "fb_ch_solve: which must be 0 (current) or 1 (heat)");
    FB_REQUIRE(c, c->mesh_kind == 1 && c->ch_assembled[which] && c->ch_active == which,
               "fb_ch_solve: assemble that system first (the Dirichlet mask and right-hand side belong to the system assembled last)");
    return fb_poisson_solve(c, max_iter, abs_tol, precond, n_iter, final_residual);
}

// x.p of the engine temporarily points at the requested system
struct ChView {
    fb_ctx* c; bool swapped;
    ChView(fb_ctx* c_, int which) : c(c_), swapped(c_->mesh_kind == 1 && which != c_->ch_active) { if (swapped) std::swap(c->d_x.p, c->d_x_other.p); }
    ~ChView() { if (swapped) std::swap(c->d_x.p, c->d_x_other.p); }
};

int fb_ch_export_solution(fb_ctx* c, int which, double* vertex_values) {
    FB_REQUIRE(c, c->mesh_kind == 1 && (which == 0 || which == 1), "fb_ch_export_solution: needs a bulk mesh and which in {0, 1}");
    ChView v(c, which);
    return fb_export_solution(c, vertex_values);
}
int fb_ch_import_solution(fb_ctx* c, int which, const double* vertex_values) {
    FB_REQUIRE(c, c->mesh_kind == 1 && (which == 0 || which == 1), "fb_ch_import_solution: needs a bulk mesh and which in {0, 1}");
    ChView v(c, which);
    return fb_import_solution(c, vertex_values);
}
int fb_ch_export_solution_grad(fb_ctx* c, int which, double* grad3) {
    FB_REQUIRE(c, c->mesh_kind == 1 && (which == 0 || which == 1), "fb_ch_export_solution_grad: needs a bulk mesh and which in {0, 1}");
    ChView v(c, which);
    return fb_export_solution_grad(c, grad3);
}
int fb_ch_check_limits(fb_ctx* c, int which, double lo, double hi, int* bad, double* mn, double* mx) {
    FB_REQUIRE(c, c->mesh_kind == 1 && (which == 0 || which == 1), "fb_ch_check_limits: needs a bulk mesh and which in {0, 1}");
    ChView v(c, which);
    return fb_check_limits(c, lo, hi, bad, mn, mx);
}

// DealSolver::export_surface_centroids (DealSolver.cpp:229-245): centres of the copper_surface faces in cell / face order,
// the order of the per-face data of fb_current_assemble / fb_heat_assemble.  xyz3 == NULL: count only.
int fb_export_surface_centroids(fb_ctx* c, double* xyz3, int* n_faces) {
    FB_REQUIRE(c, c->mesh_ok, "fb_export_surface_centroids: no mesh");
    static const int FV[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
    int n = 0;
    for (const auto& bf : c->bfaces) {
        if (bf.id != 2) continue;
        if (xyz3) {
            double s[3] = {0, 0, 0};
            for (int k = 0; k < 4; ++k) {
                const double* p = &c->xyz[3 * (size_t) c->vert2node[c->dof2vertex[c->cells_dof[8 * (size_t) bf.cell + FV[bf.face][k]]]]];
                s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
            }
            xyz3[3 * (size_t) n] = s[0] / 4.0; xyz3[3 * (size_t) n + 1] = s[1] / 4.0; xyz3[3 * (size_t) n + 2] = s[2] / 4.0;
        }
        ++n;
    }
    if (n_faces) *n_faces = n;
    return FB_OK;
}

// ------------------------------------------------------------------------------------------
int fb_interp_initialize(fb_ctx* c, const int* node_marker, const int* tet4, const int* tet_nbr4, const int* tet_marker, int n_tet,
                         const int* tri3, const int* tri2tet, const double* tri_norm3, int n_tri,
                         const int* quad4, const int* quad2hex, int n_quad, double tet_edgemax,
                         const int* voro_off, const int* voro_list, int n_voro) {
    FB_REQUIRE(c, c->mesh_ok, "fb_interp_initialize: call fb_import_mesh first");
    FB_REQUIRE(c, c->world == 1, "fb_interp_initialize: the interpolator works on replicated (un-partitioned) meshes only");
    FB_REQUIRE(c, node_marker && tet4 && tet_nbr4 && tet_marker && n_tet > 0, "fb_interp_initialize: tetrahedra missing");
    FB_REQUIRE(c, 4L * n_tet == c->n_hex, "fb_interp_initialize: expected 4 hexahedra per tetrahedron");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    c->interp_ok = false;
    const bool verbose = getenv("FB_VERBOSE") != nullptr;
    const double t_begin = omp_get_wtime();
    fb_interp_tables T;
    fb_host_interp_tables(c, node_marker, tet4, tet_nbr4, tet_marker, n_tet, tri3, tri_norm3, n_tri, quad4, n_quad, T);
    c->n_tet = n_tet; c->n_tri = n_tri; c->n_quad = n_quad; c->n_voro = n_voro;
    c->decay_factor = -1.0 / tet_edgemax;            // InterpolatorCells.cpp:534
    FB_CUDA(c, c->d_nxyz.upload(c->xyz, s));
    FB_CUDA(c, c->d_hex8.upload(c->hex8, s));
    FB_CUDA(c, c->d_node2vert.upload(c->node2vert, s));
    FB_CUDA(c, c->d_nodal.alloc(5 * (size_t) c->n_nodes));
    FB_CUDA(c, cudaMemsetAsync(c->d_nodal.p, 0, 5 * (size_t) c->n_nodes * sizeof(double), s));   // empty_value = 0
    FB_CUDA(c, c->d_n2c_off.upload(T.n2c_off, s)); FB_CUDA(c, c->d_n2c_list.upload(T.n2c_list, s));
    FB_CUDA(c, c->d_tet.upload(T.tet, s)); FB_CUDA(c, c->d_tet_cent.upload(T.tet_cent, s)); FB_CUDA(c, c->d_tet_mark.upload(T.tet_mark, s));
    FB_CUDA(c, c->d_tet_nbr_off.upload(T.tet_nbr_off, s)); FB_CUDA(c, c->d_tet_nbr.upload(T.tet_nbr, s));
    FB_CUDA(c, c->d_tet4.upload(tet4, 4 * (size_t) n_tet, s));
    FB_CUDA(c, c->d_hex.upload(T.hex, s));
    FB_CUDA(c, c->d_qtet.upload(T.qtet, s));
    if (n_tri > 0) {
        FB_CUDA(c, c->d_tri.upload(T.tri, s)); FB_CUDA(c, c->d_tri_cent.upload(T.tri_cent, s));
        FB_CUDA(c, c->d_tri_nbr_off.upload(T.tri_nbr_off, s)); FB_CUDA(c, c->d_tri_nbr.upload(T.tri_nbr, s));
        FB_CUDA(c, c->d_tri2tet.upload(tri2tet, 2 * (size_t) n_tri, s));
        FB_CUDA(c, c->d_qtri.upload(T.qtri, s));
    }
    if (n_quad > 0) FB_CUDA(c, c->d_quad2hex.upload(quad2hex, 2 * (size_t) n_quad, s));
    if (n_voro > 0) {
        FB_CUDA(c, c->d_voro_off.upload(voro_off, (size_t) n_voro + 1, s));
        FB_CUDA(c, c->d_voro_list.upload(voro_list, (size_t) std::max(1, voro_off[n_voro]), s));
    }
    int rc = sync_check(c, "fb_interp_initialize");
    if (rc) return rc;
    const double t_upload = omp_get_wtime();
    {   // uniform-grid filter of the tetrahedron scan: box of all mesh nodes on the host, lists on the device
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = 0; i < c->n_nodes; ++i)
            for (int d = 0; d < 3; ++d) { const double x = c->xyz[3 * (size_t) i + d]; lo[d] = std::min(lo[d], x); hi[d] = std::max(hi[d], x); }
        if ((rc = fb::launch_build_cell_grid(c, lo, hi))) return rc;
    }
    if (verbose) fprintf(stderr, "[fb] fb_interp_initialize (ms): host tables + uploads %.2f, cell grid %.2f\n", 1e3 * (t_upload - t_begin), 1e3 * (omp_get_wtime() - t_upload));
    c->interp_ok = true;
    return FB_OK;
}

int fb_extract_solution(fb_ctx* c, int smoothen) {
    FB_REQUIRE(c, c->interp_ok, "fb_extract_solution: interpolator not initialised");
    cudaSetDevice(c->device);
    fb::launch_extract(c, smoothen);
    return sync_check(c, "fb_extract_solution");
}

int fb_get_nodal_solutions(fb_ctx* c, double* sol5) {
    FB_REQUIRE(c, c->interp_ok && sol5, "fb_get_nodal_solutions: interpolator not initialised");
    cudaSetDevice(c->device);
    FB_CUDA(c, cudaMemcpyAsync(sol5, c->d_nodal.p, 5 * (size_t) c->n_nodes * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_get_nodal_solutions");
}

int fb_set_nodal_solutions(fb_ctx* c, const double* sol5) {
    FB_REQUIRE(c, c->interp_ok && sol5, "fb_set_nodal_solutions: interpolator not initialised");
    cudaSetDevice(c->device);
    FB_CUDA(c, cudaMemcpyAsync(c->d_nodal.p, sol5, 5 * (size_t) c->n_nodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return sync_check(c, "fb_set_nodal_solutions");
}

static int check_dim_rank(fb_ctx* c, int dim, int rank) {
    FB_REQUIRE(c, c->interp_ok, "interpolator not initialised");
    FB_REQUIRE(c, dim == 2 || dim == 3, "invalid interpolation dimension");       // SolutionReader.h:61
    FB_REQUIRE(c, rank >= 1 && rank <= 3, "invalid interpolation rank");          // SolutionReader.h:62
    FB_REQUIRE(c, dim == 3 || c->n_tri > 0, "surface interpolation needs triangles");
    return FB_OK;
}

static int reserve_query(fb_ctx* c, long n) {
    FB_CUDA(c, c->d_cellsA.alloc(n)); FB_CUDA(c, c->d_cellsB.alloc(n)); FB_CUDA(c, c->d_scan.alloc(n)); FB_CUDA(c, c->d_scan2.alloc(n));
    FB_CUDA(c, c->d_dirtyA.alloc(n)); FB_CUDA(c, c->d_dirtyB.alloc(n));
    return FB_OK;
}

int fb_locate_interpolate_dev(fb_ctx* c, int dim, int rank, long n, const double* xyz_dev, int* cells_dev, double* sol5_dev) {
    int rc = check_dim_rank(c, dim, rank);
    if (rc) return rc;
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    if ((rc = reserve_query(c, n))) return rc;
    int* base = nullptr;
    if ((rc = fb::launch_locate_chain(c, dim, rank, n, xyz_dev, &base))) return rc;
    fb::launch_finish_interp(c, dim, rank, n, xyz_dev, base, 0, cells_dev, sol5_dev);
    return FB_OK;
}

// host staging: x/y/z with stride -> packed device points
static int stage_points(fb_ctx* c, long n, const double* x, const double* y, const double* z, int stride) {
    FB_REQUIRE(c, x && y && z && stride >= 1, "point arrays missing");
    FB_CUDA(c, c->d_pts.alloc(3 * (size_t) n));
    cudaStream_t s = c->stream;
    if (stride == 3 && y == x + 1 && z == x + 2) {
        FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    } else if (stride == 1) {
        FB_CUDA(c, c->d_sol.alloc(3 * (size_t) n));
        FB_CUDA(c, cudaMemcpyAsync(c->d_sol.p, x, n * sizeof(double), cudaMemcpyHostToDevice, s));
        FB_CUDA(c, cudaMemcpyAsync(c->d_sol.p + n, y, n * sizeof(double), cudaMemcpyHostToDevice, s));
        FB_CUDA(c, cudaMemcpyAsync(c->d_sol.p + 2 * n, z, n * sizeof(double), cudaMemcpyHostToDevice, s));
        fb::launch_pack_points(c, n, c->d_sol.p, c->d_sol.p + n, c->d_sol.p + 2 * n, 1, c->d_pts.p);
        FB_CUDA(c, cudaStreamSynchronize(s));      // d_sol is reused for the results below
    } else {
        std::vector<double> tmp(3 * (size_t) n);
        for (long i = 0; i < n; ++i) { tmp[3 * i] = x[i * stride]; tmp[3 * i + 1] = y[i * stride]; tmp[3 * i + 2] = z[i * stride]; }
        FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, tmp.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
        FB_CUDA(c, cudaStreamSynchronize(s));
    }
    return FB_OK;
}

int fb_locate_interpolate(fb_ctx* c, int dim, int rank, long n, const double* x, const double* y, const double* z, int stride,
                          int* cells_out, double* sol5_out) {
    int rc = check_dim_rank(c, dim, rank);
    if (rc) return rc;
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    if ((rc = stage_points(c, n, x, y, z, stride))) return rc;
    if ((rc = reserve_query(c, n))) return rc;
    FB_CUDA(c, c->d_sol.alloc(5 * (size_t) n));
    int* base = nullptr;
    if ((rc = fb::launch_locate_chain(c, dim, rank, n, c->d_pts.p, &base))) return rc;
    int* final_cells = c->d_cellsA.p;          // the chain buffers are free once the fix-point is in d_scan2
    fb::launch_finish_interp(c, dim, rank, n, c->d_pts.p, base, 0, final_cells, c->d_sol.p);
    if (cells_out) FB_CUDA(c, cudaMemcpyAsync(cells_out, final_cells, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (sol5_out) FB_CUDA(c, cudaMemcpyAsync(sol5_out, c->d_sol.p, 5 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_locate_interpolate");
}

int fb_locate_interpolate_chains(fb_ctx* c, int dim, int rank, long n_chains, long chain_len, const double* xyz,
                                 int* cells_out, double* sol5_out) {
    int rc = check_dim_rank(c, dim, rank);
    if (rc) return rc;
    FB_REQUIRE(c, chain_len >= 1 && n_chains >= 0, "fb_locate_interpolate_chains: invalid chain shape");
    const long n = n_chains * chain_len;
    if (n <= 0) return FB_OK;
    FB_REQUIRE(c, xyz, "fb_locate_interpolate_chains: point array missing");
    cudaSetDevice(c->device);
    if ((rc = stage_points(c, n, xyz, xyz + 1, xyz + 2, 3))) return rc;
    if ((rc = reserve_query(c, n))) return rc;
    FB_CUDA(c, c->d_sol.alloc(5 * (size_t) n));
    int* base = nullptr;
    if ((rc = fb::launch_locate_chain(c, dim, rank, n, c->d_pts.p, &base, chain_len))) return rc;
    int* final_cells = c->d_cellsA.p;
    fb::launch_finish_interp(c, dim, rank, n, c->d_pts.p, base, 0, final_cells, c->d_sol.p);
    if (cells_out) FB_CUDA(c, cudaMemcpyAsync(cells_out, final_cells, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (sol5_out) FB_CUDA(c, cudaMemcpyAsync(sol5_out, c->d_sol.p, 5 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_locate_interpolate_chains");
}

int fb_interpolate(fb_ctx* c, int dim, int rank, long n, const double* x, const double* y, const double* z, int stride,
                   const int* cells, double* sol5_out) {
    int rc = check_dim_rank(c, dim, rank);
    if (rc) return rc;
    if (n <= 0) return FB_OK;
    FB_REQUIRE(c, cells && sol5_out, "fb_interpolate: arrays missing");
    cudaSetDevice(c->device);
    if ((rc = stage_points(c, n, x, y, z, stride))) return rc;
    FB_CUDA(c, c->d_cellsA.alloc(n)); FB_CUDA(c, c->d_sol.alloc(5 * (size_t) n));
    FB_CUDA(c, cudaMemcpyAsync(c->d_cellsA.p, cells, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    fb::launch_finish_interp(c, dim, rank, n, c->d_pts.p, c->d_cellsA.p, 1, nullptr, c->d_sol.p);
    FB_CUDA(c, cudaMemcpyAsync(sol5_out, c->d_sol.p, 5 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_interpolate");
}

int fb_particle_cells_dev(fb_ctx* c, long n, const double* xyz_dev, int* cell_inout_dev) {
    FB_REQUIRE(c, c->interp_ok, "fb_particle_cells: interpolator not initialised");
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    fb::launch_particle_cells(c, n, xyz_dev, cell_inout_dev);
    return FB_OK;
}

int fb_particle_cells(fb_ctx* c, long n, const double* xyz, int* cell_inout) {
    FB_REQUIRE(c, c->interp_ok && xyz && cell_inout, "fb_particle_cells: interpolator not initialised");
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    FB_CUDA(c, c->d_pts.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_cellsA.alloc(n));
    FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(c, cudaMemcpyAsync(c->d_cellsA.p, cell_inout, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    fb::launch_particle_cells(c, n, c->d_pts.p, c->d_cellsA.p);
    FB_CUDA(c, cudaMemcpyAsync(cell_inout, c->d_cellsA.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_particle_cells");
}

int fb_particle_field_dev(fb_ctx* c, long n, const double* xyz_dev, const int* cells_dev, double* E3_dev) {
    FB_REQUIRE(c, c->interp_ok, "fb_particle_field: interpolator not initialised");
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    fb::launch_particle_field(c, n, xyz_dev, cells_dev, E3_dev);
    return FB_OK;
}

int fb_particle_field(fb_ctx* c, long n, const double* xyz, const int* cells, double* E3) {
    FB_REQUIRE(c, c->interp_ok && xyz && cells && E3, "fb_particle_field: interpolator not initialised");
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    FB_CUDA(c, c->d_pts.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_cellsA.alloc(n)); FB_CUDA(c, c->d_sol.alloc(3 * (size_t) n));
    FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(c, cudaMemcpyAsync(c->d_cellsA.p, cells, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    fb::launch_particle_field(c, n, c->d_pts.p, c->d_cellsA.p, c->d_sol.p);
    FB_CUDA(c, cudaMemcpyAsync(E3, c->d_sol.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_check(c, "fb_particle_field");
}

// ---- PIC push: Pic::update_positions / update_velocities with the particles on the device ----
static int pic_positions_impl(fb_ctx* c, long n, double* d_pos, double* d_vel, int* d_cell, double dt, const double* box6, int periodic,
                              long* n_lost) {
    cudaStream_t s = c->stream;
    fb::launch_pic_move(c, n, d_pos, d_vel, d_cell, dt, box6, periodic);
    fb::launch_particle_cells(c, n, d_pos, d_cell, true);
    const size_t nb = (size_t) (n + 1023) / 1024;
    FB_CUDA(c, c->d_pic_pos.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_pic_vel.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_pic_cell.alloc(n));
    FB_CUDA(c, c->d_pic_blk.alloc(nb + 2));
    long* d_total = (long*) c->d_minmax.p;                   // 16 bytes of device scratch
    fb::launch_pic_compact(c, n, d_pos, d_vel, d_cell, c->d_pic_blk.p, d_total, c->d_pic_pos.p, c->d_pic_vel.p, c->d_pic_cell.p);
    long* h = (long*) c->pin_out.p;
    FB_CUDA(c, cudaMemcpyAsync(h, d_total, sizeof(long), cudaMemcpyDeviceToHost, s));
    FB_CUDA(c, cudaStreamSynchronize(s));
    const long kept = *h;
    if (kept < n) {                                          // somebody was lost: the compacted copy replaces the arrays
        FB_CUDA(c, cudaMemcpyAsync(d_pos, c->d_pic_pos.p, 3 * kept * sizeof(double), cudaMemcpyDeviceToDevice, s));
        FB_CUDA(c, cudaMemcpyAsync(d_vel, c->d_pic_vel.p, 3 * kept * sizeof(double), cudaMemcpyDeviceToDevice, s));
        FB_CUDA(c, cudaMemcpyAsync(d_cell, c->d_pic_cell.p, kept * sizeof(int), cudaMemcpyDeviceToDevice, s));
    }
    if (n_lost) *n_lost = n - kept;
    return FB_OK;
}

int fb_pic_update_positions_dev(fb_ctx* c, long n, double* pos3, double* vel3, int* cell, double dt, const double* box6, int periodic,
                                long* n_lost) {
    FB_REQUIRE(c, c->interp_ok, "fb_pic_update_positions: interpolator not initialised");
    FB_REQUIRE(c, box6 && box6[1] > box6[0] && box6[3] > box6[2], "fb_pic_update_positions: invalid simulation box");
    if (n_lost) *n_lost = 0;
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    int rc = pic_positions_impl(c, n, pos3, vel3, cell, dt, box6, periodic, n_lost);
    if (rc) return rc;
    return sync_check(c, "fb_pic_update_positions");
}

int fb_pic_update_positions(fb_ctx* c, long n, double* pos3, double* vel3, int* cell, double dt, const double* box6, int periodic,
                            long* n_lost) {
    FB_REQUIRE(c, c->interp_ok, "fb_pic_update_positions: interpolator not initialised");
    FB_REQUIRE(c, box6 && box6[1] > box6[0] && box6[3] > box6[2], "fb_pic_update_positions: invalid simulation box");
    if (n_lost) *n_lost = 0;
    if (n <= 0) return FB_OK;
    FB_REQUIRE(c, pos3 && vel3 && cell, "fb_pic_update_positions: particle arrays missing");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    FB_CUDA(c, c->d_pts.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_sol.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_cellsA.alloc(n));
    FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, pos3, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    FB_CUDA(c, cudaMemcpyAsync(c->d_sol.p, vel3, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    FB_CUDA(c, cudaMemcpyAsync(c->d_cellsA.p, cell, n * sizeof(int), cudaMemcpyHostToDevice, s));
    long lost = 0;
    int rc = pic_positions_impl(c, n, c->d_pts.p, c->d_sol.p, c->d_cellsA.p, dt, box6, periodic, &lost);
    if (rc) return rc;
    const long kept = n - lost;
    if (kept > 0) {
        FB_CUDA(c, cudaMemcpyAsync(pos3, c->d_pts.p, 3 * kept * sizeof(double), cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaMemcpyAsync(vel3, c->d_sol.p, 3 * kept * sizeof(double), cudaMemcpyDeviceToHost, s));
        FB_CUDA(c, cudaMemcpyAsync(cell, c->d_cellsA.p, kept * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    if (n_lost) *n_lost = lost;
    return sync_check(c, "fb_pic_update_positions");
}

int fb_pic_update_velocities_dev(fb_ctx* c, long n, const double* pos3, const int* cell, double* vel3, double dt, double q_over_m) {
    FB_REQUIRE(c, c->interp_ok, "fb_pic_update_velocities: interpolator not initialised");
    if (n <= 0) return FB_OK;
    cudaSetDevice(c->device);
    fb::launch_pic_velocities(c, n, pos3, cell, vel3, dt * q_over_m);
    return FB_OK;
}

int fb_pic_update_velocities(fb_ctx* c, long n, const double* pos3, const int* cell, double* vel3, double dt, double q_over_m) {
    FB_REQUIRE(c, c->interp_ok, "fb_pic_update_velocities: interpolator not initialised");
    if (n <= 0) return FB_OK;
    FB_REQUIRE(c, pos3 && vel3 && cell, "fb_pic_update_velocities: particle arrays missing");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    FB_CUDA(c, c->d_pts.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_sol.alloc(3 * (size_t) n)); FB_CUDA(c, c->d_cellsA.alloc(n));
    FB_CUDA(c, cudaMemcpyAsync(c->d_pts.p, pos3, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    FB_CUDA(c, cudaMemcpyAsync(c->d_sol.p, vel3, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    FB_CUDA(c, cudaMemcpyAsync(c->d_cellsA.p, cell, n * sizeof(int), cudaMemcpyHostToDevice, s));
    fb::launch_pic_velocities(c, n, c->d_pts.p, c->d_cellsA.p, c->d_sol.p, dt * q_over_m);
    FB_CUDA(c, cudaMemcpyAsync(vel3, c->d_sol.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    return sync_check(c, "fb_pic_update_velocities");
}

}  // extern "C"
