// Internal context of libfemocs_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/femocs_b200.h"

namespace fb {

// ---- tiny RAII device buffer ---------------------------------------------------------
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;       // keep capacity
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**) &p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t upload(const std::vector<T>& h, cudaStream_t s) { return upload(h.data(), h.size(), s); }
    cudaError_t zero(cudaStream_t s) { return p ? cudaMemsetAsync(p, 0, n * sizeof(T), s) : cudaSuccess; }
};

// pinned host staging buffer (grows, never shrinks)
struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    cudaError_t reserve(size_t b) {
        if (b <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMallocHost(&p, b);
        if (e == cudaSuccess) bytes = b;
        return e;
    }
};

// cell tables of the interpolator, flattened for the device ----------------------------
struct TetRec { double det0; double d[4][4]; };                 // 136 B: LinearTetrahedra det0..det4
struct TriRec { double vert0[3], edge1[3], edge2[3], pvec[3], norm[3], maxd; };   // 128 B
struct HexRec { double f[8][3]; };                              // 192 B: LinearHexahedra f0..f7

// ---- partitioned CG without NCCL inside the iteration: peer-mapped (CUDA IPC) buffers over NVLink ------------------
// Every rank owns one P2pSlots record in IPC-shared memory; PEERS write into it, the owner polls it.
constexpr int P2P_MAX = 16;
constexpr int TL_SLOT = 8200;           // coarse unknowns (+ 2 scalars) of the two-level preconditioner a rank can exchange
struct P2pSlots {
    double red[2][P2P_MAX][2];          // [reduction of the iteration: 0 = after the SpMV, 1 = after the update][source rank][2 sums]
    long long red_seq[2][P2P_MAX];      // sequence number of the sums above (release / acquire at system scope)
    long long halo_flag[P2P_MAX];       // [source rank]: number of halo exchanges whose values have landed in the ghost segment
    long long tl_seq[P2P_MAX];          // two-level preconditioner: sequence number of the partial restriction below
    double tl[P2P_MAX][TL_SLOT];        // [source rank]: that rank's P^T g (n_c values) followed by its g.Dinv g and g.g
};
struct P2pDesc {                        // device-resident descriptor read by the kernels
    int rank, world;
    P2pSlots* slots[P2P_MAX];           // [rank] = own record, the others are peer mappings
    double* peer_vec[P2P_MAX];          // peers' search-direction vector (owned rows, then the ghost segment)
    int dst_base[P2P_MAX];              // first entry of THIS rank's segment in peer p's vector: n_dofs_p + recv_off_p[rank]
    int send_off[P2P_MAX + 1];          // this rank's send list, per peer
    int n_recv[P2P_MAX];                // ghost values expected from peer p
    long long red_count[2];             // reductions completed (per kind); identical on all ranks
    long long halo_count;               // halo exchanges completed
    long long tl_count;                 // restriction exchanges completed
};

struct CgScalars {          // device-resident CG state (one struct, updated by the kernels)
    double gh;              // g.h of the previous iteration
    double res2;            // ||g||^2 after the last update
    double tol2;            // abs_tol^2
    int it;                 // iterations completed
    int done;               // 1 = converged, 2 = max_iter reached
    int max_iter;
    int pad;
    double* red;            // multi-GPU: kernels deposit their LOCAL sums here (all-reduced over the ranks by NCCL,
                            // then k_cg_scalars_* finishes the step); nullptr on one GPU
    P2pDesc* p2p;           // multi-GPU, peer-mapped mode: the last block all-reduces over NVLink itself (no NCCL, no extra kernel);
                            // done = 3 reports a peer that never answered
};

}  // namespace fb

struct fb_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    long launches = 0;

    // options
    int cg_graph_iters = 32;
    int cheb_degree = 2;
    int dof_order = 0;
    int spmv_split = 32;                     // spmv_kernel 306: rows longer than this are stored as chained segments
    int spmv_occ = 6;                        // CTAs per SM the grid of the 512-row block-JDS kernels is sized for (tuning point)
    int spmv_kernel = -1;                    // -1 auto, 0 row-block stream kernel, 2..32 lanes per row
    int cg_persistent = -1;                  // -1 auto (single cooperative launch when the system fits on chip), 0 off
    int cg_profile = 0;                      // iterations per solve bracketed with CUDA events (0 = off)
    int fe_degree = 1;                       // 1 = FE_Q(1) (the reference build), 2 = FE_Q(2) (q2.cu); read by the next fb_import_mesh

    // ---- multi-GPU partition (one process per GPU; see partition.cpp) ----
    int rank = 0, world = 1;
    void* nccl_comm = nullptr;               // ncclComm_t
    bool host_only = false;                  // plan-only context for CPU tests of the partition logic (no CUDA)
    int n_cols = 0;                          // local columns = owned rows + ghosts (== n_dofs on one GPU)
    int part_n_owned = -1;                   // >= 0: partitioned import, local vertices [0, part_n_owned) are owned
    long n_dofs_global = 0; int n_vert_global = 0, n_cells_global = 0;
    std::vector<int> part_l2g;               // local vertex (= local dof) -> global solver vertex id
    std::vector<int> part_owner;             // owner rank of every local vertex
    std::vector<int> part_cell_g;            // local cell -> global solver cell id
    std::vector<int> send_off, send_idx, recv_off;   // halo plan: per peer, owned dofs to send / ghost segment to receive
    fb::DevBuf<int> d_send_idx, d_l2g, d_gcell2local;
    // peer-mapped iteration (option cg_p2p, default on): IPC mappings of the peers' direction vectors and slot records
    int cg_p2p = 1; bool p2p_ready = false;
    fb::DevBuf<fb::P2pSlots> d_p2p_slots; fb::DevBuf<fb::P2pDesc> d_p2p; fb::DevBuf<unsigned> d_p2p_counter;
    std::vector<void*> p2p_mapped;           // pointers obtained from cudaIpcOpenMemHandle (closed on re-import / destroy)
    fb::DevBuf<double> d_sendbuf, d_red;
    // import intermediates (phase 1 -> phase 2)
    std::vector<int> h_cv, h_v2c_off, h_v2c; std::vector<unsigned char> h_isb;
    double bb_mn[3] = {0, 0, 0}, bb_mx[3] = {0, 0, 0};

    // ---- host copies of the mesh (femocs numbering) ----
    int mesh_reuse = 1; bool mesh_flipped = false, last_import_reused = false; int imported_kind = 0;   // fb_host_try_reuse
    int mesh_kind = 0;                       // 0 = vacuum hexahedra (PoissonSolver), 1 = bulk hexahedra (CurrentHeatSolver)
    int n_nodes = 0, n_hex = 0;
    std::vector<double> xyz;
    std::vector<int> hex8, hex_marker;
    std::vector<int> node2vert, vert2node, hex2cell, cell2hex;
    std::vector<int> vertex2dof, dof2vertex;
    std::vector<int> cells_dof;              // 8 dof ids per solver cell, lexicographic (deal) order
    // fe_degree 2: 27 dofs per cell (local node i + 3 j + 9 k), support points per dof, 9 dofs per top face
    std::vector<int> cells27, topfaces9; std::vector<double> q2_xyz; int imported_degree = 1;
    int n_vert = 0, n_cells = 0, n_dofs = 0;
    long nnz = 0;
    std::vector<int> rowptr, col;            // CSR pattern, columns sorted
    std::vector<int> rowblk;                 // row blocks of the streaming SpMV (first row of each block, n_rowblk + 1 entries)
    int n_rowblk = 0, rowblk_chunk = 0, rowblk_maxrows = 0;
    std::vector<unsigned short> col16;       // windowed SpMV: column position inside the block's window
    std::vector<int> win_off, win_list;      // per row block: sorted distinct columns (CSR over blocks)
    int win_max = 0, win_cap = 0;
    // block-JDS layout of the HBM-roofline SpMV
    int jds_R = 0, jds_nb = 0, jds_maxlen = 0; bool jds_ready = false, jds_val_dirty = true, jds_sym = false, h_needs_zero = false;
    std::vector<unsigned short> jds_perm, jds_len, jds_slot; std::vector<int> jds_jdp, jds_jd, jds_base; int jds_size = 0;
    // rows split into segments (spmv_kernel 306): first row of every block, next segment of a slot's row (0xFFFF: none), cap in use
    std::vector<int> jds_rowbeg; std::vector<unsigned short> jds_link; int jds_split = 0, jds_pad = 2;
    struct BFace { int cell, face, id; };
    std::vector<BFace> bfaces;
    std::vector<int> copper_dofs, top_dofs;  // Dirichlet candidates
    int n_top_faces = 0;
    bool mesh_ok = false, setup_ok = false, assembled = false, matrix_ok = false;
    double applied_field = 0, applied_potential = 0;
    int anode_dirichlet = 0;
    int n_dirichlet = 0;

    // ---- device: solver ----
    fb::DevBuf<double> d_vxyz;               // coordinates per DoF (3*n_dofs)
    fb::DevBuf<int> d_cells;                 // 8*n_cells dof ids (lexicographic)
    fb::DevBuf<int> d_cells27, d_topfaces9;  // fe_degree 2
    fb::DevBuf<int> d_asm_map; bool asm_map_ready = false; int asm_map_opt = 1;   // scatter map of the assembly (64 positions per hexahedron)
    fb::DevBuf<int> d_rowptr, d_col, d_diagpos, d_rowblk, d_win_off, d_win_list;
    fb::DevBuf<unsigned short> d_col16, d_jds_perm, d_jds_len, d_jds_slot;
    fb::DevBuf<int> d_jds_jdp, d_jds_jd, d_jds_base, d_jds_rowbeg; fb::DevBuf<unsigned short> d_jds_link; fb::DevBuf<double> d_val_jds, d_diag;
    // CurrentHeatSolver (mesh_kind 1): the CG engine works on d_val_save / d_x; the system that is NOT active is parked in
    // d_val_other / d_x_other (ch_active: 0 = current, 1 = heat)
    fb::DevBuf<double> d_val_other, d_x_other, d_res_T, d_res_rho;
    int ch_active = 0, ch_n_table = 0; bool ch_setup_ok = false, ch_matrix_ok[2] = {false, false}, ch_assembled[2] = {false, false};
    double ch_T_ambient = 300.0, ch_lorentz = 2.44e-8;
    fb::DevBuf<double> d_face_bc;            // per-face Neumann data (emission current density / Nottingham heat)
    fb::DevBuf<double> d_val_save;           // K before boundary conditions (the reference's system_matrix_save); never rewritten between assemblies
    fb::DevBuf<int> d_bc_dofs;               // Dirichlet candidates: copper dofs, then top dofs (uploaded once per mesh)
    int n_dirichlet_cu = 0, n_dirichlet_cu_top = 0;      // owned constrained rows: copper only / copper + top
    fb::DevBuf<double> d_rhs, d_x, d_g, d_d, d_h, d_dinv, d_z, d_w;
    fb::DevBuf<double> d_rho; int want_charge_density = 0; bool rho_valid = false;   // PoissonSolver::charge_density (filled when the host is about to write files)
    fb::DevBuf<int> d_topfaces;              // 4 dof ids per top (Neumann) face
    fb::DevBuf<int> d_bcflag; fb::DevBuf<double> d_bcval;
    fb::DevBuf<double> d_partial;            // block partials for dot products
    fb::DevBuf<fb::CgScalars> d_cg;
    fb::DevBuf<int> d_vertex2dof;            // n_vert
    fb::DevBuf<int> d_vert_lastcell;         // n_vert: 8 * vertex2cell + vertex2node (DealSolver.cpp:317-341), uploaded on first use
    long n_mesh_faces = -1, n_mesh_edges = -1;   // counted on first use (fb_get_mesh_counts)
    fb::DevBuf<int> d_cell2hex, d_hex2cell;
    fb::DevBuf<double> d_minmax;
    // persistent cooperative CG (native meshes): row slice per CTA
    std::vector<int> pers_cta_row; int pers_grid = 0, pers_cap = 0, pers_rmax = 0; bool pers_uploaded = false;
    fb::DevBuf<int> d_cta_row, d_pers_flags;
    fb::DevBuf<long long> d_dbg; int cg_debug = 0, pers_ctas = 0;
    cudaGraphExec_t cg_graph = nullptr;
    int cg_graph_precond = -1, cg_graph_n = 0;
    double last_solve_ms = 0; int last_iters = 0; long last_spmv = 0; int last_kernel = -1;
    double cheb_lmax = 0, cheb_ratio = 30.0, cheb_inv_theta = 0, cheb_c1[17] = {0}, cheb_c2[17] = {0};
    int cheb_k = 0, cheb_lanes = 0, cheb_power_iters = 15; bool cheb_active = false; double cheb_gershgorin = 0;
    fb::DevBuf<double> d_cheb_p, d_cheb_r;
    // two-level preconditioner (twolevel.cu): Morton aggregates (per mesh), dense inverse of the Galerkin matrix (per matrix)
    bool tl_active = false, tl_ready = false, tl_agg_ready = false; int tl_agg_opt = 0, tl_agg = 0, tl_nc = 0; void* tl_solver = nullptr;
    fb::DevBuf<int> d_tl_perm, d_tl_agg, d_tl_aoff; fb::DevBuf<double> d_tl_inv, d_tl_rc, d_tl_ec, d_tl_rc_part;
    std::vector<double> part_gxyz;           // partitioned import: coordinates of ALL global solver vertices (global Morton aggregates)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> prof_ev;        // 3 events per profiled iteration
    double prof_spmv_ms = 0, prof_vec_ms = 0; int prof_samples = 0;

    // ---- interpolator ----
    bool interp_ok = false;
    int n_tet = 0, n_tri = 0, n_quad = 0, n_voro = 0;
    double decay_factor = -1;
    fb::DevBuf<double> d_nxyz;               // all femocs nodes (3*n_nodes)
    fb::DevBuf<int> d_hex8, d_node2vert;
    fb::DevBuf<double> d_nodal;              // 5*n_nodes
    fb::DevBuf<int> d_n2c_off, d_n2c_list;   // node -> (hex*8 + local) CSR
    fb::DevBuf<int> d_voro_off, d_voro_list;
    fb::DevBuf<fb::TetRec> d_tet; fb::DevBuf<double> d_tet_cent; fb::DevBuf<int> d_tet_mark, d_tet_nbr_off, d_tet_nbr, d_tet4;
    fb::DevBuf<fb::TriRec> d_tri; fb::DevBuf<double> d_tri_cent; fb::DevBuf<int> d_tri_nbr_off, d_tri_nbr, d_tri2tet;
    fb::DevBuf<fb::HexRec> d_hex; fb::DevBuf<int> d_quad2hex;
    fb::DevBuf<int> d_qtet, d_qtri;          // 10 / 6 node ids
    // uniform-grid filter of the tetrahedron scan (option cell_grid, default on)
    int cell_grid = 1; bool grid_on = false; int grid_g[3] = {1, 1, 1}, grid_entries = 0; double grid_lo[3] = {0, 0, 0}, grid_h[3] = {1, 1, 1};
    fb::DevBuf<int> d_grid_off, d_grid_list, d_grid_coff, d_grid_clist, d_grid_cnt;
    // scratch for queries
    fb::DevBuf<double> d_pts; fb::DevBuf<int> d_cellsA, d_cellsB, d_scan, d_scan2, d_flag;
    fb::DevBuf<double> d_pic_pos, d_pic_vel; fb::DevBuf<int> d_pic_cell, d_pic_blk;   // compaction targets of the PIC push
    fb::DevBuf<int> d_needy;                 // [0] = count, [1..] = indices of the points deferred to the block-cooperative scan
    int chain_blocks_per_sm = 0; fb::DevBuf<double> d_sol;
    fb::DevBuf<unsigned char> d_dirtyA, d_dirtyB;
    fb::PinnedBuf pin_in, pin_out;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
};

#define FB_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return (ctx)->fail(FB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// implemented in host_setup.cpp
long fb_host_count_edges(const fb_ctx* c);
void fb_host_vertex_lastcell(const fb_ctx* c, std::vector<int>& out);
bool fb_host_row_blocks(fb_ctx* c, int chunk, int maxrows);
bool fb_host_col_windows(fb_ctx* c, int max_window);
bool fb_host_jds_build(fb_ctx* c, int R, int max_window, bool sym, int split = 0, int pad = 2);
int fb_host_import_mesh(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex);
bool fb_host_try_reuse(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex, int mesh_kind);
int fb_host_import_phase1(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex);
int fb_host_import_phase2(fb_ctx* c);
int fb_host_q2_phase2(fb_ctx* c);         // q2.cu: numbering, sparsity and boundary sets of FE_Q(2)
// partition.cpp: cuts the mesh for c->rank of c->world and runs phase 1 on the local sub-mesh
int fb_host_partition_phase1(fb_ctx* c, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex);
struct fb_interp_tables {
    std::vector<fb::TetRec> tet; std::vector<double> tet_cent; std::vector<int> tet_mark, tet_nbr_off, tet_nbr;
    std::vector<fb::TriRec> tri; std::vector<double> tri_cent; std::vector<int> tri_nbr_off, tri_nbr;
    std::vector<fb::HexRec> hex;
    std::vector<int> qtet, qtri, n2c_off, n2c_list;
};
void fb_host_interp_tables(const fb_ctx* c, const int* node_marker, const int* tet4, const int* tet_nbr4,
                           const int* tet_marker, int n_tet, const int* tri3, const double* tri_norm3, int n_tri,
                           const int* quad4, int n_quad, fb_interp_tables& out);
