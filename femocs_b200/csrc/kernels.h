// Launch wrappers shared between the translation units of libfemocs_b200 (internal).
#pragma once
#include <utility>

#include "ctx.h"

namespace fb {

// poisson_kernels.cu
int choose_lanes(const fb_ctx* c);
void stream_block_shape(int kernel, int& chunk, int& maxrows);
void launch_assemble_stiffness(fb_ctx* c);
void launch_cell_volumes(fb_ctx* c, double* d_cell_vol);
void launch_charge_density(fb_ctx* c, double* d_scratch);
void launch_neumann(fb_ctx* c);
void launch_assemble_heat(fb_ctx* c, double gamma, const double* d_T_prev, const double* d_phi);
void launch_set_bc(fb_ctx* c, const int* d_dofs, int n, double value);
void launch_bc_prepare(fb_ctx* c);
void launch_bc_solution(fb_ctx* c);
void launch_materialize_eliminated(fb_ctx* c, double* d_val_out, double* d_rhs_out, double* d_lift, double* d_diag_inv, int* d_diagpos_tmp);
void launch_csr_to_jds(fb_ctx* c);
void launch_cg_init(fb_ctx* c, int lanes);
void launch_cg_iteration(fb_ctx* c, int lanes);
void launch_cg_spmv(fb_ctx* c, int lanes);
void launch_cg_vectors(fb_ctx* c);
bool persistent_eligible(fb_ctx* c);
cudaError_t cheb_prepare(fb_ctx* c, int lanes);
// multi-GPU
void launch_pack(fb_ctx* c, const double* v);
void launch_pack_p2p(fb_ctx* c, const double* v);
void launch_flags_to_double(fb_ctx* c, double* out);
void launch_double_to_ghost_flags(fb_ctx* c, const double* in);
void launch_cg_scalars(fb_ctx* c, int which);
void launch_cg_init_spmv(fb_ctx* c, int lanes);
void launch_cg_init_direction(fb_ctx* c);
void launch_cg_update_only(fb_ctx* c);
void launch_cg_direction_only(fb_ctx* c);
cudaError_t launch_cg_persistent(fb_ctx* c);
void launch_minmax(fb_ctx* c);
void launch_solution_grad(fb_ctx* c, double* d_grad3);
void launch_gather(fb_ctx* c, int n, const int* idx, const double* src, double* dst);
void launch_scatter(fb_ctx* c, int n, const int* idx, const double* src, double* dst);

// q2.cu
void launch_q2_stiffness(fb_ctx* c);
void launch_q2_neumann(fb_ctx* c);

// twolevel.cu
int tl_prepare(fb_ctx* c);
void tl_release(fb_ctx* c);
void launch_tl_init_tail(fb_ctx* c);
void launch_tl_vectors(fb_ctx* c);

// interp_kernels.cu
void launch_extract(fb_ctx* c, int smoothen);
int launch_build_cell_grid(fb_ctx* c, const double* bb_lo, const double* bb_hi);
void launch_pack_points(fb_ctx* c, long n, const double* x, const double* y, const double* z, int stride, double* out);
int launch_locate_chain(fb_ctx* c, int dim, int rank, long n, const double* d_pts, int** result, long chain_len = 0);
void launch_finish_interp(fb_ctx* c, int dim, int rank, long n, const double* d_pts, const int* d_base, int final_cells,
                          int* d_cells_out, double* d_sol);
void launch_particle_cells(fb_ctx* c, long n, const double* d_pts, int* d_cells, bool after_move = false);
void launch_particle_field(fb_ctx* c, long n, const double* d_pts, const int* d_cells, double* d_E);
void launch_space_charge(fb_ctx* c, long n, const double* d_pts, const int* d_pcell, double charge_factor);
void launch_pic_move(fb_ctx* c, long n, double* d_pos, const double* d_vel, int* d_cells, double dt, const double* box6, int periodic);
void launch_pic_velocities(fb_ctx* c, long n, const double* d_pos, const int* d_cells, double* d_vel, double dt_q_over_m);
void launch_pic_compact(fb_ctx* c, long n, const double* d_pos, const double* d_vel, const int* d_cells, int* d_block_count,
                        long* d_total, double* pos_out, double* vel_out, int* cell_out);

}  // namespace fb
