/*
 * femocs_b200.h -- C ABI of libfemocs_b200.so, the B200 (sm_100a) implementation of the
 * per-step electrostatic hot path of FEMOCS (veskem/femocs).
 *
 * This is the drop-in boundary: the reference's C++ classes femocs::PoissonSolver<3> /
 * DealSolver<3>, Interpolator::extract_solution, the SolutionReader interpolation loops and
 * the Pic<3> locate / field look-ups are re-implemented as thin forwards to the entry
 * points below (see INTEGRATION.md and femocs_b200/cxx/ for the seam).  Every entry point
 * cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: opaque handle, plain pointers and sizes; no C++/torch types cross this line;
 *   - every function returning int returns 0 on success and a non-zero FB_ERR_* code on
 *     failure (fb_last_error gives the text); no exception ever crosses this boundary,
 *     matching the reference's "return code, never throw across the ABI" protocol
 *     (include/Femocs_wrap.h:17, src/DealSolver.cpp:193-205,446-457);
 *   - host pointers unless the name ends in _dev; device work is ordered on the context's
 *     stream and every non-_dev call is synchronous at return;
 *   - all floating point data is IEEE binary64, all indices are 32-bit int (deal.II's
 *     types::global_dof_index is 32-bit in the reference build);
 *   - mesh arrays are exactly what the reference's TetgenMesh hands out (SURVEY.md 8a'):
 *     hexahedron h = 4*tet + k belongs to tetrahedron `tet` (src/Tethex.cpp:1041-1096),
 *     quadrangle q = 3*tri + k to triangle `tri`; hex marker > 0 = vacuum
 *     (src/TetgenMesh.cpp:353-375); tet marker 3 = TYPES.VACUUM.
 */
#ifndef FEMOCS_B200_H_
#define FEMOCS_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_ctx fb_ctx;

enum {
    FB_OK = 0,
    FB_ERR_CUDA = 1,       /* a CUDA runtime call or kernel failed                       */
    FB_ERR_ARG = 2,        /* invalid argument / call order                              */
    FB_ERR_MESH = 3,       /* mesh could not be imported (import_mesh would return false) */
    FB_ERR_NO_DEVICE = 4   /* no usable CUDA device: there is NO CPU fallback            */
};

/* preconditioners of fb_poisson_solve */
enum {
    FB_PRECOND_JACOBI = 1,     /* diagonal scaling                                        */
    FB_PRECOND_CHEBYSHEV = 2,  /* Chebyshev polynomial of the Jacobi-scaled operator      */
    FB_PRECOND_TWOLEVEL = 3    /* Jacobi + aggregation coarse-grid correction (one GPU, multi-kernel path): same solution, ~3-5x
                                  fewer iterations on HBM-sized systems; option "tl_agg" = dofs per aggregate (0 = auto) */
};

/* ---------------------------------------------------------------------------------------
 * life cycle -- replaces the construction of PoissonSolver<3>/Interpolator members in
 * ProjectRunaway (src/ProjectRunaway.cpp:19-53).  `device` is the CUDA ordinal.
 * Returns NULL when no CUDA device can be initialised (fb_create_error() has the reason).
 * ------------------------------------------------------------------------------------- */
fb_ctx*     fb_create(int device);
void        fb_destroy(fb_ctx* ctx);
const char* fb_last_error(const fb_ctx* ctx);
const char* fb_create_error(void);
/* number of kernels launched by this context since creation (bench.py "gpu_launches") */
long        fb_kernel_launches(const fb_ctx* ctx);
/* tunables: "cg_graph_iters" (iterations per CUDA-graph launch), "cheb_degree" (polynomial degree k >= 2 of
 * FB_PRECOND_CHEBYSHEV, default 2 -> k SpMVs per CG iteration), "cheb_eig_ratio" (lmax / lmin of the interval, default 30), "cheb_power_iters" (power iterations that sharpen the
 * Gershgorin bound of lmax, default 15, 0 = Gershgorin only),
 * "dof_order" (0 = deal.II first-touch numbering, 1 = Morton order of vertex coordinates),
 * "cg_profile" (see fb_last_solve_profile),
 * "fe_degree" (element of the NEXT fb_import_mesh: 1 = FE_Q(1), the reference as shipped -- include/DealSolver.h:130
 * shape_degree = 1; 2 = FE_Q(2) with QGauss(3), what the same call sites do when that constant reads 2;
 * un-partitioned field solver only; fb_export_solution still returns the vertex values, src/DealSolver.cpp:317-341) */
int         fb_set_option(fb_ctx* ctx, const char* key, double value);

/* ---------------------------------------------------------------------------------------
 * multi-GPU (no reference counterpart: DealSolver is serial, include/DealSolver.h:135-143).  One process
 * per GPU.  Rank 0 draws a 128-byte communicator id, the host code hands it to every rank by any channel
 * (torch.distributed / MPI broadcast), and each rank calls fb_comm_init BEFORE fb_import_mesh.  With
 * world > 1 fb_import_mesh takes the SAME full mesh on every rank and keeps the rank's element partition
 * (recursive coordinate bisection of the vertices; rows of the owned vertices, ghost columns behind them);
 * assemble / solve / check_limits / export_solution then work on the distributed system: the CG exchanges
 * the halo of its search direction with NCCL point-to-point over NVLink and all-reduces its dot products.
 * fb_export_solution returns the complete potential (global solver-vertex order) on every rank.
 * The interpolator entry points need an un-partitioned context (native meshes are "replicas only").
 * ------------------------------------------------------------------------------------- */
int fb_comm_unique_id(char* id128);
int fb_comm_init(fb_ctx* ctx, int rank, int world, const char* id128);
/* out[0..9] = rank, world, owned rows, local columns (rows + ghosts), local nnz, local cells,
 * halo values sent per exchange, ghosts received, global solver vertices, global solver cells */
int fb_get_partition(const fb_ctx* ctx, long* out10);
/* how the partitioned CG communicates: 0 = one GPU; 1 = NCCL inside the iteration (grouped ncclSend/ncclRecv halo + two
 * ncclAllReduce per iteration, host-issued); 2 = peer-mapped (default when CUDA IPC works between the ranks, option
 * "cg_p2p" = 0 disables): halo values and the two pairs of sums are stored straight into the peers' memory over NVLink by
 * the kernels themselves, so that an iteration is 4 kernels, captured in a CUDA graph like on one GPU */
int fb_comm_mode(const fb_ctx* ctx);
/* host-only view of the same partition logic (no CUDA): used by the world_size-2 CPU tests.
 * phase1 returns the extremes of the local boundary-face centres (min xyz, max xyz); the caller reduces
 * them over the ranks (min / max) and passes the result to phase2.  Destroy with fb_destroy. */
fb_ctx* fb_plan_create(int rank, int world);
int fb_plan_phase1(fb_ctx* plan, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex, double* bbox6);
int fb_plan_phase2(fb_ctx* plan, const double* bbox6_global);
/* host-only, un-partitioned: everything fb_import_mesh does on the host (vertex compaction, orientation, boundary ids,
 * first-touch DoF numbering, sparsity) on a plan context; fb_plan_sizes / fb_plan_get then describe the complete system */
int fb_plan_import(fb_ctx* plan, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex);
int fb_plan_sizes(const fb_ctx* plan, long* out8);
int fb_plan_get(const fb_ctx* plan, int* local2global, int* owner, int* send_off, int* send_idx, int* recv_off, int* rowptr, int* col,
                int* cells_dof, int* local_cell2global, int* copper_flag, int* top_flag);
/* host-only: the tables fb_interp_initialize uploads (src/InterpolatorCells.cpp:523-629, 1205-1267, 1585-1637,
 * 1151-1173, 1873-1895), computed for the complete mesh on a fresh plan context; CPU parity tests compare them bit
 * for bit.  tet17 = {det0, d[4][4]}, tri16 = {vert0, edge1, edge2, pvec, norm, maxd}, hex24 = f0..f7. */
int fb_plan_interp_tables(fb_ctx* plan, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex,
                          const int* node_marker, const int* tet4, const int* tet_nbr4, const int* tet_marker, int n_tet,
                          const int* tri3, const double* tri_norm3, int n_tri, const int* quad4, int n_quad,
                          double* tet17, double* tet_cent3, int* tet_mark, double* hex24, double* tri16, double* tri_cent3,
                          int* qtet10, int* qtri6);
/* host-only: block-JDS tables (R rows per block; sym != 0: lower triangle only) of the plan's sparsity.
 * sizes6 = {blocks, stored slots incl. padding, window entries, longest row, largest window, diagonal offsets} */
int fb_plan_jds(fb_ctx* plan, int R, int max_window, int sym, long* sizes6);
/* the layout of spmv_kernel 306: rows longer than `split` entries stored as chained segments; rowbeg[nb + 1] = first row of
 * every block, link[nb * R] = slot of the next segment of the same row (0xFFFF: none); perm gains bit 15 = "not the head" */
int fb_plan_jds_split(fb_ctx* plan, int R, int max_window, int split, long* sizes6);
int fb_plan_jds_get_split(const fb_ctx* plan, int* rowbeg, unsigned short* link);
/* host-only: mesh kind of the next fb_plan_import: 0 = vacuum hexahedra (fb_import_mesh), 1 = bulk (fb_import_bulk_mesh);
 * fb_export_surface_centroids works on plan contexts too */
int fb_plan_set_kind(fb_ctx* plan, int kind);
int fb_plan_jds_get(const fb_ctx* plan, unsigned short* perm, unsigned short* len, unsigned short* slot, int* jdp, int* jd,
                    int* base, unsigned short* col16, int* win_off, int* win_list);

/* ---------------------------------------------------------------------------------------
 * bool DealSolver<3>::import_mesh(vector<Point<3>> vertices, vector<CellData<3>> cells)
 *   src/DealSolver.cpp:191-209 (+ mark_boundary :460-518, PoissonSolver::mark_mesh
 *   src/PoissonSolver.cpp:52-55), fed by TetgenNodes::export_dealii and
 *   Hexahedra::export_vacuum (src/TetgenCells.cpp:179-184,673-686) at
 *   src/ProjectRunaway.cpp:216.
 * Takes the FULL femocs node and hexahedron lists; hexahedra with hex_marker > 0 form the
 * solver mesh (solver cell id = rank among them, solver vertex id = rank among the nodes
 * they touch: src/InterpolatorCells.cpp:38-66,1247-1266).
 * ------------------------------------------------------------------------------------- */
int fb_import_mesh(fb_ctx* ctx, const double* xyz, int n_nodes,
                   const int* hex8, const int* hex_marker, int n_hex);

/* Mesh hand-off with unchanged topology (re-mesh decision of src/ProjectRunaway.cpp:55-67,
 * src/GeneralProject.cpp:19-55): when fb_import_mesh / fb_import_bulk_mesh receives the node count, hexahedra and markers
 * of the mesh it already holds, only the geometry is refreshed (orientation and boundary ids re-validated); numbering,
 * sparsity, SpMV tables and the captured CG graph are kept.  Returns 1 if the last import took that path.  Option
 * "mesh_reuse" = 0 forces the full import.  Un-partitioned contexts only. */
int fb_last_import_reused(const fb_ctx* ctx);

/* sizes after import: out[0]=n_dofs, [1]=n_cells, [2]=nnz, [3]=n_vertices,
 * [4]=n_boundary_faces, [5]=n_top_faces, [6]=n_dirichlet_dofs (after assemble) */
int fb_get_sizes(const fb_ctx* ctx, long* out7);
/* fe_degree 2: the 27 dofs of every cell (DoFCellAccessor::get_dof_indices re-ordered to the tensor lattice: local node
 * (i, j, k) in {0, 1, 2}^3 at index i + 3 j + 9 k); FB_ERR_ARG for an FE_Q(1) mesh */
int fb_get_cells27(const fb_ctx* ctx, int* cells27);

/* void PoissonSolver<3>::setup(double field, double potential)  src/PoissonSolver.cpp:162-167
 * (+ DealSolver::setup_system src/DealSolver.cpp:368-387: zero matrix/rhs, solution = 0).
 * anode_is_dirichlet mirrors Config::Field::anode_BC == "dirichlet" (src/Config.cpp:68). */
int fb_poisson_setup(fb_ctx* ctx, double field, double potential, int anode_is_dirichlet);

/* void PoissonSolver<3>::assemble(bool first_time)              src/PoissonSolver.cpp:170-210
 *   first_time: Q1 stiffness assembly (assemble_parallel :213-263); otherwise the saved
 *   pre-BC matrix is restored (:157-159).  Then copper Dirichlet (0 V), Neumann top faces
 *   (DealSolver::assemble_rhs src/DealSolver.cpp:389-430) or Dirichlet anode, the
 *   space-charge RHS of assemble_space_charge_fast (:299-319) when particles are given,
 *   and apply_dirichlet (src/DealSolver.cpp:437-440).
 *   particle_xyz: 3*n doubles, particle_cell: solver cell index per particle (as
 *   SuperParticle::cell), charge_factor = q_over_eps0 * Wsp.  n_particles = 0 -> Laplace. */
int fb_poisson_assemble(fb_ctx* ctx, int first_time,
                        const double* particle_xyz, const int* particle_cell,
                        long n_particles, double charge_factor);

/* int PoissonSolver<3>::solve() -> DealSolver::solve_cg(n_cg, cg_tolerance, ssor_param)
 *   src/PoissonSolver.h:54, src/DealSolver.cpp:442-458.
 * Preconditioned CG on K phi = b, warm-started from the current solution, stopping on the
 * ABSOLUTE l2 residual <= abs_tol like deal.II's SolverControl.  Returns through *n_iter the
 * reference's convention: +iterations when converged, -iterations when max_iter was hit. */
int fb_poisson_solve(fb_ctx* ctx, int max_iter, double abs_tol, int precond,
                     int* n_iter, double* final_residual);

/* void DealSolver::export_solution(vector<double>&)             src/DealSolver.cpp:269-278
 * void PoissonSolver::export_charge_dens(vector<double>&)       src/PoissonSolver.cpp:141-149
 * both in solver-vertex order, n_vertices values. */
int fb_export_solution(fb_ctx* ctx, double* phi_vertex);
int fb_export_charge_dens(fb_ctx* ctx, double* rho_vertex);   /* zeros unless option "charge_density" = 1 was set before the
                                                                   last assemble (the reference's write_time(), src/PoissonSolver.cpp:196-207):
                                                                   then rhs / dof_volume (DealSolver::calc_dof_volumes :344-366) */
/* void DealSolver::export_solution_grad(vector<Tensor<1,3>>&)    src/DealSolver.cpp:280-301
 *   grad3[3 v .. 3 v + 2] = MINUS the gradient of the solution, taken -- exactly as the reference does -- at Gauss point
 *   number vertex2node[v] of cell vertex2cell[v] (the LAST cell, in cell order, that holds vertex v; :317-341), not at
 *   the vertex itself.  n_vertices triples. */
int fb_export_solution_grad(fb_ctx* ctx, double* grad3_vertex);
/* void DealSolver::import_solution(const vector<double>*)       src/DealSolver.cpp:303-315 */
int fb_import_solution(fb_ctx* ctx, const double* phi_vertex);

/* bool DealSolver::check_limits(lo, hi) + stat.sol_min/sol_max  src/DealSolver.cpp:157-167 */
int fb_check_limits(fb_ctx* ctx, double lo, double hi, int* out_of_limits,
                    double* sol_min, double* sol_max);

/* double DealSolver::get_cell_vol(i) / int get_n_cells()        src/DealSolver.cpp:169-173 */
int fb_get_cell_volumes(fb_ctx* ctx, double* vol_cells);

/* ---------------------------------------------------------------------------------------
 * CurrentHeatSolver<3> on the BULK hexahedra: the current-continuity and heat equations of
 * ProjectRunaway::solve_heat (src/ProjectRunaway.cpp:535-571), through the same assembly + PCG engine.
 * `which`: 0 = CurrentSolver (current potential), 1 = HeatSolver (temperature).  One context per mesh kind: a
 * context that imported a bulk mesh serves fb_ch_* / fb_current_* / fb_heat_* only.
 * ------------------------------------------------------------------------------------- */
/* bool CurrentHeatSolver::import_mesh(nodes.export_dealii(), hexs.export_bulk())   src/ProjectRunaway.cpp:222,
 *   Hexahedra::export_bulk (hex_marker < 0) src/TetgenCells.cpp:688-701, mark_mesh src/CurrentHeatSolver.cpp:526-530:
 *   top + other -> copper_surface (Neumann faces with per-face data), bottom -> copper_bottom (Dirichlet). */
int fb_import_bulk_mesh(fb_ctx* ctx, const double* xyz, int n_nodes, const int* hex8, const int* hex_marker, int n_hex);
/* PhysicalQuantities::resistivity_data (src/PhysicalQuantityData.cpp:47-, rows {T [K], rho}) and Config::Heating::lorentz:
 *   sigma(T) = 1 / (10 rho(T)), kappa(T) = lorentz T sigma(T), T clamped to the table (src/PhysicalQuantities.cpp:31-64) */
int fb_ch_set_physics(fb_ctx* ctx, const double* table_T, const double* table_rho, int n_rows, double lorentz);
/* void CurrentHeatSolver::setup(double temperature)             src/CurrentHeatSolver.cpp:509-513
 *   both systems zeroed, current potential = 0, temperature = T_ambient everywhere */
int fb_ch_setup(fb_ctx* ctx, double T_ambient);
/* void DealSolver::export_surface_centroids(Medium&)            src/DealSolver.cpp:229-245
 *   centres of the copper_surface faces in cell / face order = the order of the per-face data below.
 *   xyz3 may be NULL (count only). */
int fb_export_surface_centroids(fb_ctx* ctx, double* xyz3, int* n_faces);
/* void CurrentSolver::assemble()                                src/CurrentHeatSolver.cpp:420-484
 *   stiffness matrix (sigma = 1), Neumann faces with the emission current densities
 *   (EmissionSolver::get_face_bc, include/CurrentHeatSolver.h:53-56), phi = 0 on copper_bottom */
int fb_current_assemble(fb_ctx* ctx, const double* face_current_density, int n_faces);
/* void HeatSolver::assemble(double delta_time [s])              src/CurrentHeatSolver.cpp:105-152,346-393
 *   implicit Euler: (cu_rho_cp / dt) M + K_kappa(T_prev); load = (cu_rho_cp / dt) T_prev + sigma(T_prev) |grad phi|^2 with
 *   phi the current potential solved last; Neumann faces with the Nottingham heat; T = T_ambient on copper_bottom */
int fb_heat_assemble(fb_ctx* ctx, double delta_time, const double* face_nottingham_heat, int n_faces);
/* int EmissionSolver::solve()  include/CurrentHeatSolver.h:41 -> DealSolver::solve_cg(conf->n_cg, cg_tolerance, ssor_param);
 *   the system assembled last; same conventions as fb_poisson_solve (warm start, absolute tolerance, +#CG / -#CG) */
int fb_ch_solve(fb_ctx* ctx, int which, int max_iter, double abs_tol, int precond, int* n_iter, double* final_residual);
/* heat.export_solution / current.export_solution / current.export_solution_grad (CurrentHeatSolver::export_temp_rho
 * src/CurrentHeatSolver.cpp:515-523, Interpolator::extract_solution(CurrentHeatSolver&) src/Interpolator.cpp:210-219),
 * DealSolver::import_solution (heat_transfer.interpolate_dofs hand-over, src/ProjectRunaway.cpp:282) and
 * heat.check_limits (src/ProjectRunaway.cpp:553) */
int fb_ch_export_solution(fb_ctx* ctx, int which, double* vertex_values);
int fb_ch_import_solution(fb_ctx* ctx, int which, const double* vertex_values);
int fb_ch_export_solution_grad(fb_ctx* ctx, int which, double* grad3_vertex);
int fb_ch_check_limits(fb_ctx* ctx, int which, double lo, double hi, int* out_of_limits, double* sol_min, double* sol_max);

/* operator<<(ostream&, const DealSolver&) / to_str()             include/DealSolver.h:107-117
 *   #faces and #edges of the solver mesh (tria->n_active_faces(), n_active_lines()); counted on the host on first use */
int fb_get_mesh_counts(fb_ctx* ctx, long* n_faces, long* n_edges);
/* the solver mesh as deal.II holds it (for write("*.vtk|msh"), src/DealSolver.cpp:351-366): coordinates of the
 * n_vertices solver vertices and the 8 vertex ids of every solver cell in old-style (UCD) order */
int fb_get_solver_mesh(fb_ctx* ctx, double* xyz_vertex, int* cells_ucd8);

/* test / integration hooks: the assembled system in DoF numbering (host copies).
 * Any pointer may be NULL. */
int fb_get_system(fb_ctx* ctx, int* rowptr, int* col, double* val, double* val_save,
                  double* rhs, double* sol, int* vertex2dof, int* vertex2node);

/* ---------------------------------------------------------------------------------------
 * void Interpolator::initialize(const TetgenMesh*, double empty_val, int search_region)
 *   src/Interpolator.cpp:28-77 with search_region = TYPES.VACUUM (src/ProjectRunaway.cpp:435)
 *   and the precompute() of LinearTetrahedra / LinearHexahedra / LinearTriangles /
 *   LinearQuadrangles / Quadratic* (src/InterpolatorCells.cpp:523-629,1205-1267,1585-1637,
 *   1151-1173,1873-1895).  Must follow fb_import_mesh (uses its nodes and hexahedra).
 *   voro_off/voro_list: CSR of TetgenMesh::calc_pseudo_3D_vorocells(vacuum=true)
 *   (src/TetgenMesh.cpp:819-837), n_voro rows; tet_edgemax = tets.stat.edgemax.
 * ------------------------------------------------------------------------------------- */
int fb_interp_initialize(fb_ctx* ctx, const int* node_marker,
                         const int* tet4, const int* tet_nbr4, const int* tet_marker, int n_tet,
                         const int* tri3, const int* tri2tet, const double* tri_norm3, int n_tri,
                         const int* quad4, const int* quad2hex, int n_quad,
                         double tet_edgemax,
                         const int* voro_off, const int* voro_list, int n_voro);

/* void Interpolator::extract_solution(PoissonSolver<3>&, bool smoothen)
 *   src/Interpolator.cpp:172-190 (store_solution :103-123, store_elfield :125-140,
 *   average_nodal_fields :142-170).  The nodal Solution array stays on the device. */
int fb_extract_solution(fb_ctx* ctx, int smoothen);

/* InterpolatorNodes::solutions access (vector<Solution>, 5 doubles per femocs node:
 * Ex, Ey, Ez, scalar1 (charge density), scalar2 (potential); include/Primitives.h:507-522) */
int fb_get_nodal_solutions(fb_ctx* ctx, double* sol5);
int fb_set_nodal_solutions(fb_ctx* ctx, const double* sol5);

/* void SolutionReader::calc_full_interpolation()                src/SolutionReader.cpp:136-165
 *   the chained-guess loop  cell = locate_interpolate(i, cell)  (:43-65,
 *   src/InterpolatorCells.cpp:442-445,269-307) for dim in {2,3}, rank in {1,2,3}:
 *     dim 2: lintri / quadtri / linquad, dim 3: lintet / quadtet / linhex.
 *   Points are x[i*stride], y[i*stride], z[i*stride] (stride 1 = the three arrays of
 *   femocs_interpolate_elfield, include/Femocs_wrap.h:36; stride 3 with y=x+1, z=x+2 = the
 *   Atom/Point3 layout).  cells_out[i] is what the reference stores in atom.marker
 *   (negative = outside, nearest cell), sol5_out the interpolated Solution.
 *   Cell indices are bit-exact with the reference's sequential loop. */
int fb_locate_interpolate(fb_ctx* ctx, int dim, int rank, long n,
                          const double* x, const double* y, const double* z, int stride,
                          int* cells_out, double* sol5_out);

/* void SolutionReader::calc_interpolation() with atoms already mapped to cells
 *   src/SolutionReader.cpp:167-190, :91-112 (interp_solution(i) with cell = abs(marker)) */
int fb_interpolate(fb_ctx* ctx, int dim, int rank, long n,
                   const double* x, const double* y, const double* z, int stride,
                   const int* cells, double* sol5_out);

/* void EmissionReader::emission_line(point, direction, rmax)     src/EmissionReader.cpp:51-61, batched:
 *   n_chains independent point sequences of chain_len points each (xyz: n_chains * chain_len packed triples);
 *   every sequence is located like a fresh SolutionReader (phis_on_line.reserve + calc_interpolation: the guess
 *   chain restarts from the first guess at the first point of each line), all sequences in one call. */
int fb_locate_interpolate_chains(fb_ctx* ctx, int dim, int rank, long n_chains, long chain_len,
                                 const double* xyz, int* cells_out, double* sol5_out);

/* int Pic<3>::update_point_cell(const SuperParticle&)           src/Pic.cpp:186-196
 *   cell_inout: solver cell index guess in, located solver cell (or -1) out. */
int fb_particle_cells(fb_ctx* ctx, long n, const double* xyz, int* cell_inout);

/* field look-up of Pic<3>::update_velocities                    src/Pic.cpp:198-209, :38-39
 *   E3[i] = linhex.interp_gradient(pos_i, deal2femocs(cell_i))
 *   (src/InterpolatorCells.cpp:410-423,1363-1416) */
int fb_particle_field(fb_ctx* ctx, long n, const double* xyz, const int* cells, double* E3);

/* int Pic<3>::update_positions()                                 src/Pic.cpp:137-184
 *   + ParticleSpecies::clear_lost()                              src/ParticleSpecies.cpp:16-31
 *   pos += vel dt; periodic images in x, y (src/Macros.cpp:41-48) or loss outside the x/y box; loss at
 *   z >= zmax; cell search from the previous cell (update_point_cell); particles whose cell is -1 are removed
 *   and the rest moved forward keeping their order.  pos3 / vel3 / cell are updated IN PLACE (first n - n_lost
 *   entries valid afterwards).  box6 = {xmin, xmax, ymin, ymax, zmin, zmax} (Medium::Sizes of Pic::set_params). */
int fb_pic_update_positions(fb_ctx* ctx, long n, double* pos3, double* vel3, int* cell, double dt,
                            const double* box6, int periodic, long* n_lost);

/* void Pic<3>::update_velocities()                               src/Pic.cpp:198-209
 *   vel_i += linhex.interp_gradient(pos_i, deal2femocs(cell_i)) * (dt * q_over_m) */
int fb_pic_update_velocities(fb_ctx* ctx, long n, const double* pos3, const int* cell, double* vel3,
                             double dt, double q_over_m);

/* ---------------------------------------------------------------------------------------
 * device-resident variants (inputs/outputs already in HBM; asynchronous on the context
 * stream until fb_synchronize).  Used by the benchmark's "value" leg and by callers that
 * keep atoms/particles on the GPU.
 * ------------------------------------------------------------------------------------- */
int fb_locate_interpolate_dev(fb_ctx* ctx, int dim, int rank, long n,
                              const double* xyz_dev, int* cells_dev, double* sol5_dev);
int fb_particle_cells_dev(fb_ctx* ctx, long n, const double* xyz_dev, int* cell_inout_dev);
int fb_particle_field_dev(fb_ctx* ctx, long n, const double* xyz_dev, const int* cells_dev,
                          double* E3_dev);
int fb_poisson_assemble_dev(fb_ctx* ctx, int first_time, const double* particle_xyz_dev,
                            const int* particle_cell_dev, long n_particles, double charge_factor);
/* particles resident in HBM across PIC steps; n_lost is written once the stream has been drained (the call
 * synchronises, as the caller needs the new particle count) */
int fb_pic_update_positions_dev(fb_ctx* ctx, long n, double* pos3_dev, double* vel3_dev, int* cell_dev, double dt,
                                const double* box6, int periodic, long* n_lost);
int fb_pic_update_velocities_dev(fb_ctx* ctx, long n, const double* pos3_dev, const int* cell_dev, double* vel3_dev,
                                 double dt, double q_over_m);
int fb_synchronize(fb_ctx* ctx);

/* timing hooks for the roofline report: device time (ms, CUDA events on the context's
 * stream) of the last fb_poisson_solve and its iteration count */
int fb_last_solve_stats(const fb_ctx* ctx, double* solve_ms, int* iterations, long* spmv_launches);
/* with option "cg_profile" = k > 0 the first k iterations of every solve are bracketed by CUDA
 * events on the context's stream: average device time (ms) of the SpMV+dot kernel and of the
 * vector-update kernels over the sampled (live) iterations */
int fb_last_solve_profile(const fb_ctx* ctx, double* spmv_ms_avg, double* vector_ms_avg, int* n_samples);
/* which SpMV kernel the last fb_poisson_solve ran (option "spmv_kernel" codes: -2 = single cooperative launch
 * k_cg_persistent, 2..32 = CSR lanes per row, 100.. = row-block streaming, 200.. = windowed streaming, 300..305 =
 * block-JDS (304: matrix stream read evict-first; 305: same with the load order pinned), 306 = segmented block-JDS, the
 * default for HBM-sized systems (rows longer than option "spmv_split" = 32 entries stored as chained segments), 307 = 306 with
 * the next block's index list prefetched, 308 = 306 with the matrix stream through a TMA / mbarrier ring of shared-memory
 * stages (both measured slower, kept as tested variants), 310/311 = symmetric block-JDS storing the lower triangle only;
 * option "spmv_occ" = CTAs per SM the grid of the 512-slot kernels is sized for) */
int fb_last_solve_kernel(const fb_ctx* ctx);
/* the context's cudaStream_t (for callers that record their own events around fb_*_dev calls) */
void* fb_get_stream(fb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FEMOCS_B200_H_ */
