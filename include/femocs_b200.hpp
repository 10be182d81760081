// femocs_b200.hpp -- C++ seam above the C ABI (include/femocs_b200.h).
//
// Header-only classes with the member names, argument meaning and error behaviour of the reference
// classes they stand in for, so that the call sites in src/ProjectRunaway.cpp keep their shape:
//
//   femocs::PoissonSolver<3>  (include/PoissonSolver.h:25-60, include/DealSolver.h:40-127)  -> femocs_b200::PoissonSolver
//   femocs::Interpolator      (include/Interpolator.h:24-68)                                -> femocs_b200::Interpolator
//   femocs::FieldReader / SolutionReader (include/SolutionReader.h:27-140, :143-230)        -> femocs_b200::FieldReader
//   femocs::Pic<3> locate / field look-ups (include/Pic.h, src/Pic.cpp:186-209)             -> femocs_b200::Pic
//
// Conventions kept from the reference: bool import_mesh (false = failed), int solve() = +#CG / -#CG,
// check_limits() true = OUT of limits (src/DealSolver.cpp:157-167), export_results label case decides
// append (exact case) vs overwrite (upper case) (src/SolutionReader.cpp:269-326), no exception is thrown
// on a numerical failure -- only on API misuse / CUDA errors (std::runtime_error with fb_last_error text),
// which the C wrappers of Femocs_wrap.cpp never let cross the ABI.
// There is no CPU fallback: constructing a Context without a CUDA device throws.
#ifndef FEMOCS_B200_HPP_
#define FEMOCS_B200_HPP_

#include <algorithm>
#include <cctype>
#include <cmath>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "femocs_b200.h"

namespace femocs_b200 {

class Context {
public:
    explicit Context(int device = 0) : h(fb_create(device)) {
        if (!h) throw std::runtime_error(std::string("fb_create failed: ") + fb_create_error());
    }
    ~Context() { fb_destroy(h); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    void check(int rc) const { if (rc) throw std::runtime_error(std::string("libfemocs_b200: ") + fb_last_error(h)); }
    void set_option(const char* key, double v) { check(fb_set_option(h, key, v)); }
    long kernel_launches() const { return fb_kernel_launches(h); }
    fb_ctx* h;
};

// the subset of Config::Field the hot path reads (defaults: src/Config.cpp:63-71)
struct FieldConfig {
    double E0 = 0, V0 = 0, ssor_param = 1.2, cg_tolerance = 1e-9, V_min = -1.0, V_max = 1e4;
    int n_cg = 10000;
    std::string anode_BC = "neumann", mode = "laplace";
};

// view of the TetgenMesh arrays the path consumes (SURVEY.md 8a'); pointers stay owned by the mesher
struct MeshArrays {
    const double* nodes = nullptr; int n_nodes = 0;          // TetgenNodes::get()            include/TetgenCells.h:288
    const int* node_markers = nullptr;                       // 1..4 = tet node / edge / face / cell centroid
    const int* hexs = nullptr; const int* hex_markers = nullptr; int n_hexs = 0;   // Hexahedra::get() :661
    const int* tets = nullptr; const int* tet_nbrs = nullptr; const int* tet_markers = nullptr; int n_tets = 0;
    const int* tris = nullptr; const int* tri2tet = nullptr; const double* tri_norms = nullptr; int n_tris = 0;
    const int* quads = nullptr; const int* quad2hex = nullptr; int n_quads = 0;
    double tet_edgemax = 1.0;                                // tets.stat.edgemax
    const int* voro_off = nullptr; const int* voro_list = nullptr; int n_voro = 0;   // calc_pseudo_3D_vorocells
};

class PoissonSolver {
public:
    struct Stat { double sol_min = 0, sol_max = 0; } stat;

    PoissonSolver(Context& c, const FieldConfig* conf) : ctx(c), conf(conf) {}

    // PoissonSolver::set_particles: SuperParticle positions + solver cell ids (src/PoissonSolver.cpp:299-319);
    // charge_factor = q_over_eps0 * Wsp
    void set_particles(const double* xyz, const int* cells, long n, double charge_factor) {
        p_xyz = xyz; p_cell = cells; n_parts = n; p_factor = charge_factor;
    }

    // DealSolver::import_mesh (src/DealSolver.cpp:191-209) fed with the mesher's arrays (src/ProjectRunaway.cpp:216)
    bool import_mesh(const MeshArrays& m) {
        const int rc = fb_import_mesh(ctx.h, m.nodes, m.n_nodes, m.hexs, m.hex_markers, m.n_hexs);
        if (rc == FB_ERR_MESH) return false;
        ctx.check(rc);
        long sz[7]; fb_get_sizes(ctx.h, sz);
        n_dofs = (int) sz[0]; n_cells = (int) sz[1]; n_vertices = (int) sz[3];
        return true;
    }
    // same signature as the reference: all mesh vertices + the vacuum cells (Point = anything with operator[],
    // Cell = anything with .vertices[8], e.g. dealii::Point<3> / dealii::CellData<3>)
    template <class Point, class Cell>
    bool import_mesh(const std::vector<Point>& vertices, const std::vector<Cell>& cells) {
        std::vector<double> xyz(3 * vertices.size());
        for (size_t i = 0; i < vertices.size(); ++i) for (int d = 0; d < 3; ++d) xyz[3 * i + d] = vertices[i][d];
        std::vector<int> hex(8 * cells.size()), mark(cells.size(), 1);
        for (size_t i = 0; i < cells.size(); ++i) for (int k = 0; k < 8; ++k) hex[8 * i + k] = (int) cells[i].vertices[k];
        MeshArrays m; m.nodes = xyz.data(); m.n_nodes = (int) vertices.size();
        m.hexs = hex.data(); m.hex_markers = mark.data(); m.n_hexs = (int) cells.size();
        return import_mesh(m);
    }

    void setup(double field, double potential) {                       // src/PoissonSolver.cpp:162-167
        ctx.check(fb_poisson_setup(ctx.h, field, potential, lower(conf->anode_BC) == "dirichlet"));
    }
    void assemble(bool first_time) {                                   // src/PoissonSolver.cpp:170-210
        const bool sc = conf->mode != "laplace" && n_parts > 0;
        ctx.check(fb_poisson_assemble(ctx.h, first_time, sc ? p_xyz : nullptr, sc ? p_cell : nullptr, sc ? n_parts : 0, p_factor));
    }
    int solve() {                                                      // include/PoissonSolver.h:54
        int it = 0;
        ctx.check(fb_poisson_solve(ctx.h, conf->n_cg, conf->cg_tolerance, FB_PRECOND_JACOBI, &it, &last_residual));
        return it;
    }
    bool check_limits(double low_limit, double high_limit) {           // src/DealSolver.cpp:157-167
        int bad = 0;
        ctx.check(fb_check_limits(ctx.h, low_limit, high_limit, &bad, &stat.sol_min, &stat.sol_max));
        return bad != 0;
    }
    void export_solution(std::vector<double>& solution) const {        // src/DealSolver.cpp:269-278
        solution.resize(n_vertices);
        ctx.check(fb_export_solution(ctx.h, solution.data()));
    }
    void export_charge_dens(std::vector<double>& charge_dens) const {  // src/PoissonSolver.cpp:141-149
        charge_dens.resize(n_vertices);
        ctx.check(fb_export_charge_dens(ctx.h, charge_dens.data()));
    }
    double get_cell_vol(int i) {                                       // src/DealSolver.cpp:169-173
        if (cell_vol.empty()) { cell_vol.resize(n_cells); ctx.check(fb_get_cell_volumes(ctx.h, cell_vol.data())); }
        return cell_vol[i];
    }
    int get_n_cells() const { return n_cells; }
    int size() const { return n_dofs; }
    std::string to_str() const {                                       // include/DealSolver.h:107-117
        std::ostringstream ss; ss << "#elems=" << n_cells << ", #nodes=" << n_vertices << ", #dofs=" << n_dofs; return ss.str();
    }

    Context& ctx;
    double last_residual = 0;

private:
    static std::string lower(std::string s) { for (auto& ch : s) ch = (char) std::tolower(ch); return s; }
    const FieldConfig* conf;
    const double* p_xyz = nullptr; const int* p_cell = nullptr; long n_parts = 0; double p_factor = 0;
    int n_dofs = 0, n_cells = 0, n_vertices = 0;
    std::vector<double> cell_vol;
};

class Interpolator {
public:
    explicit Interpolator(Context& c) : ctx(c) {}
    // Interpolator::initialize(mesh, 0, TYPES.VACUUM) (src/Interpolator.cpp:28-77); needs the solver's import_mesh first
    void initialize(const MeshArrays& m) {
        n_nodes = m.n_nodes;
        ctx.check(fb_interp_initialize(ctx.h, m.node_markers, m.tets, m.tet_nbrs, m.tet_markers, m.n_tets, m.tris, m.tri2tet,
                                       m.tri_norms, m.n_tris, m.quads, m.quad2hex, m.n_quads, m.tet_edgemax,
                                       m.voro_off, m.voro_list, m.n_voro));
    }
    // Interpolator::extract_solution(PoissonSolver<3>&, bool smoothen) (src/Interpolator.cpp:172-190)
    void extract_solution(PoissonSolver& fem, bool smoothen) {
        if (&fem.ctx != &ctx) throw std::runtime_error("solver and interpolator live in different contexts");
        ctx.check(fb_extract_solution(ctx.h, smoothen));
    }
    // nodes.get_solutions(): Solution{vector, scalar1 = charge density, scalar2 = potential} per femocs node
    void get_solutions(std::vector<double>& sol5) const { sol5.resize(5 * (size_t) n_nodes); ctx.check(fb_get_nodal_solutions(ctx.h, sol5.data())); }
    Context& ctx;
    int n_nodes = 0;
};

class FieldReader {
public:
    explicit FieldReader(Interpolator* i) : interpolator(i) {}
    // SolutionReader::set_preferences (include/SolutionReader.h:60-67)
    void set_preferences(bool sort_atoms, int d, int r) {
        if (sort_atoms) throw std::runtime_error("atom sorting is not part of the B200 hot path");
        if ((d != 2 && d != 3) || r < 1 || r > 3) throw std::runtime_error("Invalid interpolation dimension/rank");
        dim = d; rank = r;
    }
    int size() const { return (int) markers.size(); }
    // SolutionReader::interpolate(n, x, y, z) (src/SolutionReader.cpp:428-436): three coordinate arrays as in Femocs_wrap.h:36
    void interpolate(int n_points, const double* x, const double* y, const double* z) {
        markers.assign(n_points, 0); interpolation.assign(5 * (size_t) n_points, 0.0); ids.resize(n_points);
        for (int i = 0; i < n_points; ++i) ids[i] = i;
        px.assign(x, x + n_points); py.assign(y, y + n_points); pz.assign(z, z + n_points);
        calc_full();
    }
    // SolutionReader::calc_interpolation (src/SolutionReader.cpp:167-190): re-interpolate with the cached cells
    void calc_interpolation() {
        if (!mapped) { calc_full(); return; }
        const int n = size();
        if (n) interpolator->ctx.check(fb_interpolate(interpolator->ctx.h, dim, rank, n, px.data(), py.data(), pz.data(), 1,
                                                      markers.data(), interpolation.data()));
        norms();
    }
    void update_positions(const double* x, const double* y, const double* z) {
        px.assign(x, x + size()); py.assign(y, y + size()); pz.assign(z, z + size());
    }
    // SolutionReader::export_results (src/SolutionReader.cpp:303-398); returns 1 when there is nothing to export (:305)
    int export_results(int n_points, const std::string& data_type, double* data) const {
        if (size() == 0) return 1;
        std::string low = data_type; for (auto& ch : low) ch = (char) std::tolower(ch);
        const bool append = (data_type == low);
        const int n = size();
        if (low == "elfield") {
            if (!append) std::fill(data, data + 3 * (size_t) n_points, 0.0);
            for (int i = 0; i < n; ++i) {
                const int id = ids[i];
                if (id < 0 || id >= n_points) continue;
                for (int d = 0; d < 3; ++d) {
                    if (append) data[3 * id + d] += interpolation[5 * (size_t) i + d];
                    else data[3 * id + d] = interpolation[5 * (size_t) i + d];
                }
            }
            return 0;
        }
        int slot;
        if (low == "elfield_norm") slot = -1; else if (low == "charge_density") slot = 3; else if (low == "potential") slot = 4;
        else throw std::runtime_error("SolutionReader does not contain " + data_type);
        if (!append) std::fill(data, data + n_points, 0.0);
        for (int i = 0; i < n; ++i) {
            const int id = ids[i];
            if (id < 0 || id >= n_points) continue;
            const double v = slot < 0 ? field_norm[i] : interpolation[5 * (size_t) i + slot];
            if (append) data[id] += v; else data[id] = v;
        }
        return 0;
    }
    // flag of femocs_interpolate_*: 1 = point was located inside the mesh
    void export_flags(int n_points, int* flag) const {
        for (int i = 0; i < size() && i < n_points; ++i) flag[ids[i]] = markers[i] >= 0;
    }
    std::vector<int> markers;                // atom.marker = located cell (negative: outside, nearest cell)
    std::vector<double> interpolation;       // 5 doubles per point: Ex, Ey, Ez, charge density, potential
    std::vector<double> field_norm;
    double E_max = -1e100;

private:
    void calc_full() {                       // SolutionReader::calc_full_interpolation (src/SolutionReader.cpp:136-165)
        const int n = size();
        if (n) interpolator->ctx.check(fb_locate_interpolate(interpolator->ctx.h, dim, rank, n, px.data(), py.data(), pz.data(), 1,
                                                             markers.data(), interpolation.data()));
        mapped = true;
        norms();
    }
    void norms() {                           // FieldReader::calc_interpolation (src/SolutionReader.cpp:489-499)
        field_norm.resize(size()); E_max = -1e100;
        for (int i = 0; i < size(); ++i) {
            const double* s = &interpolation[5 * (size_t) i];
            field_norm[i] = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
            E_max = std::max(E_max, field_norm[i]);
        }
    }
    Interpolator* interpolator;
    int dim = 3, rank = 1;
    bool mapped = false;
    std::vector<int> ids;
    std::vector<double> px, py, pz;
};

class Pic {
public:
    explicit Pic(Interpolator* i) : interpolator(i) {}
    // Pic::update_point_cell for every particle (src/Pic.cpp:186-196): solver cell guess in, located cell (or -1) out
    void update_point_cells(long n, const double* xyz, int* cells) { interpolator->ctx.check(fb_particle_cells(interpolator->ctx.h, n, xyz, cells)); }
    // field look-up of Pic::update_velocities (src/Pic.cpp:198-209)
    void fields(long n, const double* xyz, const int* cells, double* E3) { interpolator->ctx.check(fb_particle_field(interpolator->ctx.h, n, xyz, cells, E3)); }
    // Pic::update_positions + ParticleSpecies::clear_lost (src/Pic.cpp:137-184, src/ParticleSpecies.cpp:16-31): arrays updated in
    // place, the first n - n_lost entries are the survivors in their original order; returns n_lost
    int update_positions(long n, double* pos3, double* vel3, int* cells, double dt, const double box6[6], bool periodic) {
        long lost = 0;
        interpolator->ctx.check(fb_pic_update_positions(interpolator->ctx.h, n, pos3, vel3, cells, dt, box6, periodic ? 1 : 0, &lost));
        return (int) lost;
    }
    // Pic::update_velocities (src/Pic.cpp:198-209)
    void update_velocities(long n, const double* pos3, const int* cells, double* vel3, double dt, double q_over_m) {
        interpolator->ctx.check(fb_pic_update_velocities(interpolator->ctx.h, n, pos3, cells, vel3, dt, q_over_m));
    }
private:
    Interpolator* interpolator;
};

}  // namespace femocs_b200
#endif  // FEMOCS_B200_HPP_
