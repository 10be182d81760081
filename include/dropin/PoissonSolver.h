/*
 * PoissonSolver.h -- drop-in replacement of the reference's include/PoissonSolver.h (+ src/PoissonSolver.cpp and the
 * part of src/DealSolver.cpp that PoissonSolver inherits).  Put this directory in front of the reference's include/
 * on the include path, drop src/PoissonSolver.cpp from the build and link libfemocs_b200.so: Femocs, Femocs_wrap,
 * ProjectRunaway / ProjectSpaceCharge / ProjectHeat, Interpolator, SolutionReader and Pic compile UNCHANGED against
 * it (tests/test_dropin_femocs.py builds exactly that from the reference's own sources).
 *
 * femocs::PoissonSolver<dim> keeps the members the rest of FEMOCS calls, with the reference's signatures and return
 * conventions (call sites in parentheses, paths relative to the reference root):
 *   PoissonSolver(const ParticleSpecies*, const Config::Field*, const LinearHexahedra*)   (src/ProjectRunaway.cpp:38)
 *   set_particles                                   (:45)        import_mesh(vertices, cells) -> bool   (:216)
 *   setup(field, potential)                         (:424,458)   assemble(first_time)                   (:425,497)
 *   solve() -> +#CG / -#CG                          (:431,498)   check_limits + stat + operator<<       (:475,502-503)
 *   to_str()                                        (:428,461)   export_solution / export_charge_dens / export_solution_grad
 *                                                                                    (src/Interpolator.cpp:175-176,196-198)
 *   get_cell_vol / get_n_cells                      (src/Pic.cpp:231,311)   write("*.vtk|vtks|msh") through FileWriter
 * Everything numerical is forwarded to the C ABI of include/femocs_b200.h (CUDA, sm_100a).  There is no CPU fallback:
 * without a usable GPU import_mesh() returns false and FEMOCS reports "Importing vacuum mesh to Deal.II failed".
 *
 * The class does NOT derive from the reference's DealSolver<dim>: CurrentHeatSolver keeps that (deal.II) base class
 * untouched, and no caller needs PoissonSolver to be one (grep "DealSolver<3>&": only SolutionReader::interpolate(
 * const DealSolver<3>&), used with the heat solver).  probe_*, shape_funs, get_triangulation and get_dof_handler have
 * no callers outside the solver and are not provided.
 */
#ifndef LAPLACE_H_      /* the include guard of the header this one replaces */
#define LAPLACE_H_

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "DealSolver.h"          /* whichever is on the path: brings dealii::Point / Tensor / CellData as in the reference */
#include "Config.h"
#include "InterpolatorCells.h"
#include "ParticleSpecies.h"

#include "femocs_b200.h"

namespace femocs {
using namespace dealii;
using namespace std;

template<int dim>
class PoissonSolver : public FileWriter {
public:
    PoissonSolver() : PoissonSolver(NULL, NULL, NULL) {}

    PoissonSolver(const ParticleSpecies* particles_, const Config::Field* conf_, const LinearHexahedra* interpolator_) :
            particles(particles_), conf(conf_), interpolator(interpolator_), ctx(NULL), n_verts(0), n_cells(0), n_dofs(0),
            n_faces(0), n_edges(0), precond(FB_PRECOND_JACOBI)
    {
        stat.sol_min = 0; stat.sol_max = 0;
        static_assert(dim == 3, "libfemocs_b200 solves on the 3D hexahedral mesh");
        // FEMOCS_B200_DEVICE picks the CUDA ordinal, FEMOCS_B200_PRECOND=chebyshev | twolevel another preconditioner
        const char* p = getenv("FEMOCS_B200_PRECOND");
        if (p && string(p) == "chebyshev") precond = FB_PRECOND_CHEBYSHEV;
        if (p && string(p) == "twolevel") precond = FB_PRECOND_TWOLEVEL;
    }

    PoissonSolver(const PoissonSolver&) = delete;
    PoissonSolver& operator=(const PoissonSolver&) = delete;

    ~PoissonSolver() { if (ctx) fb_destroy(ctx); }

    void set_particles(const ParticleSpecies* parts) { particles = parts; }

    /** DealSolver::import_mesh(vertices, cells), src/DealSolver.cpp:191-209: all mesh nodes and the vacuum hexahedra
     *  in old-style vertex order, as TetgenNodes::export_dealii / Hexahedra::export_vacuum hand them over by value */
    bool import_mesh(vector<Point<dim>> vertices, vector<CellData<dim>> cells) {
        if (!ctx) {
            const char* dev = getenv("FEMOCS_B200_DEVICE");
            ctx = fb_create(dev ? atoi(dev) : 0);
            if (!ctx) { write_silent_msg(string("libfemocs_b200: ") + fb_create_error()); return false; }
            // the element is a compile-time constant of the reference (include/DealSolver.h:130); here: shape_degree below
            if (shape_degree != 1 && fb_set_option(ctx, "fe_degree", (double) shape_degree)) { complain("fe_degree"); return false; }
        }
        const size_t nv = vertices.size(), nc = cells.size();
        vector<double> xyz(3 * nv);
        for (size_t i = 0; i < nv; ++i)
            for (int d = 0; d < 3; ++d) xyz[3 * i + d] = vertices[i][d];
        vector<int> hex8(8 * nc), marker(nc, 1);
        for (size_t i = 0; i < nc; ++i)
            for (int k = 0; k < 8; ++k) hex8[8 * i + k] = (int) cells[i].vertices[k];
        if (nv == 0 || nc == 0) return false;
        if (fb_import_mesh(ctx, xyz.data(), (int) nv, hex8.data(), marker.data(), (int) nc)) return complain("import_mesh"), false;
        long sz[7];
        fb_get_sizes(ctx, sz);
        n_dofs = (int) sz[0]; n_cells = (int) sz[1]; n_verts = (int) sz[3];
        fb_get_mesh_counts(ctx, &n_faces, &n_edges);
        cell_vol.clear();
        return true;
    }

    /** src/PoissonSolver.cpp:162-167 (+ DealSolver::setup_system: zero system, solution = 0) */
    void setup(const double field, const double potential) {
        require(conf, "NULL conf can't be used!");
        if (fb_poisson_setup(ctx, field, potential, conf->anode_BC == "dirichlet")) complain("setup");
    }

    /** src/PoissonSolver.cpp:170-210: stiffness matrix (first_time) or the saved one, copper Dirichlet, Neumann or
     *  Dirichlet anode, space charge of the super particles (assemble_space_charge_fast, :299-319) */
    void assemble(const bool first_time) {
        require(conf, "NULL conf can't be used!");
        require(conf->anode_BC == "neumann" || conf->anode_BC == "dirichlet", "Unimplemented anode BC: " + conf->anode_BC);
        long n = 0;
        double charge_factor = 0;
        if (conf->mode != "laplace" && particles && particles->size() > 0) {
            n = particles->size();
            charge_factor = particles->q_over_eps0 * particles->get_Wsp();
            part_xyz.resize(3 * n); part_cell.resize(n);
            long i = 0;
            for (SuperParticle const &sp : *particles) {
                part_xyz[3 * i] = sp.pos.x; part_xyz[3 * i + 1] = sp.pos.y; part_xyz[3 * i + 2] = sp.pos.z;
                part_cell[i++] = sp.cell;
            }
        }
        // src/PoissonSolver.cpp:196-207: the charge density (rhs / dof volume) is kept only when a file is about to be written
        fb_set_option(ctx, "charge_density", this->write_time() ? 1.0 : 0.0);
        if (fb_poisson_assemble(ctx, first_time, n ? part_xyz.data() : NULL, n ? part_cell.data() : NULL, n, charge_factor))
            complain("assemble");
    }

    /** include/PoissonSolver.h:54 -> DealSolver::solve_cg(n_cg, cg_tolerance, ssor_param), src/DealSolver.cpp:442-458:
     *  # CG iterations, negative when n_cg was reached.  ssor_param is accepted and unused: the GPU path
     *  preconditions with Jacobi or a Chebyshev polynomial (north star) and stops at the same absolute residual. */
    int solve() {
        require(conf, "NULL conf can't be used!");
        int it = 0; double res = 0;
        if (fb_poisson_solve(ctx, conf->n_cg, conf->cg_tolerance, precond, &it, &res)) { complain("solve"); return -abs(conf->n_cg) - 1; }
        return it;
    }

    /** src/DealSolver.cpp:157-167 */
    bool check_limits(const double low_limit, const double high_limit) {
        int bad = 1;
        if (fb_check_limits(ctx, low_limit, high_limit, &bad, &stat.sol_min, &stat.sol_max)) { complain("check_limits"); return true; }
        return bad != 0;
    }

    /** src/DealSolver.cpp:152-155 (linfty norm of the solution) */
    double max_solution() const {
        int bad; double lo = 0, hi = 0;
        fb_check_limits(ctx, -1e300, 1e300, &bad, &lo, &hi);
        return max(fabs(lo), fabs(hi));
    }

    /** src/DealSolver.cpp:269-278 / src/PoissonSolver.cpp:141-149 / src/DealSolver.cpp:280-301: vertex order */
    void export_solution(vector<double> &solution) const {
        solution.resize(n_verts);
        if (n_verts && fb_export_solution(ctx, solution.data())) complain("export_solution");
    }
    void export_charge_dens(vector<double> &charge_dens) const {
        charge_dens.resize(n_verts);
        if (n_verts && fb_export_charge_dens(ctx, charge_dens.data())) complain("export_charge_dens");
    }
    void export_solution_grad(vector<Tensor<1, dim>> &grads) const {
        grads.resize(n_verts);
        vector<double> g(3 * (size_t) n_verts);
        if (n_verts && fb_export_solution_grad(ctx, g.data())) complain("export_solution_grad");
        for (int i = 0; i < n_verts; ++i)
            for (int d = 0; d < 3; ++d) grads[i][d] = g[3 * (size_t) i + d];
    }

    /** src/DealSolver.cpp:303-315 */
    void import_solution(const vector<double>* new_solution) {
        require(new_solution, "Can't use NULL solution vector!");
        require((int) new_solution->size() == n_verts, "Mismatch between #vertices and solution vector size: "
                + d2s(n_verts) + " vs " + d2s(new_solution->size()));
        if (fb_import_solution(ctx, new_solution->data())) complain("import_solution");
    }

    /** src/DealSolver.cpp:169-173; Pic calls it once per cell, so the volumes are fetched once per mesh */
    double get_cell_vol(const int i) const {
        if (cell_vol.empty() && n_cells > 0) {
            cell_vol.resize(n_cells);
            if (fb_get_cell_volumes(ctx, cell_vol.data())) complain("get_cell_vol");
        }
        return cell_vol[i];
    }
    int get_n_cells() const { return n_cells; }
    int size() const { return n_dofs; }

    /** include/DealSolver.h:107-117 */
    friend ostream& operator <<(ostream &os, const PoissonSolver<dim>& d) {
        os << "#elems=" << d.n_cells << ", #faces=" << d.n_faces << ", #edges=" << d.n_edges
                << ", #nodes=" << d.n_verts << ", #dofs=" << d.n_dofs;
        return os;
    }
    string to_str() const { ostringstream ss; ss << (*this); return ss.str(); }

    /** include/DealSolver.h:119-127 */
    struct Stat {
        double sol_min;
        double sol_max;
        friend ostream& operator <<(ostream &os, const Stat &s) {
            os << "min=" << s.sol_min << ", max=" << s.sol_max;
            return os;
        }
    } stat;

    /** the device context (for host codes that also route their interpolation through libfemocs_b200) */
    fb_ctx* context() const { return ctx; }

protected:
    bool valid_extension(const string &ext) const { return ext == "vtk" || ext == "vtks" || ext == "msh"; }

    /** potential on the solver mesh as legacy VTK (the reference: deal.II DataOut, src/PoissonSolver.cpp:322-336) */
    void write_vtk(ofstream& out) const {
        vector<double> xyz(3 * (size_t) n_verts), phi;
        vector<int> cells(8 * (size_t) n_cells);
        if (fb_get_solver_mesh(ctx, xyz.data(), cells.data())) { complain("write_vtk"); return; }
        export_solution(phi);
        out << "# vtk DataFile Version 3.0\n# " << to_str() << "\nASCII\nDATASET UNSTRUCTURED_GRID\n";
        out << "POINTS " << n_verts << " double\n";
        for (int i = 0; i < n_verts; ++i) out << xyz[3 * (size_t) i] << ' ' << xyz[3 * (size_t) i + 1] << ' ' << xyz[3 * (size_t) i + 2] << '\n';
        out << "CELLS " << n_cells << ' ' << 9 * (size_t) n_cells << '\n';
        for (int c = 0; c < n_cells; ++c) {
            out << 8;
            for (int k = 0; k < 8; ++k) out << ' ' << cells[8 * (size_t) c + k];
            out << '\n';
        }
        out << "CELL_TYPES " << n_cells << '\n';
        for (int c = 0; c < n_cells; ++c) out << "12\n";        // VTK_HEXAHEDRON (old-style vertex order)
        out << "POINT_DATA " << n_verts << "\nSCALARS potential double 1\nLOOKUP_TABLE default\n";
        for (int i = 0; i < n_verts; ++i) out << phi[i] << '\n';
    }

    /** the solver mesh as Gmsh 2 ASCII (the reference: deal.II GridOut::write_msh, src/DealSolver.cpp:351-355) */
    void write_msh(ofstream& out) const {
        vector<double> xyz(3 * (size_t) n_verts);
        vector<int> cells(8 * (size_t) n_cells);
        if (fb_get_solver_mesh(ctx, xyz.data(), cells.data())) { complain("write_msh"); return; }
        out << "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n" << n_verts << '\n';
        for (int i = 0; i < n_verts; ++i)
            out << i + 1 << ' ' << xyz[3 * (size_t) i] << ' ' << xyz[3 * (size_t) i + 1] << ' ' << xyz[3 * (size_t) i + 2] << '\n';
        out << "$EndNodes\n$Elements\n" << n_cells << '\n';
        for (int c = 0; c < n_cells; ++c) {
            out << c + 1 << " 5 2 0 0";
            for (int k = 0; k < 8; ++k) out << ' ' << cells[8 * (size_t) c + k] + 1;
            out << '\n';
        }
        out << "$EndElements\n";
    }

private:
    /// degree of the shape functions, as include/DealSolver.h:130 of the reference; 2 selects the library's FE_Q(2) path
    /// (QGauss(3); space charge through the general path of PoissonSolver.cpp:276-296)
    static constexpr unsigned int shape_degree = 1;

    const ParticleSpecies* particles;
    const Config::Field* conf;
    const LinearHexahedra* interpolator;      ///< kept for signature compatibility; the space-charge weights are computed on the device

    fb_ctx* ctx;
    int n_verts, n_cells, n_dofs;
    long n_faces, n_edges;
    int precond;
    vector<double> part_xyz;
    vector<int> part_cell;
    mutable vector<double> cell_vol;

    void complain(const char* where) const {
        write_silent_msg(string("libfemocs_b200 ") + where + ": " + (ctx ? fb_last_error(ctx) : fb_create_error()));
    }
};

} // namespace femocs

#endif /* LAPLACE_H_ */
