// TEST INFRASTRUCTURE (oracle/_ref/libfemocs_dropin.so build only) -- not part of the product.
// This image has no deal.II, so the reference's include/DealSolver.h (deal.II Triangulation / SparseMatrix members)
// cannot be compiled here.  In a real FEMOCS build that header stays as it is (the heat solver derives from it); in
// this build a declaration-level stand-in takes its place so that the REST of the reference -- Femocs, Femocs_wrap,
// ProjectRunaway, Interpolator, SolutionReader, Pic ... -- compiles verbatim against the product header
// include/dropin/PoissonSolver.h.  Only what those sources reference is declared; nothing here computes.
#ifndef DEALSOLVER_H_
#define DEALSOLVER_H_

#include "deal.II/numerics/vector_tools.h"
#include "Globals.h"
#include "Medium.h"
#include "FileWriter.h"

using namespace dealii;
using namespace std;

namespace femocs {

template<int dim> class PoissonSolver;
template<int dim> class CurrentHeatSolver;

template<int dim>
class DealSolver : public FileWriter {
public:
    DealSolver() { stat.sol_min = 0; stat.sol_max = 0; }
    virtual ~DealSolver() {}

    void export_solution(vector<double>& sol) const { sol.clear(); }
    void export_solution_grad(vector<Tensor<1, dim>>& grads) const { grads.clear(); }
    void export_surface_centroids(Medium&) const {}
    void export_vertices(Medium&) {}
    void export_dofs(vector<Point<dim>>& points) const { points.clear(); }
    void import_solution(const vector<double>*) {}
    bool import_mesh(vector<Point<dim>>, vector<CellData<dim>>) { return true; }
    bool check_limits(const double, const double) { return false; }
    int size() const { return 0; }
    double get_cell_vol(const int) const { return 0; }
    int get_n_cells() const { return 0; }
    double max_solution() const { return 0; }
    string to_str() const { return "heat solver not built in this image"; }

    struct Stat {
        double sol_min, sol_max;
        friend ostream& operator <<(ostream &os, const Stat &s) { os << "min=" << s.sol_min << ", max=" << s.sol_max; return os; }
    } stat;

protected:
    bool valid_extension(const string&) const { return false; }
};

}  // namespace femocs
#endif
