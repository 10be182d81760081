// TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
// Declaration-level stand-in for the reference's include/DealSolver.h so that the
// reference's *non-solver* sources (Interpolator.cpp, SolutionReader.cpp, ...)
// can be compiled verbatim in an image without deal.II.  Only the members those
// sources call are provided (reference call sites: Interpolator.cpp:175-198,
// SolutionReader.cpp:424,672,682); the arithmetic of the real class is restated
// in oracle/femocs_oracle.cpp, never here.  Vertex-ordered vectors are injected
// by the driver via set_exported().
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
    DealSolver() {}
    virtual ~DealSolver() {}

    void export_solution(vector<double>& sol) const { sol = vertex_solution; }
    void export_solution_grad(vector<Tensor<1, dim>>& grads) const { grads = vertex_grads; }
    void export_surface_centroids(Medium&) const {}
    void export_vertices(Medium&) {}
    void import_solution(const vector<double>* s) { if (s) vertex_solution = *s; }
    int size() const { return (int) vertex_solution.size(); }
    double get_cell_vol(const int) const { return 0; }
    int get_n_cells() const { return 0; }
    double max_solution() const { return 0; }

    struct Stat { double sol_min = 0, sol_max = 0; } stat;

    /** oracle hook: vertex-ordered solution to be handed to Interpolator::extract_solution */
    vector<double> vertex_solution;
    vector<Tensor<1, dim>> vertex_grads;

protected:
    bool valid_extension(const string&) const { return false; }
};

}  // namespace femocs
#endif
