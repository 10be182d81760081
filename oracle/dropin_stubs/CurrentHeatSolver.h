// TEST INFRASTRUCTURE (oracle/_ref/libfemocs_dropin.so build only) -- not part of the product.
// Stand-in for the reference's include/CurrentHeatSolver.h (deal.II, not buildable in this image; see DealSolver.h
// beside this file).  The current / heat solvers are row f3 of SURVEY.md section 8; with heat_mode = none and
// field_mode = laplace -- the configuration the drop-in test runs -- none of these members is ever called.
#ifndef CURRENTSANDHEATING_H_
#define CURRENTSANDHEATING_H_

#include "DealSolver.h"
#include "PhysicalQuantities.h"
#include "Config.h"
#include "PoissonSolver.h"

namespace femocs {

class EmissionReader;

template<int dim> class EmissionSolver : public DealSolver<dim> {
public:
    int solve() { return 0; }
};
template<int dim> class CurrentSolver : public EmissionSolver<dim> {
public:
    void assemble() {}
};
template<int dim> class HeatSolver : public EmissionSolver<dim> {
public:
    void assemble(const double) {}
};

template<int dim>
class CurrentHeatSolver : public DealSolver<dim> {
public:
    CurrentHeatSolver() {}
    CurrentHeatSolver(PhysicalQuantities*, const Config::Heating*, EmissionReader*) {}
    void export_temp_rho(vector<double>& temp, vector<Tensor<1, dim>>& rho) const { temp.clear(); rho.clear(); }
    void set_dependencies(PhysicalQuantities*, const Config::Heating*) {}
    void setup(const double) {}
    int size() const { return heat.size(); }
    HeatSolver<dim> heat;
    CurrentSolver<dim> current;
};

}  // namespace femocs
#endif
