// TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
// Stand-in for the reference's include/CurrentHeatSolver.h (see DealSolver.h stub).
// The current/heat solvers are outside the hot path.
#ifndef CURRENTSANDHEATING_H_
#define CURRENTSANDHEATING_H_

#include "DealSolver.h"
#include "PhysicalQuantities.h"
#include "Config.h"
#include "PoissonSolver.h"

namespace femocs {

template<int dim> class EmissionSolver : public DealSolver<dim> {};
template<int dim> class CurrentSolver : public EmissionSolver<dim> {};
template<int dim> class HeatSolver : public EmissionSolver<dim> {};

template<int dim>
class CurrentHeatSolver : public DealSolver<dim> {
public:
    CurrentHeatSolver() {}
    void export_temp_rho(vector<double>& temp, vector<Tensor<1, dim>>& rho) const { temp.clear(); rho.clear(); }
    int size() const { return heat.size(); }
    HeatSolver<dim> heat;
    CurrentSolver<dim> current;
};

}  // namespace femocs
#endif
