// TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
// Stand-in for the reference's include/PoissonSolver.h (see DealSolver.h stub).
#ifndef LAPLACE_H_
#define LAPLACE_H_

#include "DealSolver.h"
#include "Config.h"
#include "InterpolatorCells.h"
#include "ParticleSpecies.h"

namespace femocs {

template<int dim>
class PoissonSolver : public DealSolver<dim> {
public:
    PoissonSolver() {}
    PoissonSolver(const ParticleSpecies*, const Config::Field*, const LinearHexahedra*) {}
    void set_particles(const ParticleSpecies*) {}
    void export_charge_dens(vector<double>& rho) const { rho = vertex_charge; }

    vector<double> vertex_charge;   ///< oracle hook, vertex-ordered charge density
};

}  // namespace femocs
#endif
