/* TEST INFRASTRUCTURE (oracle/_ref/libfemocs_dropin.so build only) -- not part of the product.
 * GETELEC is Fortran (no gfortran in this image) and field emission is outside the hot path: the two entry points
 * EmissionReader calls (reference src/EmissionReader.cpp:200-213) report zero emission. */
#include <stddef.h>
#include "getelec.h"

int cur_dens_c(struct emission* e) { e->Jem = 0; e->heat = 0; e->theta = 1; e->ierr = 0; e->regime = 0; e->sharp = 0; return 0; }
int cur_dens_SC(struct emission* e) { return cur_dens_c(e); }
