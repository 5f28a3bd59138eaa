// Temporary: entry points that are being implemented (removed as each lands).
#include "atx_potential_common.cuh"
#define NOTIMPL(name) atx_set_error(name ": not implemented yet"); return ATX_ERROR_UNSPECIFIED;
extern "C" int atx_rebo2_create(atx_ctx *, const atx_rebo2_params *, atx_rebo2 **) { NOTIMPL("atx_rebo2_create") }
extern "C" int atx_rebo2_destroy(atx_rebo2 *) { return 0; }
extern "C" int atx_rebo2_bind_to(atx_rebo2 *, atx_particles *, atx_neighbors *, int, const int *) { NOTIMPL("atx_rebo2_bind_to") }
extern "C" int atx_rebo2_energy_and_forces(atx_rebo2 *, atx_particles *, atx_neighbors *, double *, double *, double *, double *, double *, double *, double *, double *) { NOTIMPL("atx_rebo2_energy_and_forces") }

int atx_rebo2_compute_device(atx_rebo2 *, atx_particles *, atx_neighbors *, const PotOut &) { NOTIMPL("atx_rebo2_compute_device") }
