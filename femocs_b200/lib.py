"""ctypes binding of libfemocs_b200.so (the C ABI declared in include/femocs_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device can be
initialised, importing / creating a context raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libfemocs_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)
vp = C.c_void_p

# name -> (restype, argtypes); pointers are passed as raw addresses (host or device)
SIGNATURES = {
    "fb_create": (vp, [C.c_int]),
    "fb_destroy": (None, [vp]),
    "fb_last_error": (C.c_char_p, [vp]),
    "fb_create_error": (C.c_char_p, []),
    "fb_kernel_launches": (C.c_long, [vp]),
    "fb_set_option": (C.c_int, [vp, C.c_char_p, C.c_double]),
    "fb_comm_unique_id": (C.c_int, [vp]),
    "fb_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "fb_get_partition": (C.c_int, [vp, vp]),
    "fb_get_cells27": (C.c_int, [vp, vp]),
    "fb_plan_create": (vp, [C.c_int, C.c_int]),
    "fb_plan_phase1": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "fb_plan_phase2": (C.c_int, [vp, vp]),
    "fb_plan_import": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int]),
    "fb_plan_sizes": (C.c_int, [vp, vp]),
    "fb_plan_get": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "fb_import_mesh": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int]),
    "fb_get_sizes": (C.c_int, [vp, vp]),
    "fb_poisson_setup": (C.c_int, [vp, C.c_double, C.c_double, C.c_int]),
    "fb_poisson_assemble": (C.c_int, [vp, C.c_int, vp, vp, C.c_long, C.c_double]),
    "fb_poisson_solve": (C.c_int, [vp, C.c_int, C.c_double, C.c_int, c_int_p, c_double_p]),
    "fb_export_solution": (C.c_int, [vp, vp]),
    "fb_export_charge_dens": (C.c_int, [vp, vp]),
    "fb_import_solution": (C.c_int, [vp, vp]),
    "fb_export_solution_grad": (C.c_int, [vp, vp]),
    "fb_get_mesh_counts": (C.c_int, [vp, c_long_p, c_long_p]),
    "fb_get_solver_mesh": (C.c_int, [vp, vp, vp]),
    "fb_check_limits": (C.c_int, [vp, C.c_double, C.c_double, c_int_p, c_double_p, c_double_p]),
    "fb_get_cell_volumes": (C.c_int, [vp, vp]),
    "fb_get_system": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "fb_interp_initialize": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int,
                                       C.c_double, vp, vp, C.c_int]),
    "fb_extract_solution": (C.c_int, [vp, C.c_int]),
    "fb_get_nodal_solutions": (C.c_int, [vp, vp]),
    "fb_set_nodal_solutions": (C.c_int, [vp, vp]),
    "fb_locate_interpolate": (C.c_int, [vp, C.c_int, C.c_int, C.c_long, vp, vp, vp, C.c_int, vp, vp]),
    "fb_locate_interpolate_chains": (C.c_int, [vp, C.c_int, C.c_int, C.c_long, C.c_long, vp, vp, vp]),
    "fb_interpolate": (C.c_int, [vp, C.c_int, C.c_int, C.c_long, vp, vp, vp, C.c_int, vp, vp]),
    "fb_particle_cells": (C.c_int, [vp, C.c_long, vp, vp]),
    "fb_particle_field": (C.c_int, [vp, C.c_long, vp, vp, vp]),
    "fb_locate_interpolate_dev": (C.c_int, [vp, C.c_int, C.c_int, C.c_long, vp, vp, vp]),
    "fb_particle_cells_dev": (C.c_int, [vp, C.c_long, vp, vp]),
    "fb_particle_field_dev": (C.c_int, [vp, C.c_long, vp, vp, vp]),
    "fb_poisson_assemble_dev": (C.c_int, [vp, C.c_int, vp, vp, C.c_long, C.c_double]),
    "fb_pic_update_positions": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_double, vp, C.c_int, c_long_p]),
    "fb_pic_update_velocities": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_double, C.c_double]),
    "fb_pic_update_positions_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_double, vp, C.c_int, c_long_p]),
    "fb_pic_update_velocities_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_double, C.c_double]),
    "fb_synchronize": (C.c_int, [vp]),
    "fb_last_solve_stats": (C.c_int, [vp, c_double_p, c_int_p, c_long_p]),
    "fb_last_solve_profile": (C.c_int, [vp, c_double_p, c_double_p, c_int_p]),
    "fb_last_solve_kernel": (C.c_int, [vp]),
    "fb_plan_interp_tables": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int,
                                        vp, vp, vp, vp, vp, vp, vp, vp]),
    "fb_plan_jds": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
    "fb_plan_jds_split": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
    "fb_plan_jds_get_split": (C.c_int, [vp, vp, vp]),
    "fb_plan_jds_get": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "fb_get_stream": (vp, [vp]),
    "fb_comm_mode": (C.c_int, [vp]),
    "fb_last_import_reused": (C.c_int, [vp]),
    "fb_plan_set_kind": (C.c_int, [vp, C.c_int]),
    "fb_import_bulk_mesh": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int]),
    "fb_ch_set_physics": (C.c_int, [vp, vp, vp, C.c_int, C.c_double]),
    "fb_ch_setup": (C.c_int, [vp, C.c_double]),
    "fb_export_surface_centroids": (C.c_int, [vp, vp, c_int_p]),
    "fb_current_assemble": (C.c_int, [vp, vp, C.c_int]),
    "fb_heat_assemble": (C.c_int, [vp, C.c_double, vp, C.c_int]),
    "fb_ch_solve": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, c_int_p, c_double_p]),
    "fb_ch_export_solution": (C.c_int, [vp, C.c_int, vp]),
    "fb_ch_import_solution": (C.c_int, [vp, C.c_int, vp]),
    "fb_ch_export_solution_grad": (C.c_int, [vp, C.c_int, vp]),
    "fb_ch_check_limits": (C.c_int, [vp, C.c_int, C.c_double, C.c_double, c_int_p, c_double_p, c_double_p]),
}

_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(femocs_b200 has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)           # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
