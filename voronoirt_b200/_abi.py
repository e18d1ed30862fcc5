"""ctypes mirror of include/vrt.h (struct layouts and function prototypes).

Nothing here computes anything: it only describes the C ABI so that the host layer (api.py) and the
tests can call libvrt.so with plain pointers and sizes.
"""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)


class vrt_line(C.Structure):
    """include/vrt.h `vrt_line` — scalar fields of HydrogenicLine (reference src/line.jl:14-72)."""
    _fields_ = [
        ("nlam", C.c_int64),
        ("lidx", C.c_int64 * 4),
        ("lambda0", C.c_double),
        ("Aji", C.c_double), ("Bji", C.c_double), ("Bij", C.c_double),
        ("chi_i", C.c_double), ("chi_j", C.c_double), ("chi_inf", C.c_double),
        ("gi", C.c_int64), ("gj", C.c_int64), ("Z", C.c_int64),
        ("atom_weight", C.c_double),
        ("c_unsold", C.c_double), ("gamma_natural", C.c_double),
        ("c_linear_stark", C.c_double), ("c_quadratic_stark", C.c_double),
    ]


class vrt_site_data(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in (
        "temperature", "electron_density", "hydrogen_density",
        "velocity_z", "velocity_x", "velocity_y", "doppler_width",
        "alpha_cont", "destruction", "C", "lte_pops")]


class vrt_quadrature(C.Structure):
    _fields_ = [("n_dirs", C.c_int64), ("weights", C.c_void_p), ("theta", C.c_void_p), ("phi", C.c_void_p)]


class vrt_config(C.Structure):
    _fields_ = [
        ("n_sweeps", C.c_int32), ("dir_begin", C.c_int32),
        ("p", C.c_double),
        ("lam_begin", C.c_int64), ("lam_end", C.c_int64), ("lam_chunk", C.c_int64),
        ("prune", C.c_int32), ("dir_end", C.c_int32),
        ("cell_shard_rank", C.c_int32), ("cell_shard_count", C.c_int32),
    ]


class vrt_iter_info(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("reserved0", C.c_int32),
        ("diff", C.c_double),
        ("t_opacity_ms", C.c_double), ("t_sweep_ms", C.c_double), ("t_source_ms", C.c_double),
        ("t_rates_ms", C.c_double), ("t_stateq_ms", C.c_double), ("t_total_ms", C.c_double),
        ("updates", C.c_double),
    ]


class vrt_result(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("diff", C.c_double), ("seconds", C.c_double)]


vrt_allreduce_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p)
vrt_iter_cb = C.CFUNCTYPE(C.c_int, C.POINTER(vrt_iter_info), C.c_void_p)

P = C.c_void_p  # array pointers are passed as raw addresses (host or device)

# name -> (restype, argtypes); every symbol include/vrt.h declares
PROTOTYPES = {
    "vrt_abi_version": (C.c_int, []),
    "vrt_last_error": (C.c_char_p, []),
    "vrt_device_count": (C.c_int, [c_int32_p]),
    "vrt_set_device": (C.c_int, [C.c_int32]),
    "vrt_read_neighbours": (C.c_int, [C.c_char_p, C.c_int64, P, C.c_int64, c_int64_p]),
    "vrt_grid_create": (C.c_int, [C.c_int64, P, P, C.c_int64, P, C.POINTER(C.c_void_p)]),
    "vrt_grid_destroy": (None, [C.c_void_p]),
    "vrt_grid_size": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "vrt_grid_num_layers": (C.c_int, [C.c_void_p, C.c_int32, c_int64_p]),
    "vrt_grid_get_layers": (C.c_int, [C.c_void_p, C.c_int32, P, P]),
    "vrt_grid_get_delaunay_lines": (C.c_int, [C.c_void_p, P]),
    "vrt_grid_get_stencil": (C.c_int, [C.c_void_p, P, C.c_double, P, P, P, P]),
    "vrt_grid_get_schedule": (C.c_int, [C.c_void_p, P, C.c_int32, C.c_int32, C.c_int32, P, P, P, c_int64_p, c_int64_p]),
    "vrt_grid_release_schedules": (C.c_int, [C.c_void_p]),
    "vrt_formal_solve": (C.c_int, [C.c_void_p, P, C.c_int32, C.c_double, C.c_int32, C.c_int64, P, P, P, P]),
    "vrt_regular_formal_solve": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, P, P, P, P, C.c_int32, C.c_int32, C.c_int64,
                                           P, P, P, P, P]),
    "vrt_regular_mean_intensity": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, P, P, P, C.POINTER(vrt_quadrature), C.c_int32, C.c_int64,
                                             P, P, P, P, P]),
    "vrt_regular_lambda_iterate": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, P, P, P, C.POINTER(vrt_quadrature), C.c_int32, P, P, P,
                                             C.c_double, C.c_int32, vrt_iter_cb, C.c_void_p, P, P, C.POINTER(vrt_result)]),
    "vrt_voronoi_neighbours": (C.c_int, [C.c_int64, P, P, P, C.c_int64, c_int64_p]),
    "vrt_trilinear": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, P, P, P, P, C.c_int64, P, P]),
    "vrt_nearest_site": (C.c_int, [C.c_int64, P, P, C.c_int64, P, P, P]),
    "vrt_rejection_sampling": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, P, P, P, P, C.c_uint64, P, c_double_p]),
    "vrt_nearest_sites": (C.c_int, [C.c_int64, P, P, C.c_int64, P, C.c_int32, P, P]),
    "vrt_nearest_corner": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, P, P, P, P, C.c_int64, P, P]),
    "vrt_regular_grid_create": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, P, P, P, C.POINTER(C.c_void_p)]),
    "vrt_regular_release_workspace": (C.c_int, []),
    "vrt_solver_create_line": (C.c_int, [C.c_void_p, C.POINTER(vrt_line), P, C.POINTER(vrt_site_data),
                                         C.POINTER(vrt_quadrature), C.POINTER(vrt_config), C.POINTER(C.c_void_p)]),
    "vrt_solver_create_continuum": (C.c_int, [C.c_void_p, P, P, P, C.POINTER(vrt_quadrature),
                                              C.POINTER(vrt_config), C.POINTER(C.c_void_p)]),
    "vrt_solver_destroy": (None, [C.c_void_p]),
    "vrt_solver_set_allreduce": (C.c_int, [C.c_void_p, vrt_allreduce_fn, C.c_void_p]),
    "vrt_solver_set_field": (C.c_int, [C.c_void_p, C.c_int32, P]),
    "vrt_solver_nlam_local": (C.c_int, [C.c_void_p, c_int64_p]),
    "vrt_mean_intensity": (C.c_int, [C.c_void_p, P, P, P, P]),
    "vrt_calculate_R": (C.c_int, [C.c_void_p, P, P, P]),
    "vrt_get_revised_populations": (C.c_int, [C.c_int64, P, P, P, P]),
    "vrt_lambda_iterate": (C.c_int, [C.c_void_p, C.c_double, C.c_int32, vrt_iter_cb, C.c_void_p, C.POINTER(vrt_result)]),
    "vrt_get_state": (C.c_int, [C.c_void_p, P, P, P]),
    "vrt_set_state": (C.c_int, [C.c_void_p, P, P]),
    "vrt_state_checksum": (C.c_int, [C.c_void_p, c_double_p]),
    "vrt_output_create": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "vrt_output_create_regular": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "vrt_output_dataset_size": (C.c_int, [C.c_void_p, C.c_char_p, c_int64_p]),
    "vrt_output_write": (C.c_int, [C.c_void_p, C.c_char_p, P, C.c_int64]),
    "vrt_output_write_convergence": (C.c_int, [C.c_void_p, C.c_int64, C.c_double]),
    "vrt_output_write_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrt_output_close": (C.c_int, [C.c_void_p]),
    "vrt_nccl_available": (C.c_int, []),
    "vrt_nccl_version": (C.c_int, [c_int32_p]),
    "vrt_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "vrt_solver_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]),
    "vrt_solver_peer_handle": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vrt_solver_peer_attach": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "vrt_solver_peer_detach": (C.c_int, [C.c_void_p]),
    "vrt_solver_set_direction_lambda": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64]),
    "vrt_solver_direction_visits": (C.c_int, [C.c_void_p, c_int64_p, c_double_p, C.c_int64]),
    "vrt_solver_cell_slice": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "vrt_get_state_slice": (C.c_int, [C.c_void_p, P, P, P]),
    "vrt_set_state_slice": (C.c_int, [C.c_void_p, P, P]),
    "vrt_last_stats": (C.c_int, [c_double_p]),
}


def bind(lib):
    """Attach restype/argtypes for every ABI symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
