"""ctypes binding of libb200mf.so (the C ABI declared in include/b200mf.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C dealii_b200/csrc``.
There is no fallback: if the shared library is missing, importing the engine fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mf.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOCONVERGENCE, ERR_COMM = 0, -1, -2, -3, -4, -5
F64, F32 = 0, 1
GEOMETRY_Q1_VERTICES, GEOMETRY_JACOBIANS = 0, 1
CELLS_CARTESIAN, CELLS_AFFINE, CELLS_GENERAL = 0, 1, 2
L2G_CONSTRAINED = 0x80000000
PRECOND_NONE, PRECOND_JACOBI, PRECOND_CHEBYSHEV = 0, 1, 2
MESH_MORTON, MESH_LEXICOGRAPHIC = 0, 1
GHOSTS_RELEVANT, GHOSTS_TOUCHED = 0, 1
DEFORM_NONE, DEFORM_SINE = 0, 1

u64, u32p, u16p, f64p, vp = C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_double), C.c_void_p


class SetupDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("degree", C.c_int), ("n_q_points_1d", C.c_int), ("number", C.c_int),
                ("n_cells", u64), ("n_owned_dofs", u64), ("n_ghost_dofs", u64),
                ("local_to_global", vp), ("constraint_mask", vp),
                ("geometry", C.c_int), ("cell_vertices", vp), ("inv_jacobian", vp), ("JxW", vp),
                ("shape_values", vp), ("shape_gradients_collocation", vp), ("quadrature_weights", vp),
                ("subface_interpolation_matrix", vp),
                ("constrained_dofs", vp), ("n_constrained_dofs", u64),
                ("n_cells_interior", u64)]


class SetupInfo(C.Structure):
    _fields_ = [("dim", C.c_int), ("degree", C.c_int), ("n_q_points_1d", C.c_int), ("number", C.c_int),
                ("n_cells", u64), ("n_owned_dofs", u64), ("n_ghost_dofs", u64), ("n_constrained_dofs", u64),
                ("cell_kind", C.c_int), ("n_distinct_geometries", u64), ("device_bytes", u64),
                ("geometry_bytes", u64), ("index_bytes", u64),
                ("n_bricks", u64), ("cells_per_brick", u64)]


class BulkInfo(C.Structure):
    _fields_ = [("n_bricks", u64), ("n_patterns", u64), ("n_own", u64), ("n_first_scalar", u64),
                ("n_later", u64), ("n_zero", u64), ("n_general_cells", u64), ("n_boundary_bricks", u64),
                ("usable", C.c_int), ("enabled", C.c_int),
                ("tuned_ms_index_map", C.c_double), ("tuned_ms_bulk", C.c_double),
                ("tuned_ms_coloured", C.c_double), ("n_colours", C.c_int), ("n_coloured_launches", C.c_int),
                ("n_zero_coloured", u64), ("path", C.c_int), ("strided", C.c_int)]


class Operator(C.Structure):
    _fields_ = [("grad_coefficient", vp), ("mass_coefficient", vp),
                ("grad_constant", C.c_double), ("mass_constant", C.c_double)]


class SolverDesc(C.Structure):
    _fields_ = [("preconditioner", C.c_int), ("inverse_diagonal", vp),
                ("chebyshev_degree", C.c_int), ("smoothing_range", C.c_double),
                ("eig_cg_n_iterations", C.c_int), ("safety_factor", C.c_double),
                ("tolerance", C.c_double), ("max_iterations", C.c_int),
                ("first_owned_global_index", u64), ("check_every", C.c_int),
                ("max_eigenvalue", C.c_double), ("eig_keep_constrained_entries", C.c_int)]


class SolverResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("residual", C.c_double), ("initial_residual", C.c_double),
                ("chebyshev_max_eigenvalue", C.c_double), ("chebyshev_min_eigenvalue", C.c_double),
                ("operator_applications", u64)]


class MgDesc(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("levels", C.POINTER(vp)), ("operators", C.POINTER(Operator)),
                ("child_cells", C.POINTER(vp)), ("smoother_degree", C.c_int32),
                ("smoothing_range", C.c_double), ("eig_cg_n_iterations", C.c_int32),
                ("coarse_tolerance", C.c_double), ("safety_factor", C.c_double),
                ("partitioners", C.POINTER(vp))]


class MgLevelInfo(C.Structure):
    _fields_ = [("eig_min", C.c_double), ("eig_max", C.c_double), ("degree", C.c_int32),
                ("eig_cg_iterations", C.c_int32), ("n_dofs", u64), ("inverse_diagonal", vp)]


class PartitionerInfo(C.Structure):
    _fields_ = [("n_owned", u64), ("n_ghost", u64), ("n_import", u64),
                ("n_ghost_targets", C.c_int), ("n_import_targets", C.c_int),
                ("ghost_target_ranks", C.POINTER(C.c_int)), ("ghost_target_counts", C.POINTER(u64)),
                ("import_target_ranks", C.POINTER(C.c_int)), ("import_target_counts", C.POINTER(u64)),
                ("import_indices", u32p)]


class MeshDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("degree", C.c_int), ("cells_per_direction", C.c_int),
                ("cell_order", C.c_int), ("left", C.c_double), ("right", C.c_double),
                ("deformation", C.c_int), ("deformation_amplitude", C.c_double),
                ("dirichlet_boundary", C.c_int), ("mark_constrained_l2g", C.c_int), ("dof_numbering", C.c_int)]


class PartitionDesc(C.Structure):
    _fields_ = [("mesh", MeshDesc), ("coarse", C.c_int * 3), ("n_ranks", C.c_int), ("rank", C.c_int),
                ("ghost_mode", C.c_int), ("want_lattice_ids", C.c_int)]


class PartitionView(C.Structure):
    _fields_ = [("n_global_dofs", u64), ("n_global_cells", u64), ("first_owned_global", u64),
                ("n_owned", u64), ("n_ghost", u64), ("n_cells_interior", u64),
                ("rank_offsets", C.POINTER(u64)), ("ghost_global", C.POINTER(u64)),
                ("lattice_ids", C.POINTER(u64)), ("cell_morton_position", C.POINTER(u64))]


class AdaptiveDesc(C.Structure):
    _fields_ = [("part", PartitionDesc), ("ball_radius", C.c_double), ("brick_friendly_order", C.c_int)]


class AdaptiveView(C.Structure):
    _fields_ = [("constraint_mask", u16p), ("active_cell_index", u32p), ("n_hanging_dofs", u64),
                ("n_masked_cells", u64), ("dof_coords", f64p)]


class MeshView(C.Structure):
    _fields_ = [("n_cells", u64), ("n_dofs", u64), ("n_boundary_dofs", u64),
                ("dofs_per_cell", C.c_int), ("vertices_per_cell", C.c_int), ("dim", C.c_int),
                ("local_to_global", u32p), ("cell_vertices", f64p), ("boundary_dofs", u32p)]


# every symbol include/b200mf.h declares: (restype, argtypes)
SYMBOLS = {
    "b200mf_setup_create": (C.c_int, [C.POINTER(SetupDesc), C.POINTER(vp)]),
    "b200mf_setup_destroy": (C.c_int, [vp]),
    "b200mf_setup_get_info": (C.c_int, [vp, C.POINTER(SetupInfo)]),
    "b200mf_get_quadrature_points": (C.c_int, [vp, vp]),
    "b200mf_cell_loop": (C.c_int, [vp, C.POINTER(Operator), vp, vp, vp]),
    "b200mf_vmult": (C.c_int, [vp, C.POINTER(Operator), vp, vp, vp]),
    "b200mf_copy_constrained_values": (C.c_int, [vp, vp, vp, vp]),
    "b200mf_set_constrained_values": (C.c_int, [vp, vp, C.c_double, vp]),
    "b200mf_compute_diagonal": (C.c_int, [vp, C.POINTER(Operator), vp, vp]),
    "b200mf_vmult_host": (C.c_int, [vp, C.POINTER(Operator), vp, vp]),
    "b200mf_vec_set": (C.c_int, [C.c_int, vp, C.c_double, u64, vp]),
    "b200mf_vec_axpy": (C.c_int, [C.c_int, vp, C.c_double, vp, u64, vp]),
    "b200mf_vec_sadd": (C.c_int, [C.c_int, vp, C.c_double, C.c_double, vp, u64, vp]),
    "b200mf_vec_scale_by": (C.c_int, [C.c_int, vp, vp, vp, u64, vp]),
    "b200mf_vec_dot": (C.c_int, [C.c_int, vp, vp, u64, f64p, vp]),
    "b200mf_vec_dot_device": (C.c_int, [C.c_int, vp, vp, u64, vp, vp]),
    "b200mf_vec_norm_sqr": (C.c_int, [C.c_int, vp, u64, f64p, vp]),
    "b200mf_vec_norm_2": (C.c_int, [C.c_int, vp, u64, f64p, vp]),
    "b200mf_vec_norm_1": (C.c_int, [C.c_int, vp, u64, f64p, vp]),
    "b200mf_vec_norm_linfty": (C.c_int, [C.c_int, vp, u64, f64p, vp]),
    "b200mf_vec_equ": (C.c_int, [C.c_int, vp, C.c_double, vp, C.c_double, vp, u64, vp]),
    "b200mf_vec_sadd_xavbw": (C.c_int, [C.c_int, vp, C.c_double, C.c_double, vp, C.c_double, vp, u64, vp]),
    "b200mf_vec_scale": (C.c_int, [C.c_int, vp, C.c_double, vp, u64, vp]),
    "b200mf_vec_add_and_dot": (C.c_int, [C.c_int, vp, C.c_double, vp, vp, u64, f64p, vp]),
    "b200mf_comm_get_unique_id": (C.c_int, [vp]),
    "b200mf_comm_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "b200mf_comm_destroy": (C.c_int, [vp]),
    "b200mf_comm_allreduce_sum": (C.c_int, [vp, vp, C.c_int, vp]),
    "b200mf_partitioner_create": (C.c_int, [vp, C.POINTER(u64), C.POINTER(u64), u64, C.c_int, C.POINTER(vp)]),
    "b200mf_partitioner_create_host": (C.c_int, [C.c_int, C.c_int, C.POINTER(u64), C.POINTER(u64), u64, C.c_int,
                                                 C.POINTER(C.c_int), C.POINTER(u64), C.c_int, C.POINTER(u64),
                                                 C.POINTER(vp)]),
    "b200mf_partitioner_destroy": (C.c_int, [vp]),
    "b200mf_partitioner_get_info": (C.c_int, [vp, C.POINTER(PartitionerInfo)]),
    "b200mf_update_ghost_values": (C.c_int, [vp, vp, vp]),
    "b200mf_compress_add": (C.c_int, [vp, vp, vp]),
    "b200mf_zero_out_ghost_values": (C.c_int, [vp, vp, vp]),
    "b200mf_dist_vmult": (C.c_int, [vp, vp, C.POINTER(Operator), vp, vp, vp]),
    "b200mf_dist_vmult_host_batch": (C.c_int, [vp, vp, C.POINTER(Operator), C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "b200mf_dist_compute_diagonal": (C.c_int, [vp, vp, C.POINTER(Operator), vp, vp]),
    "b200mf_dist_cg_solve": (C.c_int, [vp, vp, C.POINTER(Operator), C.POINTER(SolverDesc), vp, vp,
                                       C.POINTER(SolverResult), vp]),
    "b200mf_cg_solve": (C.c_int, [vp, C.POINTER(Operator), C.POINTER(SolverDesc), vp, vp,
                                  C.POINTER(SolverResult), vp]),
    "b200mf_mg_create": (C.c_int, [C.POINTER(MgDesc), C.POINTER(vp), vp]),
    "b200mf_mg_destroy": (None, [vp]),
    "b200mf_mg_prolongation_matrix_1d": (C.c_int, [C.c_int, f64p]),
    "b200mf_mg_get_level_info": (C.c_int, [vp, C.c_int, C.POINTER(MgLevelInfo)]),
    "b200mf_mg_prolongate": (C.c_int, [vp, C.c_int, vp, vp, vp]),
    "b200mf_mg_restrict_and_add": (C.c_int, [vp, C.c_int, vp, vp, vp]),
    "b200mf_mg_vcycle": (C.c_int, [vp, C.c_int, vp, vp, vp]),
    "b200mf_mg_cg_solve": (C.c_int, [vp, vp, C.POINTER(Operator), C.c_double, C.c_int, vp, vp,
                                     C.POINTER(SolverResult), vp]),
    "b200mf_mg_dist_cg_solve": (C.c_int, [vp, vp, vp, C.POINTER(Operator), C.c_double, C.c_int, vp, vp,
                                          C.POINTER(SolverResult), vp]),
    "b200mf_cg_solve_host": (C.c_int, [vp, C.POINTER(Operator), C.POINTER(SolverDesc), vp, vp,
                                       C.POINTER(SolverResult)]),
    "b200mf_mesh_create": (C.c_int, [C.POINTER(MeshDesc), C.POINTER(vp)]),
    "b200mf_mesh_create_partitioned": (C.c_int, [C.POINTER(PartitionDesc), C.POINTER(vp)]),
    "b200mf_mesh_create_adaptive": (C.c_int, [C.POINTER(AdaptiveDesc), C.POINTER(vp)]),
    "b200mf_mesh_adaptive_view_get": (C.c_int, [vp, C.POINTER(AdaptiveView)]),
    "b200mf_mesh_partition_view_get": (C.c_int, [vp, C.POINTER(PartitionView)]),
    "b200mf_cell_loop_range": (C.c_int, [vp, C.POINTER(Operator), vp, vp, u64, u64, vp]),
    "b200mf_cell_loop_range_dot": (C.c_int, [vp, C.POINTER(Operator), vp, vp, u64, u64, vp, vp]),
    "b200mf_copy_constrained_values_dot": (C.c_int, [vp, vp, vp, vp, vp]),
    "b200mf_debug_resolve_hanging_nodes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint16, C.c_int, vp]),
    "b200mf_vmult_host_batch": (C.c_int, [vp, C.POINTER(Operator), C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "b200mf_brick_probe": (C.c_int, [C.POINTER(SetupDesc), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
    "b200mf_bulk_probe": (C.c_int, [C.POINTER(SetupDesc), C.POINTER(BulkInfo)]),
    "b200mf_setup_enable_bulk": (C.c_int, [vp, C.c_int]),
    "b200mf_setup_select_brick_path": (C.c_int, [vp, C.c_int]),
    "b200mf_setup_enable_strided": (C.c_int, [vp, C.c_int]),
    "b200mf_setup_get_bulk_info": (C.c_int, [vp, C.POINTER(BulkInfo)]),
    "b200mf_vmult_prepare": (C.c_int, [vp, C.POINTER(Operator), vp, vp]),
    "b200mf_vmult_range": (C.c_int, [vp, C.POINTER(Operator), vp, vp, u64, u64, vp, vp]),
    "b200mf_cg_init": (C.c_int, [C.c_int, vp, vp, vp, vp, vp, u64, vp, vp]),
    "b200mf_cg_post": (C.c_int, [C.c_int, vp, vp, vp, u64, vp, C.c_int, vp]),
    "b200mf_cg_pre": (C.c_int, [C.c_int, vp, vp, vp, vp, u64, vp, C.c_int, vp]),
    "b200mf_cg_final": (C.c_int, [C.c_int, vp, vp, u64, vp, C.c_int, vp]),
    "b200mf_ghost_pack": (C.c_int, [C.c_int, vp, vp, vp, u64, vp]),
    "b200mf_ghost_unpack_add": (C.c_int, [C.c_int, vp, vp, vp, u64, vp]),
    "b200mf_mesh_view_get": (C.c_int, [vp, C.POINTER(MeshView)]),
    "b200mf_mesh_destroy": (C.c_int, [vp]),
    "b200mf_setup_create_from_mesh": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "b200mf_last_error": (C.c_char_p, []),
    "b200mf_version": (C.c_int, []),
    "b200mf_kernel_launch_count": (u64, []),
}

_lib = None


class B200MFError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"b200mf error {code}: {message}")
        self.code = code


def load():
    """Load libb200mf.so (once) and declare the prototypes of every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError => a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code):
    if code != OK:
        raise B200MFError(code, load().b200mf_last_error().decode())
    return code
