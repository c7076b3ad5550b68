"""ctypes mirror of include/mavmap_b200.h (structs, constants, prototypes).

Pure declarations: no library is loaded here, so both the product loader
(mavmap_b200._lib) and the test-only oracle wrapper (oracle/orc.py) can share it.
"""
import ctypes as C

MM_OK = 0
MM_ERR_INVALID_ARG = 1
MM_ERR_DATUM = 2
MM_ERR_MIN_TRACK_LEN = 3
MM_ERR_NO_DEVICE = 4
MM_ERR_CUDA = 5
MM_ERR_ALLOC = 6
MM_ERR_NUMERICAL = 7
MM_ERR_UNSUPPORTED = 8

MM_MODEL_PINHOLE, MM_MODEL_OPENCV, MM_MODEL_CATA = 1, 2, 3
MM_INTR_STRIDE = 9
MODEL_NUM_PARAMS = {1: 4, 2: 8, 3: 9}

MM_MATCH_IMPL_AUTO, MM_MATCH_IMPL_SIMT, MM_MATCH_IMPL_TCGEN05 = 0, 1, 2
MM_POSE_FREE, MM_POSE_FIXED, MM_POSE_FIXED_X = 0, 1, 2
MM_LOSS_TRIVIAL, MM_LOSS_CAUCHY = 0, 1
MM_SOLVER_PCG, MM_SOLVER_CHOLESKY = 0, 1
MM_PRECOND_AUTO, MM_PRECOND_TWO_LEVEL, MM_PRECOND_TILE_CHOLESKY = 0, 1, 2
MM_BA_TRACE_MAX = 512
TERMINATION = {0: "NO_CONVERGENCE", 1: "FUNCTION_TOLERANCE", 2: "GRADIENT_TOLERANCE",
               3: "PARAMETER_TOLERANCE", 4: "NUMERICAL_FAILURE", 5: "EMPTY"}

p_f64 = C.POINTER(C.c_double)
p_f32 = C.POINTER(C.c_float)
p_i32 = C.POINTER(C.c_int32)
p_i64 = C.POINTER(C.c_int64)
p_u8 = C.POINTER(C.c_uint8)


class MatchOptions(C.Structure):
    _fields_ = [("ratio_test", C.c_int32), ("max_ratio", C.c_double),
                ("max_distance", C.c_double), ("impl", C.c_int32)]


class BAProblem(C.Structure):
    _fields_ = [("n_img", C.c_int32), ("n_cam", C.c_int32), ("n_pt", C.c_int32),
                ("n_obs", C.c_int64),
                ("poses", p_f64), ("pose_const", p_u8), ("img_cam", p_i32),
                ("intr", p_f64), ("cam_model", p_i32), ("intr_const", p_u8),
                ("pts", p_f64), ("pt_const", p_u8),
                ("obs_xy", p_f64), ("obs_img", p_i32), ("obs_pt", p_i32),
                ("pt_err", p_f64), ("rot_prior", p_f64), ("rot_prior_w", p_f64)]


class BAOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("loss_type", C.c_int32),
                ("loss_scale", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("jacobi_scaling", C.c_int32),
                ("max_num_consecutive_invalid_steps", C.c_int32),
                ("linear_solver", C.c_int32), ("pcg_tolerance", C.c_double),
                ("pcg_max_iterations", C.c_int32), ("print_progress", C.c_int32),
                ("pcg_preconditioner", C.c_int32), ("tile_cholesky_tolerance", C.c_double)]


class BASummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("num_residuals", C.c_int64),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("termination", C.c_int32), ("num_iterations", C.c_int32),
                ("return_value", C.c_double),
                ("trace_cost", C.c_double * MM_BA_TRACE_MAX),
                ("trace_radius", C.c_double * MM_BA_TRACE_MAX),
                ("trace_gradient_max_norm", C.c_double * MM_BA_TRACE_MAX),
                ("trace_accepted", C.c_int32 * MM_BA_TRACE_MAX),
                ("trace_linear_iterations", C.c_int32 * MM_BA_TRACE_MAX),
                ("ms_setup", C.c_double), ("ms_linearize", C.c_double),
                ("ms_schur", C.c_double), ("ms_pcg", C.c_double),
                ("ms_update", C.c_double), ("ms_total", C.c_double)]

    def as_dict(self):
        n = min(self.num_iterations, MM_BA_TRACE_MAX)
        return {
            "initial_cost": self.initial_cost, "final_cost": self.final_cost,
            "num_residuals": self.num_residuals,
            "num_successful_steps": self.num_successful_steps,
            "num_unsuccessful_steps": self.num_unsuccessful_steps,
            "termination": TERMINATION.get(self.termination, str(self.termination)),
            "num_iterations": self.num_iterations, "return_value": self.return_value,
            "trace_cost": list(self.trace_cost[:n]), "trace_radius": list(self.trace_radius[:n]),
            "trace_gradient_max_norm": list(self.trace_gradient_max_norm[:n]),
            "trace_accepted": list(self.trace_accepted[:n]),
            "trace_linear_iterations": list(self.trace_linear_iterations[:n]),
            "ms": {"setup": self.ms_setup, "linearize": self.ms_linearize, "schur": self.ms_schur,
                   "pcg": self.ms_pcg, "update": self.ms_update, "total": self.ms_total},
        }


# symbol -> (restype, argtypes) for every entry point include/mavmap_b200.h declares
# int (*mm_allreduce_fn)(void* user, double* buf, int64_t count, void* stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

PROTOTYPES = {
    "mm_abi_version": (C.c_int, []),
    "mm_last_error": (C.c_char_p, []),
    "mm_device_count": (C.c_int, []),
    "mm_kernel_launch_count": (C.c_uint64, []),
    "mm_match_pair_counters": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mm_camera_model_name_to_code": (C.c_int, [C.c_char_p]),
    "mm_camera_model_num_params": (C.c_int, [C.c_int]),
    "mm_camera_image2world_threshold": (C.c_double, [C.c_double, C.c_int, p_f64]),
    "mm_camera_world2image": (C.c_int, [C.c_int, p_f64, C.c_int64, p_f64, p_f64]),
    "mm_camera_image2world": (C.c_int, [C.c_int, p_f64, C.c_int64, p_f64, p_f64]),
    "mm_camera_image2world_normalized": (C.c_int, [C.c_int, p_f64, C.c_int64, p_f64, p_f64]),
    "mm_triangulate_two_view": (C.c_int, [p_f64, p_f64, C.c_int64, p_f64, p_f64, p_f64,
                                          p_f64, p_f64, p_f64, p_f64, p_f64]),
    "mm_reproj_errors": (C.c_int, [p_f64, C.c_int64, p_f64, p_f64, p_f64, p_f64]),
    "mm_tri_angles": (C.c_int, [p_f64, p_f64, C.c_int64, p_f64, p_f64]),
    "mm_match_options_default": (None, [C.POINTER(MatchOptions)]),
    "mm_match_pair": (C.c_int, [p_f32, C.c_int32, p_f32, C.c_int32, C.c_int32, p_f32, p_f32,
                                C.POINTER(MatchOptions), p_i32, p_i32, p_f32, p_i32]),
    "mm_match_set_create": (C.c_int, [p_f32, p_f32, p_i32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_void_p)]),
    "mm_match_set_create_dev": (C.c_int, [C.c_void_p, C.c_void_p, p_i32, C.c_int32, C.c_int32,
                                          C.POINTER(C.c_void_p)]),
    "mm_match_set_destroy": (None, [C.c_void_p]),
    "mm_match_set_pairs": (C.c_int, [C.c_void_p, p_i32, p_i32, C.c_int32,
                                     C.POINTER(MatchOptions), p_i64, p_i32, p_i32, p_f32,
                                     C.c_int64]),
    "mm_match_set_pairs_dev": (C.c_int, [C.c_void_p, p_i32, p_i32, C.c_int32,
                                         C.POINTER(MatchOptions), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_ransac_score": (C.c_int, [C.c_int32, p_f64, C.c_int32, C.c_int64, p_f64, p_f64, C.c_double, p_i32, p_f64, p_i32, p_f64, p_u8]),
    "mm_feature_cache_info": (C.c_int, [C.c_char_p, C.c_char_p, p_i32, p_i32, p_i32, p_i32]),
    "mm_feature_cache_read": (C.c_int, [C.c_char_p, C.c_char_p, p_f32, p_f32, C.c_int32, C.c_int32]),
    "mm_match_set_create_from_cache": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_void_p)]),
    "mm_ba_options_default": (None, [C.POINTER(BAOptions)]),
    "mm_ba_solve": (C.c_int, [C.POINTER(BAProblem), C.POINTER(BAOptions), C.POINTER(BASummary)]),
    "mm_ba_session_create": (C.c_int, [C.POINTER(BAProblem), C.POINTER(BAOptions), C.c_void_p,
                                       C.POINTER(C.c_void_p)]),
    "mm_ba_session_create_sharded": (C.c_int, [C.POINTER(BAProblem), C.POINTER(BAOptions), C.c_void_p, C.c_int32, C.c_int32,
                                               ALLREDUCE_FN, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mm_ba_session_reset": (C.c_int, [C.c_void_p]),
    "mm_ba_session_iterate": (C.c_int, [C.c_void_p, C.c_int32, p_i32]),
    "mm_ba_session_download": (C.c_int, [C.c_void_p, p_f64, p_f64, p_f64, p_f64]),
    "mm_ba_session_summary": (C.c_int, [C.c_void_p, C.POINTER(BASummary)]),
    "mm_ba_session_time_kernel": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, p_f64]),
    "mm_ba_session_num_blocks": (C.c_int64, [C.c_void_p]),
    "mm_ba_session_coarse_dim": (C.c_int32, [C.c_void_p]),
    "mm_ba_session_solver_info": (C.c_int, [C.c_void_p, p_f64]),
    "mm_debug_spd_inverse": (C.c_int, [p_f64, C.c_int32]),
    "mm_debug_tilechol_plan_create": (C.c_int, [C.c_int32, C.c_int32, p_i32, p_i32, p_f64, C.c_int32, C.POINTER(C.c_void_p)]),
    "mm_debug_tilechol_plan_array": (C.c_int64, [C.c_void_p, C.c_int32, p_i64, C.c_int64]),
    "mm_debug_tilechol_plan_destroy": (None, [C.c_void_p]),
    "mm_debug_tilechol_solve": (C.c_int, [C.c_int32, C.c_int32, p_i32, p_i32, p_f64, p_f64, C.c_int32, p_f64, p_f64, p_f64, p_f64, C.c_int32, p_f64, p_f64]),
    "mm_ba_session_destroy": (None, [C.c_void_p]),
    "mm_pose_refine": (C.c_int, [p_f64, p_f64, C.c_int, p_f64, C.c_int64, p_f64, p_f64, p_u8,
                                 C.POINTER(BAOptions), C.POINTER(BASummary), p_f64]),
    "mm_pose_refine_batch": (C.c_int, [C.c_int32, p_f64, p_f64, p_i32, p_f64, p_i64, p_f64, p_f64,
                                       C.POINTER(BAOptions), C.POINTER(BASummary), p_f64]),
}


def bind(lib, prefix_from="mm_", prefix_to="mm_", names=None):
    """Attach restype/argtypes to `lib` for the given symbols (renaming the prefix)."""
    for name, (res, args) in PROTOTYPES.items():
        if names is not None and name not in names:
            continue
        sym = prefix_to + name[len(prefix_from):]
        fn = getattr(lib, sym)
        fn.restype = res
        fn.argtypes = args
    return lib


def as_ptr(arr, ctype_ptr):
    """numpy array -> ctypes pointer (None passes through)."""
    if arr is None:
        return None
    return arr.ctypes.data_as(ctype_ptr)
