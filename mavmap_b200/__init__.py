"""mavmap_b200 — B200-native bundle adjustment, descriptor matching, triangulation and camera
models behind the MAVMAP (mavmap/mavmap) call surface.

Host-side mirror of the reference interface for the hot path only; every function marshals
into the C ABI of libmavmap_b200.so (include/mavmap_b200.h), hand-written sm_100a CUDA.
There is no CPU fallback: compute calls raise if the library or a CUDA device is missing.
"""
from .ba import (BA_POSE_FIXED, BA_POSE_FIXED_X, BA_POSE_FREE, BundleAdjustmentOptions, FlatProblem,
                 bundle_adjustment, pose_refinement, pose_refinement_batch)
from .fm import FeatureManager
from .geometry import (CAMERA_MODEL_NAME_TO_CODE, calc_tri_angles, camera_model_image2world,
                       camera_model_image2world_threshold, camera_model_name_to_code,
                       camera_model_world2image, triangulate_points, triangulate_two_view)
from .matching import MatchSet, match_brute_force

__all__ = [
    "BA_POSE_FREE", "BA_POSE_FIXED", "BA_POSE_FIXED_X", "BundleAdjustmentOptions", "FlatProblem",
    "bundle_adjustment", "pose_refinement", "pose_refinement_batch", "FeatureManager", "CAMERA_MODEL_NAME_TO_CODE",
    "camera_model_name_to_code", "camera_model_world2image", "camera_model_image2world",
    "camera_model_image2world_threshold", "triangulate_points", "triangulate_two_view",
    "calc_tri_angles", "match_brute_force", "MatchSet",
]
