"""Host-side mirror of the reference geometry helpers on the hot path.

  camera_model_name_to_code / world2image / image2world / image2world_threshold
      src/base3d/camera_models.h:375-423, camera_models.cc:12-52
  triangulate_points, calc_tri_angles      src/base3d/triangulation.cc:53-147
  calc_reproj_errors, calc_depth           src/base3d/projection.cc:107-149
All arithmetic runs in the CUDA library through the C ABI (include/mavmap_b200.h).
"""
import numpy as np

from ._abi import MODEL_NUM_PARAMS, as_ptr, p_f64
from ._lib import check, lib

CAMERA_MODEL_NAME_TO_CODE = {"PINHOLE": 1, "OPENCV": 2, "CATA": 3}


def camera_model_name_to_code(name):
    return lib().mm_camera_model_name_to_code(name.encode())


def _params(model_code, params):
    n = MODEL_NUM_PARAMS[int(model_code)]
    p = np.zeros(9)
    p[:n] = np.asarray(params, dtype=np.float64)[:n]
    return p


def camera_model_world2image(xyz, model_code, params):
    """xyz [n,3] camera-frame points -> uv [n,2] pixels."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    uv = np.empty((len(xyz), 2))
    p = _params(model_code, params)
    check(lib().mm_camera_world2image(int(model_code), as_ptr(p, p_f64), len(xyz), as_ptr(xyz, p_f64), as_ptr(uv, p_f64)))
    return uv


def camera_model_image2world(uv, model_code, params, normalized=False):
    """uv [n,2] -> xyz [n,3] (scalar overload) or, normalized=True, (x/z, y/z) [n,2]
    (vector overload, camera_models.cc:24-44)."""
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    p = _params(model_code, params)
    if normalized:
        out = np.empty((len(uv), 2))
        check(lib().mm_camera_image2world_normalized(int(model_code), as_ptr(p, p_f64), len(uv), as_ptr(uv, p_f64), as_ptr(out, p_f64)))
    else:
        out = np.empty((len(uv), 3))
        check(lib().mm_camera_image2world(int(model_code), as_ptr(p, p_f64), len(uv), as_ptr(uv, p_f64), as_ptr(out, p_f64)))
    return out


def camera_model_image2world_threshold(threshold, model_code, params):
    p = _params(model_code, params)
    return lib().mm_camera_image2world_threshold(float(threshold), int(model_code), as_ptr(p, p_f64))


def triangulate_two_view(proj_matrix1, proj_matrix2, points1, points2):
    """Fused triangulate_points + calc_reproj_errors (both views) + calc_depth (both views) +
    calc_tri_angles, as SequentialMapper::process evaluates them (sequential_mapper.cc:786-801).
    Returns a dict with X [n,3], reproj1, reproj2, depth1, depth2, angle [n]."""
    P1 = np.ascontiguousarray(proj_matrix1, dtype=np.float64).reshape(3, 4)
    P2 = np.ascontiguousarray(proj_matrix2, dtype=np.float64).reshape(3, 4)
    x1 = np.ascontiguousarray(points1, dtype=np.float64).reshape(-1, 2)
    x2 = np.ascontiguousarray(points2, dtype=np.float64).reshape(-1, 2)
    n = len(x1)
    keys = ("reproj1", "reproj2", "depth1", "depth2", "angle")
    out = {k: np.empty(n) for k in keys}
    X = np.empty((n, 3))
    check(lib().mm_triangulate_two_view(as_ptr(P1, p_f64), as_ptr(P2, p_f64), n, as_ptr(x1, p_f64), as_ptr(x2, p_f64),
                                        as_ptr(X, p_f64), *[as_ptr(out[k], p_f64) for k in keys]))
    out["X"] = X
    return out


def triangulate_points(proj_matrix1, proj_matrix2, points1, points2):
    return triangulate_two_view(proj_matrix1, proj_matrix2, points1, points2)["X"]


def calc_tri_angles(proj_matrix1, proj_matrix2, points3D):
    """calc_tri_angles (triangulation.cc:101-147): angle between the two rays of each GIVEN 3-D point."""
    P1 = np.ascontiguousarray(proj_matrix1, dtype=np.float64).reshape(3, 4)
    P2 = np.ascontiguousarray(proj_matrix2, dtype=np.float64).reshape(3, 4)
    X = np.ascontiguousarray(points3D, dtype=np.float64).reshape(-1, 3)
    ang = np.empty(len(X))
    check(lib().mm_tri_angles(as_ptr(P1, p_f64), as_ptr(P2, p_f64), len(X), as_ptr(X, p_f64), as_ptr(ang, p_f64)))
    return ang


def calc_reproj_errors(points2D, points3D, proj_matrix):
    """calc_reproj_errors (projection.cc:107-130): ||pi(P X) - x|| on the normalised plane."""
    P = np.ascontiguousarray(proj_matrix, dtype=np.float64).reshape(3, 4)
    x = np.ascontiguousarray(points2D, dtype=np.float64).reshape(-1, 2)
    X = np.ascontiguousarray(points3D, dtype=np.float64).reshape(-1, 3)
    err = np.empty(len(X))
    check(lib().mm_reproj_errors(as_ptr(P, p_f64), len(X), as_ptr(x, p_f64), as_ptr(X, p_f64), as_ptr(err, p_f64), None))
    return err


def calc_depth(proj_matrix, points3D):
    """calc_depth (projection.cc:133-149), batched over points."""
    P = np.ascontiguousarray(proj_matrix, dtype=np.float64).reshape(3, 4)
    X = np.ascontiguousarray(points3D, dtype=np.float64).reshape(-1, 3)
    d = np.empty(len(X))
    check(lib().mm_reproj_errors(as_ptr(P, p_f64), len(X), None, as_ptr(X, p_f64), None, as_ptr(d, p_f64)))
    return d


RANSAC_P3P, RANSAC_HOMOGRAPHY, RANSAC_ESSENTIAL = 0, 1, 2


def ransac_score(kind, models, x, y, threshold, _fn=None):
    """Scores RANSAC hypotheses as util/estimation.cc:83-126 does (inliers |r| <= threshold, sum of |r| over them) with the
    residuals of p3p.cc:172-199 (kind 0: x = points2D, y = points3D, models [h,3,4]), projective_transform.cc:48-74 (kind 1:
    src, dst, models [h,3,3]) or essential_matrix.cc:131-162 (kind 2).  Returns a dict: num_inliers [h], residual_sum [h],
    best (index by the reference's rule), residuals [n] and inlier_mask [n] of the best model."""
    import ctypes as C
    from ._abi import p_i32, p_u8
    msize = 12 if kind == RANSAC_P3P else 9
    M = np.ascontiguousarray(models, dtype=np.float64).reshape(-1, msize)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 2)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, 3 if kind == RANSAC_P3P else 2)
    assert len(x) == len(y)
    cnt = np.zeros(len(M), np.int32); rs = np.zeros(len(M)); best = C.c_int32(-1)
    res = np.zeros(len(x)); mask = np.zeros(len(x), np.uint8)
    fn = _fn or lib().mm_ransac_score
    rc = fn(int(kind), as_ptr(M, p_f64), len(M), len(x), as_ptr(x, p_f64), as_ptr(y, p_f64), float(threshold),
            as_ptr(cnt, p_i32), as_ptr(rs, p_f64), C.byref(best), as_ptr(res, p_f64), as_ptr(mask, p_u8))
    if _fn is None:
        check(rc)
    return {"num_inliers": cnt, "residual_sum": rs, "best": best.value, "residuals": res, "inlier_mask": mask.astype(bool)}
