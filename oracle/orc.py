"""ctypes wrapper around oracle/liborc.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (mavmap_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from mavmap_b200 import _abi
from mavmap_b200._abi import (BAOptions, BASummary, MatchOptions, as_ptr, p_f32, p_f64, p_i32, p_u8)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liborc.so")
_LIB = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("orc_camera.c", "orc_geometry.c", "orc_match.c", "orc_ba.c", "orc.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_num_threads.restype = C.c_int
        L.orc_ba_options_default.argtypes = [C.POINTER(BAOptions)]
        L.orc_ba_solve.restype = C.c_int
        L.orc_ba_solve.argtypes = [C.POINTER(_abi.BAProblem), C.POINTER(BAOptions), C.POINTER(BASummary)]
        L.orc_rot_prior.restype = C.c_double
        L.orc_rot_prior.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
        L.orc_ba_cost.restype = C.c_double
        L.orc_ba_cost.argtypes = [C.POINTER(_abi.BAProblem), C.POINTER(BAOptions)]
        L.orc_pose_refine.restype = C.c_int
        L.orc_pose_refine.argtypes = _abi.PROTOTYPES["mm_pose_refine"][1]
        for n in ("camera_world2image", "camera_image2world", "camera_image2world_normalized"):
            f = getattr(L, "orc_" + n)
            f.restype = C.c_int
            f.argtypes = _abi.PROTOTYPES["mm_" + n][1]
        L.orc_triangulate_two_view.restype = C.c_int
        L.orc_triangulate_two_view.argtypes = _abi.PROTOTYPES["mm_triangulate_two_view"][1]
        L.orc_match_pair.restype = C.c_int
        L.orc_match_pair.argtypes = _abi.PROTOTYPES["mm_match_pair"][1]
        L.orc_knn2.restype = C.c_int
        L.orc_knn2.argtypes = [p_f32, C.c_int32, p_f32, C.c_int32, C.c_int32, p_f32, p_f32,
                               C.c_double, p_i32, p_f32]
        L.orc_ba_residual_jet.argtypes = [C.c_int, p_f64, p_f64, p_f64, p_f64, p_f64, p_f64]
        L.orc_ba_residual.argtypes = [C.c_int, p_f64, p_f64, p_f64, p_f64, p_f64]
        _LIB = L
    return _LIB


def num_threads():
    return lib().orc_num_threads()


def _chk(rc):
    if rc == _abi.MM_OK:
        return
    if rc in (_abi.MM_ERR_INVALID_ARG, _abi.MM_ERR_DATUM, _abi.MM_ERR_MIN_TRACK_LEN):
        raise ValueError("oracle: invalid argument (%d)" % rc)
    raise RuntimeError("oracle error %d" % rc)


# ---- camera models -----------------------------------------------------------
def _params9(params):
    p = np.zeros(9)
    p[:len(params)] = params
    return p


def camera_world2image(model, params, xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    uv = np.empty((len(xyz), 2))
    p = _params9(params)
    _chk(lib().orc_camera_world2image(model, as_ptr(p, p_f64), len(xyz), as_ptr(xyz, p_f64), as_ptr(uv, p_f64)))
    return uv


def camera_image2world(model, params, uv):
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    xyz = np.empty((len(uv), 3))
    p = _params9(params)
    _chk(lib().orc_camera_image2world(model, as_ptr(p, p_f64), len(uv), as_ptr(uv, p_f64), as_ptr(xyz, p_f64)))
    return xyz


def camera_image2world_normalized(model, params, uv):
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    xy = np.empty((len(uv), 2))
    p = _params9(params)
    _chk(lib().orc_camera_image2world_normalized(model, as_ptr(p, p_f64), len(uv), as_ptr(uv, p_f64), as_ptr(xy, p_f64)))
    return xy


def ba_residual_jet(model, pose6, X, intr9, obs):
    pose6 = np.ascontiguousarray(pose6, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
    intr9 = _params9(intr9); obs = np.ascontiguousarray(obs, dtype=np.float64)
    r = np.empty(2); J = np.empty((2, 18))
    lib().orc_ba_residual_jet(model, as_ptr(pose6, p_f64), as_ptr(X, p_f64), as_ptr(intr9, p_f64),
                              as_ptr(obs, p_f64), as_ptr(r, p_f64), as_ptr(J, p_f64))
    return r, J


# ---- triangulation -------------------------------------------------------------
def triangulate_two_view(P1, P2, x1, x2):
    P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(3, 4)
    P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(3, 4)
    x1 = np.ascontiguousarray(x1, dtype=np.float64).reshape(-1, 2)
    x2 = np.ascontiguousarray(x2, dtype=np.float64).reshape(-1, 2)
    n = len(x1)
    out = {k: np.empty(n) for k in ("reproj1", "reproj2", "depth1", "depth2", "angle")}
    X = np.empty((n, 3))
    _chk(lib().orc_triangulate_two_view(as_ptr(P1, p_f64), as_ptr(P2, p_f64), n, as_ptr(x1, p_f64),
                                        as_ptr(x2, p_f64), as_ptr(X, p_f64),
                                        *[as_ptr(out[k], p_f64) for k in ("reproj1", "reproj2", "depth1", "depth2", "angle")]))
    out["X"] = X
    return out


def tri_angles(P1, P2, X):
    """triangulation.cc:101-147 for given points"""
    P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(3, 4); P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(3, 4)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
    ang = np.empty(len(X))
    L = lib(); L.orc_tri_angles.restype = C.c_int; L.orc_tri_angles.argtypes = [p_f64, p_f64, C.c_int64, p_f64, p_f64]
    _chk(L.orc_tri_angles(as_ptr(P1, p_f64), as_ptr(P2, p_f64), len(X), as_ptr(X, p_f64), as_ptr(ang, p_f64)))
    return ang


# ---- matching --------------------------------------------------------------------
def match_options(ratio_test=True, max_ratio=0.6, max_distance=-1.0, impl=0):
    o = MatchOptions()
    o.ratio_test = int(bool(ratio_test)); o.max_ratio = float(max_ratio)
    o.max_distance = float(max_distance); o.impl = int(impl)
    return o


def match_pair(d1, d2, xy1=None, xy2=None, ratio_test=True, max_ratio=0.6, max_distance=-1.0):
    d1 = np.ascontiguousarray(d1, dtype=np.float32); d2 = np.ascontiguousarray(d2, dtype=np.float32)
    n1, n2 = len(d1), len(d2)
    k = d1.shape[1] if d1.ndim == 2 else d2.shape[1]
    if xy1 is not None:
        xy1 = np.ascontiguousarray(xy1, dtype=np.float32); xy2 = np.ascontiguousarray(xy2, dtype=np.float32)
    cap = max(min(n1, n2), 1)
    q = np.empty(cap, dtype=np.int32); t = np.empty(cap, dtype=np.int32); dist = np.empty(cap, dtype=np.float32)
    n_out = C.c_int32(0)
    o = match_options(ratio_test, max_ratio, max_distance)
    _chk(lib().orc_match_pair(as_ptr(d1, p_f32), n1, as_ptr(d2, p_f32), n2, k, as_ptr(xy1, p_f32),
                              as_ptr(xy2, p_f32), C.byref(o), as_ptr(q, p_i32), as_ptr(t, p_i32),
                              as_ptr(dist, p_f32), C.byref(n_out)))
    m = n_out.value
    return q[:m].copy(), t[:m].copy(), dist[:m].copy()


def match_pair_cv2(d1, d2, xy1=None, xy2=None, ratio_test=True, max_ratio=0.6, max_distance=-1.0):
    """feature.cc:52-133 executed with the library the reference delegates to (cv2.BFMatcher)."""
    import cv2
    d1 = np.ascontiguousarray(d1, dtype=np.float32); d2 = np.ascontiguousarray(d2, dtype=np.float32)
    n1, n2 = len(d1), len(d2)
    if n1 == 0 or n2 == 0:
        return (np.zeros(0, np.int32),) * 2 + (np.zeros(0, np.float32),)
    m12 = m21 = None
    if max_distance != -1:
        a = np.asarray(xy1, dtype=np.float32).astype(np.float64); b = np.asarray(xy2, dtype=np.float32).astype(np.float64)
        dd = (a[:, None, 0] - b[None, :, 0]) ** 2 + (a[:, None, 1] - b[None, :, 1]) ** 2
        m12 = (dd < max_distance * max_distance).astype(np.uint8)
        m21 = np.ascontiguousarray(m12.T)
    q, t, dist = [], [], []
    if ratio_test:
        bf = cv2.BFMatcher(cv2.NORM_L2, False)
        k12 = bf.knnMatch(d1, d2, 2, mask=m12) if m12 is not None else bf.knnMatch(d1, d2, 2)
        k21 = bf.knnMatch(d2, d1, 2, mask=m21) if m21 is not None else bf.knnMatch(d2, d1, 2)

        def ratio(rows):
            out = []
            for r in rows:
                r = list(r)
                if len(r) > 1 and float(np.float32(r[0].distance) / np.float32(r[1].distance)) > max_ratio:
                    r = []
                out.append(r)
            return out
        k12, k21 = ratio(k12), ratio(k21)
        for r in k12:
            if len(r) < 2:
                continue
            j = r[0].trainIdx
            r2 = k21[j]
            if len(r2) < 2:
                continue
            if r2[0].trainIdx == r[0].queryIdx and r2[0].queryIdx == r[0].trainIdx:
                q.append(r[0].queryIdx); t.append(r[0].trainIdx); dist.append(r[0].distance)
    elif max_distance == -1:
        for m in cv2.BFMatcher(cv2.NORM_L2, True).match(d1, d2):
            q.append(m.queryIdx); t.append(m.trainIdx); dist.append(m.distance)
    else:
        bf = cv2.BFMatcher(cv2.NORM_L2, False)
        a12 = bf.match(d1, d2, mask=m12); a21 = bf.match(d2, d1, mask=m21)
        for m in a12:
            if m.trainIdx < len(a21) and m.queryIdx == a21[m.trainIdx].trainIdx:
                q.append(m.queryIdx); t.append(m.trainIdx); dist.append(m.distance)
    return np.array(q, np.int32), np.array(t, np.int32), np.array(dist, np.float32)


# ---- bundle adjustment -------------------------------------------------------------
def default_options():
    o = BAOptions()
    lib().orc_ba_options_default(C.byref(o))
    return o


def ba_solve_fn(p, o, s):
    _chk(lib().orc_ba_solve(p, o, s))


def pose_refine_fn(*a):
    _chk(lib().orc_pose_refine(*a))


def solve_flat(flat, c_options):
    """Oracle counterpart of mavmap_b200.ba.solve_flat (in place)."""
    from mavmap_b200.ba import solve_flat as _sf
    return _sf(flat, c_options, ba_solve_fn)


def ba_cost(flat, c_options):
    cp = flat.to_c()
    return lib().orc_ba_cost(C.byref(cp), C.byref(c_options))


def ransac_score(kind, models, x, y, threshold):
    from mavmap_b200.geometry import ransac_score as _rs
    L = lib()
    L.orc_ransac_score.restype = C.c_int
    L.orc_ransac_score.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
    return _rs(kind, models, x, y, threshold, _fn=L.orc_ransac_score)


def rot_prior(rvec, rvec0, weight):
    """BARotationConstraintCostFunction (bundle_adjustment.cc:57-111): (residual, d residual / d rvec [3])"""
    a = np.ascontiguousarray(rvec, dtype=np.float64); b = np.ascontiguousarray(rvec0, dtype=np.float64); J = np.zeros(3)
    p = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    r = lib().orc_rot_prior(p(a), p(b), float(weight), p(J))
    return r, J


def bundle_adjustment(fm, free, fixed, fixed_x, options, point3D_errors, rotation_constraints=None, gcp_ids=()):
    from mavmap_b200.ba import bundle_adjustment as _ba
    return _ba(fm, free, fixed, fixed_x, options, point3D_errors, rotation_constraints, gcp_ids,
               _solve_fn=ba_solve_fn)


def pose_refinement(rvec, tvec, camera_params, points2D, points3D, inlier_mask, options):
    from mavmap_b200.ba import pose_refinement as _pr
    return _pr(rvec, tvec, camera_params, points2D, points3D, inlier_mask, options, _refine_fn=pose_refine_fn)
