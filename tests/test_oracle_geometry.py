"""CPU tests: the oracle restatement of the camera models and triangulation against the
reference's own test vectors (camera_models_test.cc, triangulation_test.cc) and against the
reference code itself (tests/golden/camera_ref.npz, made from oracle/_ref)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from mavmap_b200 import synthetic

CAMERA_SETS = [   # src/base3d/camera_models_test.cc:59-82
    (1, [651.123, 655.123, 386.123, 511.123]),
    (3, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001, 0]),
    (3, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001, 1]),
    (3, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001, 0.5]),
    (2, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001]),
]


@pytest.mark.parametrize("code,params", CAMERA_SETS)
def test_camera_round_trips_like_reference_test(orc, code, params):
    # test_model<CameraModel>() of camera_models_test.cc:16-55, same tolerances
    uv = orc.camera_world2image(code, params, [[0.5, 0.23, 1.0]])
    xyz = orc.camera_image2world(code, params, uv)[0]
    assert abs(xyz[0] / xyz[2] - 0.5) < 1e-5 and abs(xyz[1] / xyz[2] - 0.23) < 1e-5
    xyz = orc.camera_image2world(code, params, [[200.0, 100.0]])
    uv = orc.camera_world2image(code, params, xyz)[0]
    assert abs(uv[0] - 200) < 1e-1 and abs(uv[1] - 100) < 1e-1
    uv = orc.camera_world2image(code, params, [[0.0, 0.0, 1.0]])[0]
    assert abs(uv[0] - params[2]) < 1e-6 and abs(uv[1] - params[3]) < 1e-6
    xyz = orc.camera_image2world(code, params, [[params[2], params[3]]])[0]
    assert abs(xyz[0] / xyz[2]) < 1e-4 and abs(xyz[1] / xyz[2]) < 1e-4


def test_camera_oracle_equals_reference_code(orc):
    g = np.load(os.path.join(GOLDEN, "camera_ref.npz"))
    for k in range(int(g["n"])):
        code = int(g["code%d" % k]); prm = g["params%d" % k]
        np.testing.assert_array_equal(orc.camera_world2image(code, prm, g["xyz%d" % k]), g["uv%d" % k])
        np.testing.assert_array_equal(orc.camera_image2world(code, prm, g["uv_in%d" % k]), g["xyz_out%d" % k])
        np.testing.assert_array_equal(orc.camera_image2world_normalized(code, prm, g["uv_in%d" % k]), g["xy_norm%d" % k])


def test_camera_oracle_equals_live_reference_build(orc):
    import ctypes as C
    path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libref_camera.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built here")
    ref = C.CDLL(path)
    p = C.POINTER(C.c_double)
    rng = np.random.default_rng(3)
    for code, params in CAMERA_SETS:
        prm = np.zeros(9); prm[:len(params)] = params
        xyz = rng.uniform([-0.5, -0.5, 0.9], [0.5, 0.5, 2.0], (200, 3)); uv = np.empty((200, 2))
        ref.ref_world2image(code, prm.ctypes.data_as(p), C.c_long(200), xyz.ctypes.data_as(p), uv.ctypes.data_as(p))
        np.testing.assert_array_equal(orc.camera_world2image(code, prm, xyz), uv)


@pytest.mark.parametrize("model", [1, 2, 3])
def test_ba_residual_jets_match_finite_differences(orc, model):
    rng = np.random.default_rng(model)
    intr = np.zeros(9); intr[:len(synthetic.INTRINSICS[model])] = synthetic.INTRINSICS[model]
    for trial in range(5):
        pose = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.3, 3) + [0, 0, 2.0]])
        if trial == 4:
            pose[:3] = 0.0          # theta == 0: first-order branch of AngleAxisRotatePoint
        X = rng.uniform([-1, -1, 4], [1, 1, 8]); obs = rng.uniform(300, 700, 2)
        r, J = orc.ba_residual_jet(model, pose, X, intr, obs)
        x0 = np.concatenate([pose, X, intr])
        Jn = np.zeros((2, 18))
        for k in range(18):
            h = 1e-6 * max(1.0, abs(x0[k]))
            xp, xm = x0.copy(), x0.copy(); xp[k] += h; xm[k] -= h
            rp, _ = orc.ba_residual_jet(model, xp[:6], xp[6:9], xp[9:], obs)
            rm, _ = orc.ba_residual_jet(model, xm[:6], xm[6:9], xm[9:], obs)
            Jn[:, k] = (rp - rm) / (2 * h)
        Jn[:, 9 + len(synthetic.INTRINSICS[model]):] = 0
        if trial == 4:
            # at theta == 0 the Jet derivative is that of pt + w x pt; FD crosses into the Rodrigues branch (same to O(h))
            np.testing.assert_allclose(J[:, :3], Jn[:, :3], rtol=0, atol=1e-4 * np.abs(J).max())
            J[:, :3] = Jn[:, :3] = 0
        np.testing.assert_allclose(J, Jn, rtol=0, atol=2e-7 * np.abs(J).max())


def _rodrigues(rvec):
    th = np.linalg.norm(rvec)
    if th < np.finfo(float).eps:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def reference_triangulation_vectors():
    """The cases of test_triangulate_point (src/base3d/triangulation_test.cc:16-60)."""
    pts = np.array([[0, 0.1, 0.1], [0, 1, 3], [0, 1, 2], [0.01, 0.2, 3], [-1, 0.1, 1], [0.1, 0.1, 0.2]], dtype=float)
    P1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    cases = []
    rx = 0.0
    while rx < 1:
        tx = 0.0
        while tx < 10:
            # SimilarityTransform3D(1, rx, 0.2, 0.3, tx, 2, 3) (similarity_transform.cc:61-75) = [R(rvec) | t]
            P2 = np.hstack([_rodrigues(np.array([rx, 0.2, 0.3])), np.array([[tx], [2.0], [3.0]])])
            cases.append((P1, P2, pts))
            tx += 2
        rx += 0.2
    return cases


def test_triangulation_reference_vectors(orc):
    cases = reference_triangulation_vectors()
    assert len(cases) == 25
    for P1, P2, pts in cases:
        h = np.hstack([pts, np.ones((len(pts), 1))])
        a = (P1 @ h.T).T; b = (P2 @ h.T).T
        x1 = a[:, :2] / a[:, 2:]; x2 = b[:, :2] / b[:, 2:]
        out = orc.triangulate_two_view(P1, P2, x1, x2)
        assert np.linalg.norm(out["X"] - pts, axis=1).max() < 1e-10       # ASSERT_ALMOST_EQUAL(..., 1e-10)


def test_triangulation_matches_numpy_svd_and_filters(orc):
    rng = np.random.default_rng(5)
    P1 = np.hstack([_rodrigues(rng.normal(0, 0.1, 3)), rng.normal(0, 0.2, (3, 1))])
    P2 = np.hstack([_rodrigues(rng.normal(0, 0.2, 3)), np.array([[1.0], [0.1], [-0.1]])])
    X = rng.uniform([-2, -2, 4], [2, 2, 9], (300, 3))
    h = np.hstack([X, np.ones((300, 1))])
    a = (P1 @ h.T).T; b = (P2 @ h.T).T
    x1 = a[:, :2] / a[:, 2:] + rng.normal(0, 1e-3, (300, 2)); x2 = b[:, :2] / b[:, 2:] + rng.normal(0, 1e-3, (300, 2))
    out = orc.triangulate_two_view(P1, P2, x1, x2)
    for i in range(300):
        A = np.array([x1[i, 0] * P1[2] - P1[0], x1[i, 1] * P1[2] - P1[1], x1[i, 0] * P1[1] - x1[i, 1] * P1[0],
                      x2[i, 0] * P2[2] - P2[0], x2[i, 1] * P2[2] - P2[1], x2[i, 0] * P2[1] - x2[i, 1] * P2[0]])
        v = np.linalg.svd(A)[2][3]
        np.testing.assert_allclose(out["X"][i], v[:3] / v[3], rtol=1e-9, atol=1e-9)
    Xh = np.hstack([out["X"], np.ones((300, 1))])
    q = (P2 @ Xh.T).T
    np.testing.assert_allclose(out["reproj2"], np.linalg.norm(q[:, :2] / q[:, 2:] - x2, axis=1), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["depth2"], q[:, 2] * np.linalg.norm(P2[:, 2]), rtol=1e-12)
    C1 = -np.linalg.inv(P1[:, :3]) @ P1[:, 3]; C2 = -np.linalg.inv(P2[:, :3]) @ P2[:, 3]
    r1 = np.linalg.norm(out["X"] - C1, axis=1); r2 = np.linalg.norm(out["X"] - C2, axis=1)
    ang = np.arccos((r1 ** 2 + r2 ** 2 - np.linalg.norm(C1 - C2) ** 2) / (2 * r1 * r2))
    np.testing.assert_allclose(out["angle"], ang, rtol=1e-9)


def _ransac_case(kind, n=3000, h=40, seed=0):
    """correspondences consistent with model 0 (+ noise + 30 % outliers) and h-1 perturbed hypotheses"""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed + kind)
    if kind == 0:
        R = Rotation.from_rotvec([0.1, -0.2, 0.05]).as_matrix(); t = np.array([0.3, -0.1, 0.5])
        X = rng.uniform([-2, -2, 4], [2, 2, 9], (n, 3)); Xc = X @ R.T + t
        x = Xc[:, :2] / Xc[:, 2:] + rng.normal(0, 1e-3, (n, 2)); y = X
        models = [np.hstack([R, t[:, None]])] + [np.hstack([Rotation.from_rotvec(rng.normal(0, 0.01, 3)).as_matrix() @ R, (t + rng.normal(0, 0.02, 3))[:, None]]) for _ in range(h - 1)]
    elif kind == 1:
        H = np.array([[1.1, 0.02, 5.0], [-0.03, 0.95, -3.0], [1e-4, -2e-4, 1.0]])
        x = rng.uniform(0, 1000, (n, 2)); xh = np.hstack([x, np.ones((n, 1))]) @ H.T
        y = xh[:, :2] / xh[:, 2:] + rng.normal(0, 0.5, (n, 2))
        models = [H] + [H + rng.normal(0, 1e-3, (3, 3)) * np.array([[1, 1, 100], [1, 1, 100], [1e-4, 1e-4, 0]]) for _ in range(h - 1)]
    else:
        R = Rotation.from_rotvec([0.02, 0.1, -0.03]).as_matrix(); t = np.array([1.0, 0.1, 0.05])
        X = rng.uniform([-2, -2, 4], [2, 2, 9], (n, 3)); X2 = X @ R.T + t
        x = X[:, :2] / X[:, 2:] + rng.normal(0, 1e-3, (n, 2)); y = X2[:, :2] / X2[:, 2:] + rng.normal(0, 1e-3, (n, 2))
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]); E = tx @ R
        models = [E] + [E + rng.normal(0, 0.01, (3, 3)) for _ in range(h - 1)]
    out = rng.random(n) < 0.3
    y = y.copy(); y[out] += rng.normal(0, 50.0 if kind == 1 else (1.0 if kind == 0 else 0.3), (out.sum(), y.shape[1]))
    return np.array(models), x, y, (2.0 if kind == 1 else 4e-3)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ransac_score_oracle_matches_numpy(kind):
    """oracle restatement of util/estimation.cc:83-126 + the three residual functions against plain numpy"""
    from oracle import orc
    models, x, y, thr = _ransac_case(kind)
    o = orc.ransac_score(kind, models, x, y, thr)
    for h in (0, 7):
        M = models[h]
        if kind == 0:
            p = y @ M[:, :3].T + M[:, 3]; r = np.linalg.norm(p[:, :2] / p[:, 2:] - x, axis=1)
        elif kind == 1:
            p = np.hstack([x, np.ones((len(x), 1))]) @ M.T; r = np.linalg.norm(p[:, :2] / p[:, 2:] - y, axis=1)
        else:
            x1 = np.hstack([x, np.ones((len(x), 1))]); x2 = np.hstack([y, np.ones((len(y), 1))])
            Ex1 = x1 @ M.T; Etx2 = x2 @ M; r = np.abs(np.sum(x2 * Ex1, axis=1) / np.sqrt(Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2))
        inl = r <= thr
        assert abs(int(inl.sum()) - int(o["num_inliers"][h])) <= 1            # a residual within an ulp of the threshold may flip
        assert abs(r[inl].sum() - o["residual_sum"][h]) < 1e-6 * max(1.0, r[inl].sum())
    assert o["best"] == 0 and o["inlier_mask"].sum() == o["num_inliers"][0] > 0.6 * len(x)


def test_ransac_score_oracle_reproduces_golden():
    import json, os, sys
    from conftest import GOLDEN
    from oracle import orc
    gold = json.load(open(os.path.join(GOLDEN, "extra.json")))["ransac"]
    for kind in (0, 1, 2):
        models, x, y, thr = _ransac_case(kind, n=2000, h=16, seed=21)
        r = orc.ransac_score(kind, models, x, y, thr)
        assert r["num_inliers"].tolist() == gold[str(kind)]["num_inliers"] and r["best"] == gold[str(kind)]["best"]
        np.testing.assert_allclose(r["residual_sum"], gold[str(kind)]["residual_sum"], rtol=1e-12)


def _ransac_homography(score_fn, seed=5, n=1000, n_out=400, trials=60, thr=1.0):
    """RANSAC as util/estimation.cc:26-146 runs it, with the SCORING step delegated to `score_fn` (all hypotheses of a batch at
    once): the data layout of estimation_test.cc:19-66 (first 400 of 1000 correspondences are gross outliers), for a homography."""
    rng = np.random.default_rng(seed)
    H = np.array([[1.05, 0.03, 12.0], [-0.02, 0.97, -7.0], [2e-5, -1e-5, 1.0]])
    src = np.stack([np.arange(n, dtype=float), np.sqrt(np.arange(n)) * 20 + 2], axis=1)
    d = np.hstack([src, np.ones((n, 1))]) @ H.T; dst = d[:, :2] / d[:, 2:]
    dst[:n_out] = rng.uniform(-6000, -2000, (n_out, 2))
    models = []
    for _ in range(trials):                       # 4-point DLT hypotheses on the host
        idx = rng.choice(n, 4, replace=False)
        A = []
        for (x, y), (u, v) in zip(src[idx], dst[idx]):
            A.append([-x, -y, -1, 0, 0, 0, u * x, u * y, u]); A.append([0, 0, 0, -x, -y, -1, v * x, v * y, v])
        h = np.linalg.svd(np.array(A))[2][-1]
        models.append((h / h[8]).reshape(3, 3) if abs(h[8]) > 1e-12 else np.eye(3))
    r = score_fn(1, np.array(models), src, dst, thr)
    return r, models, H, n, n_out


def test_ransac_loop_separates_outliers_like_estimation_test():
    """mirror of estimation_test.cc:19-66 (400 gross outliers of 1000 must be rejected exactly) with the oracle scorer"""
    from oracle import orc
    r, models, H, n, n_out = _ransac_homography(orc.ransac_score)
    assert r["num_inliers"][r["best"]] == n - n_out
    assert not r["inlier_mask"][:n_out].any() and r["inlier_mask"][n_out:].all()
    assert np.abs(models[r["best"]] - H).max() < 1e-6 * np.abs(H).max()
