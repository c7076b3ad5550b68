"""GPU parity tests (through the C ABI): camera models and two-view triangulation."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_geometry import CAMERA_SETS, reference_triangulation_vectors

pytestmark = pytest.mark.gpu


def test_camera_models_equal_reference_code(mm):
    g = np.load(os.path.join(GOLDEN, "camera_ref.npz"))
    for k in range(int(g["n"])):
        code = int(g["code%d" % k]); prm = g["params%d" % k]
        # same operations in the same order on IEEE doubles; FMA contraction on the device may
        # differ in the last ulp, hence 4 ulp instead of bit equality
        np.testing.assert_allclose(mm.camera_model_world2image(g["xyz%d" % k], code, prm), g["uv%d" % k], rtol=1e-15 * 4, atol=1e-12)
        np.testing.assert_allclose(mm.camera_model_image2world(g["uv_in%d" % k], code, prm), g["xyz_out%d" % k], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(mm.camera_model_image2world(g["uv_in%d" % k], code, prm, normalized=True), g["xy_norm%d" % k], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("code,params", CAMERA_SETS)
def test_camera_round_trips_like_reference_test(mm, code, params):
    uv = mm.camera_model_world2image([[0.5, 0.23, 1.0]], code, params)
    xyz = mm.camera_model_image2world(uv, code, params)[0]
    assert abs(xyz[0] / xyz[2] - 0.5) < 1e-5 and abs(xyz[1] / xyz[2] - 0.23) < 1e-5
    uv = mm.camera_model_world2image([[0.0, 0.0, 1.0]], code, params)[0]
    assert abs(uv[0] - params[2]) < 1e-6 and abs(uv[1] - params[3]) < 1e-6


def test_camera_large_batch_vs_oracle(mm, orc):
    rng = np.random.default_rng(1)
    xyz = rng.uniform([-0.5, -0.5, 0.8], [0.5, 0.5, 3.0], (200003, 3))
    for code, params in CAMERA_SETS:
        np.testing.assert_allclose(mm.camera_model_world2image(xyz, code, params), orc.camera_world2image(code, params, xyz), rtol=1e-14, atol=1e-11)
    assert mm.camera_model_world2image(np.zeros((0, 3)), 1, CAMERA_SETS[0][1]).shape == (0, 2)
    with pytest.raises(ValueError):
        from mavmap_b200._lib import check, lib
        from mavmap_b200._abi import as_ptr, p_f64
        p = np.zeros(9); check(lib().mm_camera_world2image(7, as_ptr(p, p_f64), 1, as_ptr(xyz, p_f64), as_ptr(xyz, p_f64)))


def test_triangulation_reference_vectors(mm):
    for P1, P2, pts in reference_triangulation_vectors():
        h = np.hstack([pts, np.ones((len(pts), 1))])
        a = (P1 @ h.T).T; b = (P2 @ h.T).T
        X = mm.triangulate_points(P1, P2, a[:, :2] / a[:, 2:], b[:, :2] / b[:, 2:])
        assert np.linalg.norm(X - pts, axis=1).max() < 1e-10        # triangulation_test.cc tolerance


def test_triangulation_batch_vs_oracle(mm, orc):
    rng = np.random.default_rng(11)
    P1 = np.hstack([np.eye(3), np.zeros((3, 1))]); P2 = np.hstack([np.eye(3), np.array([[-1.0], [0.05], [0.02]])])
    X = rng.uniform([-3, -3, 4], [3, 3, 20], (5000, 3))
    x1 = X[:, :2] / X[:, 2:] + rng.normal(0, 1e-3, (5000, 2))
    Xc = X + P2[:, 3]; x2 = Xc[:, :2] / Xc[:, 2:] + rng.normal(0, 1e-3, (5000, 2))
    g = mm.triangulate_two_view(P1, P2, x1, x2); o = orc.triangulate_two_view(P1, P2, x1, x2)
    np.testing.assert_allclose(g["X"], o["X"], rtol=1e-9, atol=1e-9)
    for k in ("reproj1", "reproj2", "depth1", "depth2", "angle"):
        np.testing.assert_allclose(g[k], o[k], rtol=1e-8, atol=1e-10)
    assert mm.triangulate_points(P1, P2, np.zeros((0, 2)), np.zeros((0, 2))).shape == (0, 3)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ransac_score_bit_exact_against_oracle(mm, orc, kind):
    """hypothesis scoring on the device; residuals are computed without FMA contraction in the reference's
    operation order, so residuals, inlier masks and counts are identical to the oracle and the same model wins."""
    from test_oracle_geometry import _ransac_case
    from mavmap_b200.geometry import ransac_score
    models, x, y, thr = _ransac_case(kind, n=20000, h=64, seed=3)
    g = ransac_score(kind, models, x, y, thr); o = orc.ransac_score(kind, models, x, y, thr)
    np.testing.assert_array_equal(g["num_inliers"], o["num_inliers"])
    np.testing.assert_allclose(g["residual_sum"], o["residual_sum"], rtol=1e-12)
    assert g["best"] == o["best"] == 0
    np.testing.assert_array_equal(g["residuals"], o["residuals"])
    np.testing.assert_array_equal(g["inlier_mask"], o["inlier_mask"])
    e = ransac_score(kind, models[:0], x, y, thr)
    assert e["best"] == -1



def test_tri_angles_of_given_points_vs_oracle(mm, orc):
    """calc_tri_angles (triangulation.cc:101-147) takes the 3-D points as given: device vs oracle vs the closed form, including
    points on a camera's principal plane and (numerically) at infinity, and a non-rotation projection matrix"""
    rng = np.random.default_rng(21)
    from mavmap_b200.synthetic import _rodrigues
    R2 = _rodrigues(np.array([0.1, -0.2, 0.05]))[0]
    P1 = np.hstack([np.eye(3), np.zeros((3, 1))]); P2 = np.hstack([R2, np.array([[-1.5], [0.2], [0.1]])])
    X = rng.uniform([-3, -3, 2], [3, 3, 30], (500, 3))
    X[0] = [0.3, -0.4, 0.0]; X[1] = [1e8, 2e8, 5e9]; X[2] = -R2.T @ P2[:, 3]            # principal plane, far away, ON the second centre
    a = mm.calc_tri_angles(P1, P2, X)
    # acos amplifies the last bit of its argument near +-1 (rays almost parallel): sqrt(2 ulp) ~ 2e-8 there, 1e-15 elsewhere
    o = orc.tri_angles(P1, P2, X)
    np.testing.assert_allclose(a, o, atol=5e-8); np.testing.assert_allclose(a[o > 1e-3], o[o > 1e-3], rtol=1e-12)
    c1 = np.zeros(3); c2 = -R2.T @ P2[:, 3]
    r1 = np.linalg.norm(X - c1, axis=1); r2 = np.linalg.norm(X - c2, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = np.arccos((r1 ** 2 + r2 ** 2 - np.sum((c1 - c2) ** 2)) / (2 * r1 * r2))
    ref[np.isnan(ref)] = 0.0
    np.testing.assert_allclose(a[3:], ref[3:], atol=1e-9)
    assert a[0] > 0.1 and a[1] < 1e-6 and a[2] < 1e-6
    # consistent with the fused kernel on points that it triangulated itself
    x1 = X[3:, :2] / X[3:, 2:]; Xc = X[3:] @ R2.T + P2[:, 3]; x2 = Xc[:, :2] / Xc[:, 2:]
    tri = mm.triangulate_two_view(P1, P2, x1, x2)
    np.testing.assert_allclose(mm.calc_tri_angles(P1, P2, tri["X"]), tri["angle"], atol=1e-9)
