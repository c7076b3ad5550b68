"""CPU tests: the BA oracle (Ceres-1.8-semantics LM + Schur) and the host-side flattening of
the FeatureManager (bundle_adjustment.cc:228-387, 459-471)."""
import json
import os
import sys

import numpy as np
import pytest

import mavmap_b200 as mm
from conftest import GOLDEN
from mavmap_b200 import synthetic
from mavmap_b200.ba import flatten


def _opts(orc, iters=8):
    o = orc.default_options(); o.max_num_iterations = iters; o.function_tolerance = 0; o.gradient_tolerance = 0
    return o


def test_oracle_reproduces_committed_traces(orc):
    with open(os.path.join(GOLDEN, "ba_trace.json")) as f:
        gold = json.load(f)
    for name, kw, model, refine in [("tiny_pinhole", synthetic.BA_CONFIGS["tiny"], 1, False),
                                    ("tiny_opencv", dict(synthetic.BA_CONFIGS["tiny"], seed=77), 2, False),
                                    ("tiny_cata_refine", dict(synthetic.BA_CONFIGS["tiny"], seed=78), 3, True)]:
        flat, _ = synthetic.make_ba_problem(model=model, refine_camera_params=refine, **kw)
        s = orc.solve_flat(flat, _opts(orc)).as_dict()
        np.testing.assert_allclose(s["trace_cost"], gold[name]["trace_cost"], rtol=1e-9)
        np.testing.assert_allclose(s["trace_radius"], gold[name]["trace_radius"], rtol=1e-7)
        assert s["trace_accepted"] == gold[name]["trace_accepted"]
        np.testing.assert_allclose(np.abs(flat.poses).sum(), gold[name]["poses_sum"], rtol=1e-9)


def test_lm_decreases_cost_and_recovers_truth(orc):
    flat, truth = synthetic.make_ba_problem(n_img=10, n_obs_target=3000, track_len=4, seed=5, outlier_frac=0.0, noise_px=0.0)
    c0 = orc.ba_cost(flat, _opts(orc))
    s = orc.solve_flat(flat, _opts(orc, 30)).as_dict()
    costs = [c for c, a in zip(s["trace_cost"], s["trace_accepted"]) if a]
    assert all(b <= a for a, b in zip(costs, costs[1:]))
    assert abs(s["initial_cost"] - c0) < 1e-9 * c0
    assert s["final_cost"] < 1e-12 * c0          # noise-free: exact minimum is zero
    # gauge fixed by image 0 (FIXED) and tx of image 1 (FIXED_X): truth is recovered
    np.testing.assert_allclose(flat.poses, truth["poses"], atol=1e-6)
    np.testing.assert_allclose(flat.pts, truth["pts"], atol=1e-5)


def test_termination_rules(orc):
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["tiny"])
    o = orc.default_options()       # function_tolerance 1e-4 as BundleAdjustmentOptions default
    s = orc.solve_flat(flat.copy(), o).as_dict()
    assert s["termination"] == "FUNCTION_TOLERANCE"
    assert s["return_value"] == pytest.approx(np.sqrt(s["final_cost"] / s["num_residuals"]))
    o.max_num_iterations = 2
    s2 = orc.solve_flat(flat.copy(), o).as_dict()
    assert s2["termination"] == "NO_CONVERGENCE" and s2["num_successful_steps"] + s2["num_unsuccessful_steps"] == 2


def _scene(n_img=4, n_pt=60, seed=0):
    rng = np.random.default_rng(seed)
    fm = mm.FeatureManager()
    cam = fm.add_camera([1000.0, 1000.0, 640.0, 480.0, 1])
    X = rng.uniform([-2, -2, 8], [2, 2, 12], (n_pt, 3))
    ids = []
    for i in range(n_img):
        rvec = rng.normal(0, 0.02, 3); tvec = np.array([0.5 * i, 0.0, 0.0]) + rng.normal(0, 0.01, 3)
        from mavmap_b200.synthetic import _rodrigues
        Xc = X @ _rodrigues(rvec)[0].T + tvec
        uv = np.stack([1000 * Xc[:, 0] / Xc[:, 2] + 640, 1000 * Xc[:, 1] / Xc[:, 2] + 480], axis=1)
        iid = fm.add_image(cam, uv + rng.normal(0, 0.3, uv.shape))
        fm.set_pose(iid, rvec + rng.normal(0, 0.005, 3), tvec + rng.normal(0, 0.02, 3))
        ids.append(iid)
    for p in range(n_pt):
        # the last 10 points are seen by only one image (filtered by min_track_len)
        obs = [(i, p) for i in ids] if p < n_pt - 10 else [(ids[0], p)]
        fm.add_track(X[p] + rng.normal(0, 0.05, 3), obs)
    return fm, ids


def test_flatten_follows_reference_rules():
    fm, ids = _scene()
    opt = mm.BundleAdjustmentOptions(min_track_len=2, print_summary=False)
    flat, image_ids, camera_ids, point_ids = flatten(fm, ids[2:], ids[:1], ids[1:2], opt)
    # residual-block order: free, fixed, fixed_x (bundle_adjustment.cc:511-533)
    assert image_ids == [ids[2], ids[3], ids[0], ids[1]]
    assert flat.n_obs == 4 * 50 and flat.n_pt == 50            # single-view points dropped (count < min_track_len)
    assert np.all(np.diff(flat.obs_img) >= 0)
    assert flat.pose_const.tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1], [0, 1, 0, 0]]
    assert flat.intr_const.tolist() == [1]
    opt.refine_camera_params = True
    assert flatten(fm, ids[2:], ids[:1], ids[1:2], opt)[0].intr_const.tolist() == [0]
    with pytest.raises(ValueError, match="At least 7 parameters"):
        flatten(fm, ids[1:], ids[:1], [], opt)                  # 6 fixed dof < 7
    opt.min_track_len = 1
    with pytest.raises(ValueError, match="Minimum track length"):
        flatten(fm, ids[2:], ids[:1], ids[1:2], opt)
    # GCP ids count 3 dof each and become constant points (bundle_adjustment.cc:459-461, 545-549)
    opt.min_track_len = 2
    gcp = {fm.point2D_to_point3D[fm.image_to_points2D[ids[0]][0]], fm.point2D_to_point3D[fm.image_to_points2D[ids[0]][1]],
           fm.point2D_to_point3D[fm.image_to_points2D[ids[0]][2]]}
    flat = flatten(fm, ids, [], [], opt, gcp_ids=gcp)[0]
    assert int(flat.pt_const.sum()) == 3 and not flat.pose_const.any()


def test_bundle_adjustment_call_surface_with_oracle_engine(orc):
    fm, ids = _scene(seed=3)
    errs = {}
    opt = mm.BundleAdjustmentOptions(update_point3D_errors=True, print_summary=False, max_num_iterations=20)
    before = {i: fm.rvecs[i].copy() for i in ids}
    ret = orc.bundle_adjustment(fm, ids[2:], ids[:1], ids[1:2], opt, errs)
    assert 0 < ret < 1.0                                         # ~0.3 px noise
    assert np.array_equal(fm.rvecs[ids[0]], before[ids[0]])      # FIXED pose untouched
    assert not np.array_equal(fm.rvecs[ids[2]], before[ids[2]])
    assert len(errs) == 50 and all(e >= 0 for e in errs.values())


def test_pose_refinement_oracle(orc):
    rng = np.random.default_rng(2)
    X = rng.uniform([-2, -2, 6], [2, 2, 10], (80, 3))
    from mavmap_b200.synthetic import _rodrigues, project
    rvec, tvec = np.array([0.05, -0.1, 0.02]), np.array([0.3, -0.2, 0.5])
    params = synthetic.INTRINSICS[2] + [2]
    uv = project(2, np.array(params[:8]), X @ _rodrigues(rvec)[0].T + tvec)
    mask = np.ones(80, bool); mask[::7] = False
    uv[~mask] += 50.0                                            # outliers excluded by the mask
    r0, t0 = rvec + 0.02, tvec - 0.05
    ret = orc.pose_refinement(r0, t0, params, uv, X, mask, mm.BundleAdjustmentOptions(print_summary=False, function_tolerance=1e-12))
    np.testing.assert_allclose(r0, rvec, atol=1e-6); np.testing.assert_allclose(t0, tvec, atol=1e-6)
    assert ret < 1e-6


def test_rotation_constraint_functor_matches_reference_formula_and_finite_differences():
    """BARotationConstraintCostFunction (bundle_adjustment.cc:57-111): weight * sqrt of nine squared differences between the
    column-major rotation matrices, INCLUDING the reference's index quirk (the 8th term reads rotmat[6], .cc:103)."""
    from scipy.spatial.transform import Rotation
    from oracle import orc
    rng = np.random.default_rng(4)
    for _ in range(20):
        w, w0, weight = rng.normal(0, 0.8, 3), rng.normal(0, 0.8, 3), float(rng.uniform(0.5, 20))
        rot = Rotation.from_rotvec(w).as_matrix().T.ravel()       # column-major, as ceres::AngleAxisToRotationMatrix fills it
        rot0 = Rotation.from_rotvec(w0).as_matrix().T.ravel()
        pairs = [(0, 0), (3, 1), (6, 2), (1, 3), (4, 4), (7, 5), (2, 6), (6, 7), (8, 8)]       # (rotmat index, rotmat0_ index)
        want = weight * np.sqrt(sum((rot[a] - rot0[b]) ** 2 for a, b in pairs))
        r, J = orc.rot_prior(w, w0, weight)
        assert abs(r - want) < 1e-12 * max(1.0, want)
        for k in range(3):
            e = np.zeros(3); e[k] = 1e-6
            fd = (orc.rot_prior(w + e, w0, weight)[0] - orc.rot_prior(w - e, w0, weight)[0]) / 2e-6
            assert abs(J[k] - fd) < 1e-6 * max(1.0, abs(fd))


def _constraint_scene(seed=5):
    """4-image scene + rotation constraints in the reference's convention: the functor compares R(rvec)' (camera-to-world)
    with R(rvec0) (bundle_adjustment.cc:79-81), and the scene is first rotated by M = R_FM' R_C of the first fixed image."""
    from scipy.spatial.transform import Rotation
    fm, ids = _scene(seed=seed)
    rng = np.random.default_rng(9)
    cons = {i: (Rotation.from_rotvec(fm.rvecs[i]).inv() * Rotation.from_rotvec(rng.normal(0, 0.004, 3))).as_rotvec() for i in ids}
    cons[ids[0]] = np.array(fm.rvecs[ids[0]], dtype=np.float64)        # M = identity: the frames already agree
    return fm, ids, cons


def test_rotate_into_constraint_frame_follows_reference_and_keeps_projections():
    """bundle_adjustment.cc:402-425 + similarity_transform.cc:90-121: X' = M X, [R|t] -> [R M'|t] with M = R_FM' R_C;
    camera coordinates R X + t of every observation are unchanged by construction."""
    from scipy.spatial.transform import Rotation
    from mavmap_b200.ba import _rotate_into_constraint_frame
    fm, ids = _scene(seed=6)
    cons = {ids[0]: np.array([0.4, -0.3, 0.2])}
    Xc_before = {i: Rotation.from_rotvec(fm.rvecs[i]).as_matrix() @ np.asarray(fm.points3D[1]) + fm.tvecs[i] for i in ids}
    R_fm = Rotation.from_rotvec(fm.rvecs[ids[0]]).as_matrix(); R_c = Rotation.from_rotvec(cons[ids[0]]).as_matrix()
    _rotate_into_constraint_frame(fm, ids[:1], cons)
    np.testing.assert_allclose(Rotation.from_rotvec(fm.rvecs[ids[0]]).as_matrix(), R_fm @ (R_fm.T @ R_c).T, atol=1e-12)
    for i in ids:
        np.testing.assert_allclose(Rotation.from_rotvec(fm.rvecs[i]).as_matrix() @ np.asarray(fm.points3D[1]) + fm.tvecs[i], Xc_before[i], atol=1e-10)


def test_bundle_adjustment_with_rotation_constraints_oracle():
    """constrain_rotation (bundle_adjustment.cc:390-446): one extra residual per free image (no loss function).  The
    residual is a norm, not a squared norm, and carries the reference's index quirk, so it does not vanish at the
    constraint; the test holds what the reference guarantees: the extra residuals are counted, they enter the cost, and the
    optimiser trades reprojection error for a smaller constraint residual."""
    import mavmap_b200 as mm
    from oracle import orc
    from mavmap_b200.ba import flatten
    fm, ids, cons = _constraint_scene()
    before = {i: np.array(fm.rvecs[i]) for i in ids}
    prior = lambda f: sum(orc.rot_prior(f.rvecs[i], cons[i], 1.0)[0] for i in ids[2:])
    kw = dict(print_summary=False, max_num_iterations=15, function_tolerance=0, gradient_tolerance=0)
    opt = mm.BundleAdjustmentOptions(constrain_rotation=True, constrain_rotation_weight=1e2, **kw)
    flat, image_ids, _, _ = flatten(fm, ids[2:], ids[:1], ids[1:2], opt)
    ret = orc.bundle_adjustment(fm, ids[2:], ids[:1], ids[1:2], opt, {}, cons)
    assert np.isfinite(ret)
    np.testing.assert_allclose(fm.rvecs[ids[0]], before[ids[0]], atol=1e-9)       # M = I and the image is fixed
    # same scene without the constraints: different rotations, and a return value over 2 n_obs residuals only
    fm2, ids2, _ = _constraint_scene()
    ret2 = orc.bundle_adjustment(fm2, ids2[2:], ids2[:1], ids2[1:2], mm.BundleAdjustmentOptions(**kw), {})
    assert not np.allclose(fm2.rvecs[ids2[2]], fm.rvecs[ids[2]], atol=1e-6) and ret2 < ret
    assert prior(fm) < prior(fm2)                  # the constrained solve ends closer to the constraints than the free one


def test_oracle_reproduces_extra_goldens(orc):
    """tests/golden/extra.json (made by make_golden_extra.py): rotation-constraint functor and a constrained LM trace"""
    sys.path.insert(0, GOLDEN)
    import make_golden_extra as mg
    gold = json.load(open(os.path.join(GOLDEN, "extra.json")))
    for case in gold["rot_prior"]:
        r, J = orc.rot_prior(case["rvec"], case["rvec0"], case["weight"])
        assert abs(r - case["r"]) <= 1e-12 * max(1.0, abs(case["r"]))
        np.testing.assert_allclose(J, case["J"], rtol=1e-10, atol=1e-12)
    flat = mg.constrained_problem()
    s = orc.solve_flat(flat, _opts(orc)).as_dict()
    g = gold["ba_rotation_constraints"]
    np.testing.assert_allclose(s["trace_cost"], g["trace_cost"], rtol=1e-9)
    assert s["trace_accepted"] == g["trace_accepted"] and s["num_residuals"] == g["num_residuals"]


def test_oracle_objective_and_minimum_against_independent_numpy(orc):
    """Independent cross-check of WHAT the oracle minimises (residual of bundle_adjustment.h:131-159 for the OPENCV model,
    Cauchy loss of bundle_adjustment.cc:477-478, gauge by constant parameters): a numpy objective written from
    mavmap_b200.synthetic's own camera-model code must (1) equal the oracle's cost at the start, (2) equal it at the oracle's
    final iterate, and (3) be stationary there (central-difference gradient ~ 0 against its size at the start).
    Pins the objective and the optimum, not the LM trajectory (Ceres is not available: DESIGN.md section 2)."""
    from mavmap_b200.synthetic import _rodrigues, project
    flat, _ = synthetic.make_ba_problem(n_img=5, n_obs_target=500, track_len=4, seed=91, model=2, outlier_frac=0.02)
    o = orc.default_options(); o.max_num_iterations = 100; o.function_tolerance = 1e-16; o.gradient_tolerance = 1e-14; o.parameter_tolerance = 1e-14
    f = flat.copy(); s = orc.solve_flat(f, o).as_dict()
    free_pose = np.ones((flat.n_img, 6), bool)
    for i in range(flat.n_img):
        c = flat.pose_const[i]
        free_pose[i, :3] = not c[0]; free_pose[i, 3] = not c[1]; free_pose[i, 4] = not c[2]; free_pose[i, 5] = not c[3]
    nfp = int(free_pose.sum())

    def objective(z):
        poses = flat.poses.copy(); poses[free_pose] = z[:nfp]
        pts = z[nfp:].reshape(-1, 3)
        c = 0.0
        for i in range(flat.n_img):
            sel = flat.obs_img == i
            Xc = pts[flat.obs_pt[sel]] @ _rodrigues(poses[i, :3])[0].T + poses[i, 3:]
            r = project(2, flat.intr[0], Xc) - flat.obs_xy[sel]
            c += 0.5 * np.sum(np.log1p(np.sum(r * r, axis=1)))       # rho(s) = a^2 log(1 + s / a^2), a = loss_scale_factor = 1
        return c

    def gradient(z, h=1e-6):
        g = np.empty_like(z)
        for k in range(len(z)):
            e = np.zeros_like(z); e[k] = h
            g[k] = (objective(z + e) - objective(z - e)) / (2 * h)
        return g
    z0 = np.concatenate([flat.poses[free_pose], flat.pts.ravel()])
    z1 = np.concatenate([f.poses[free_pose], f.pts.ravel()])
    assert abs(objective(z0) - s["initial_cost"]) < 1e-9 * s["initial_cost"]
    assert abs(objective(z1) - s["final_cost"]) < 1e-9 * s["final_cost"]
    assert s["final_cost"] < 0.5 * s["initial_cost"]
    g0, g1 = gradient(z0), gradient(z1)
    assert np.abs(g1).max() < 1e-5 * np.abs(g0).max()


def test_oracle_first_lm_step_against_dense_numpy(orc):
    """One Levenberg-Marquardt step restated densely in numpy with a NUMERICAL Jacobian (independent of the oracle's dual
    numbers and of its Schur elimination): Triggs-corrected Cauchy residuals (sqrt(rho') scaling, rho'' < 0), Jacobi column
    scaling 1/(1 + ||col||), D^2 = clamp(diag J'J, 1e-6, 1e32) / radius with the initial radius 1e4, step = -scale * y from
    (J'J + D^2) y = J'r.  The cost after that step must equal the oracle's trace entry 1 (SURVEY 8a-a3')."""
    from mavmap_b200.synthetic import _rodrigues, project
    flat, _ = synthetic.make_ba_problem(n_img=5, n_obs_target=500, track_len=4, seed=92, model=2, outlier_frac=0.02)
    s = orc.solve_flat(flat.copy(), _opts(orc, 1)).as_dict()
    assert s["trace_accepted"][1] == 1
    free_pose = np.ones((flat.n_img, 6), bool)
    for i in range(flat.n_img):
        c = flat.pose_const[i]
        free_pose[i, :3] = not c[0]; free_pose[i, 3] = not c[1]; free_pose[i, 4] = not c[2]; free_pose[i, 5] = not c[3]
    nfp = int(free_pose.sum())

    def raw(z):
        poses = flat.poses.copy(); poses[free_pose] = z[:nfp]
        pts = z[nfp:].reshape(-1, 3)
        r = np.empty((flat.n_obs, 2))
        for i in range(flat.n_img):
            sel = flat.obs_img == i
            Xc = pts[flat.obs_pt[sel]] @ _rodrigues(poses[i, :3])[0].T + poses[i, 3:]
            r[sel] = project(2, flat.intr[0], Xc) - flat.obs_xy[sel]
        return r
    cost = lambda z: 0.5 * np.sum(np.log1p(np.sum(raw(z) ** 2, axis=1)))
    z0 = np.concatenate([flat.poses[free_pose], flat.pts.ravel()])
    r0 = raw(z0); n = len(z0)
    J = np.empty((2 * flat.n_obs, n)); h = 1e-6
    for k in range(n):
        e = np.zeros(n); e[k] = h
        J[:, k] = ((raw(z0 + e) - raw(z0 - e)) / (2 * h)).ravel()
    w = np.sqrt(1.0 / (1.0 + np.sum(r0 ** 2, axis=1)))               # sqrt(rho'), one per observation
    rt = (r0 * w[:, None]).ravel(); Jt = J * np.repeat(w, 2)[:, None]
    scale = 1.0 / (1.0 + np.linalg.norm(Jt, axis=0))
    Js = Jt * scale
    A = Js.T @ Js
    D2 = np.clip(np.diag(A), 1e-6, 1e32) / 1e4
    y = np.linalg.solve(A + np.diag(D2), Js.T @ rt)
    z1 = z0 - scale * y
    assert abs(cost(z0) - s["trace_cost"][0]) < 1e-9 * s["trace_cost"][0]
    assert abs(cost(z1) - s["trace_cost"][1]) < 1e-7 * s["trace_cost"][1], (cost(z1), s["trace_cost"][1])
