"""GPU parity tests (through the C ABI): bundle adjustment / pose refinement against the oracle.

Tolerance (north_star): cost and every pose/point parameter within 1e-6 relative after the
same iteration count, denominators max(|x|, 1)."""
import numpy as np
import pytest

from mavmap_b200 import synthetic
from mavmap_b200.ba import default_c_options, solve_flat

pytestmark = pytest.mark.gpu
REL = 1e-6


def _both(orc, flat, iters, **kw):
    og = default_c_options(); og.max_num_iterations = iters; og.function_tolerance = 0; og.gradient_tolerance = 0
    oo = orc.default_options(); oo.max_num_iterations = iters; oo.function_tolerance = 0; oo.gradient_tolerance = 0
    for k, v in kw.items():
        setattr(og, k, v); setattr(oo, k, v)
    g, c = flat.copy(), flat.copy()
    sg = solve_flat(g, og).as_dict(); so = orc.solve_flat(c, oo).as_dict()
    return g, c, sg, so


def _opts(iters):
    o = default_c_options(); o.max_num_iterations = iters; o.function_tolerance = 0; o.gradient_tolerance = 0
    return o


def _assert_parity(g, c, sg, so, param_rel=REL):
    assert sg["trace_accepted"] == so["trace_accepted"]
    np.testing.assert_allclose(sg["trace_cost"], so["trace_cost"], rtol=REL)
    np.testing.assert_allclose(sg["trace_radius"], so["trace_radius"], rtol=1e-5)
    assert abs(sg["return_value"] - so["return_value"]) <= REL * so["return_value"]
    for a, b in ((g.poses, c.poses), (g.pts, c.pts), (g.intr, c.intr)):
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)) < param_rel


@pytest.mark.parametrize("model", [1, 2, 3])
def test_small_problem_parity_per_camera_model(mm, orc, model):
    flat, _ = synthetic.make_ba_problem(model=model, **synthetic.BA_CONFIGS["tiny"])
    _assert_parity(*_both(orc, flat, 10))


@pytest.mark.parametrize("model", [1, 2, 3])
def test_refine_camera_params_parity(mm, orc, model):
    """refine_camera_params=true (the mapper's default, mapper.cc:878-886): one shared intrinsics block,
    dense border of the reduced camera system.

    Parameter tolerance: an 8-image problem with free distortion coefficients is weakly determined; the oracle differs
    from ITSELF by 2.8e-7 / 3.4e-7 (OPENCV) and 1.9e-7 (CATA) relative between 1-, 5- and 8-thread runs (summation order only,
    measured in this container).  Cost trace, radius trace and step pattern are held to 1e-6; the parameters to 10x the
    oracle's own spread where that spread is within a factor 3 of 1e-6."""
    flat, truth = synthetic.make_ba_problem(model=model, refine_camera_params=True, **synthetic.BA_CONFIGS["tiny"])
    flat.intr[0, :2] *= 1.01; flat.intr[0, 2:4] += 3.0          # start from a perturbed calibration
    g, c, sg, so = _both(orc, flat, 10)
    _assert_parity(g, c, sg, so, param_rel=REL if model == 1 else 4e-6)
    assert not np.allclose(g.intr, flat.intr)                    # the intrinsics did move ...
    assert abs(g.intr[0, 0] - truth["intr"][0, 0]) < abs(flat.intr[0, 0] - truth["intr"][0, 0])   # ... towards the truth


def test_refine_camera_params_medium(mm, orc):
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, refine_camera_params=True, **synthetic.BA_CONFIGS["small"])
    flat.intr[0, :2] *= 0.995
    _assert_parity(*_both(orc, flat, 8))


def test_cfg1_parity(mm, orc):
    # BASELINE.json configs[0]: 20-image PINHOLE sequence
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["cfg1"])
    g, c, sg, so = _both(orc, flat, 15)
    _assert_parity(g, c, sg, so)
    assert sg["final_cost"] < 0.2 * sg["initial_cost"]


def test_medium_problem_parity(mm, orc):
    # outlier-free: every parameter is well determined, so the 1e-6 criterion applies to all of them
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, **synthetic.BA_CONFIGS["small"])
    _assert_parity(*_both(orc, flat, 12))


def test_medium_problem_with_outliers_and_huge_radius(mm, orc):
    """2 % gross outliers leave some points on flat parts of the Cauchy loss; their coordinates are
    ill-determined and even the oracle differs from itself by up to ~1e-3 relative between a 1- and an
    8-thread run (summation order only; measured).  Cost trace and step pattern must
    still agree to 1e-6; parameters are held to the oracle's own reproducibility."""
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["small"])
    for radius, ptol in ((1e4, 2e-5), (1e12, 2e-2)):
        g, c, sg, so = _both(orc, flat, 12, initial_trust_region_radius=radius)
        assert sg["trace_accepted"] == so["trace_accepted"]
        np.testing.assert_allclose(sg["trace_cost"], so["trace_cost"], rtol=REL)
        assert abs(sg["return_value"] - so["return_value"]) <= REL * so["return_value"]
        assert np.max(np.abs(g.poses - c.poses) / np.maximum(np.abs(c.poses), 1.0)) < ptol
        assert np.max(np.abs(g.pts - c.pts) / np.maximum(np.abs(c.pts), 1.0)) < ptol


def test_trivial_loss_and_termination_parity(mm, orc):
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, **synthetic.BA_CONFIGS["tiny"])
    g, c, sg, so = _both(orc, flat, 50, loss_type=0, function_tolerance=1e-6, gradient_tolerance=1e-10)
    assert sg["termination"] == so["termination"] and sg["num_iterations"] == so["num_iterations"]
    _assert_parity(g, c, sg, so)


def test_point_errors_and_constant_points(mm, orc):
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["tiny"])
    flat.pt_const[::5] = 1                       # GCP-style constant points (bundle_adjustment.cc:545-549)
    flat.pt_err = np.zeros(flat.n_pt)
    g, c, sg, so = _both(orc, flat, 8)
    _assert_parity(g, c, sg, so)
    assert np.array_equal(g.pts[::5], flat.pts[::5])
    np.testing.assert_allclose(g.pt_err, c.pt_err, rtol=1e-6, atol=1e-9)


def test_empty_and_invalid_problems(mm):
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["tiny"])
    empty = type(flat)(flat.poses, flat.pose_const, flat.img_cam, flat.intr, flat.cam_model, flat.intr_const, flat.pts, flat.pt_const,
                       np.zeros((0, 2)), np.zeros(0, np.int32), np.zeros(0, np.int32))
    s = solve_flat(empty, default_c_options()).as_dict()
    assert s["termination"] == "EMPTY" and np.isnan(s["return_value"])     # bundle_adjustment.cc:571-573 -> 0/0
    bad = flat.copy(); bad.obs_img = bad.obs_img.copy(); bad.obs_img[0] = 10 ** 6
    with pytest.raises(ValueError):
        solve_flat(bad, default_c_options())


def test_call_surface_bundle_adjustment_and_pose_refinement(mm, orc):
    from test_oracle_ba import _scene
    fm_g, ids = _scene(seed=3); fm_o, _ = _scene(seed=3)
    opt = mm.BundleAdjustmentOptions(update_point3D_errors=True, print_summary=False, max_num_iterations=10,
                                     function_tolerance=0, gradient_tolerance=0)
    eg, eo = {}, {}
    rg = mm.bundle_adjustment(fm_g, ids[2:], ids[:1], ids[1:2], opt, eg)
    ro = orc.bundle_adjustment(fm_o, ids[2:], ids[:1], ids[1:2], opt, eo)
    assert abs(rg - ro) <= REL * ro
    for i in ids:
        np.testing.assert_allclose(fm_g.rvecs[i], fm_o.rvecs[i], atol=REL); np.testing.assert_allclose(fm_g.tvecs[i], fm_o.tvecs[i], atol=REL)
    for p in eo:
        assert abs(eg[p] - eo[p]) < 1e-6
    with pytest.raises(ValueError, match="At least 7 parameters"):
        mm.bundle_adjustment(fm_g, ids[1:], ids[:1], [], opt, eg)
    # pose_refinement
    rng = np.random.default_rng(2)
    X = rng.uniform([-2, -2, 6], [2, 2, 10], (200, 3))
    from mavmap_b200.synthetic import _rodrigues, project
    rvec, tvec = np.array([0.05, -0.1, 0.02]), np.array([0.3, -0.2, 0.5])
    params = synthetic.INTRINSICS[2] + [2]
    uv = project(2, np.array(params[:8]), X @ _rodrigues(rvec)[0].T + tvec) + rng.normal(0, 0.3, (200, 2))
    mask = np.ones(200, bool); mask[::9] = False
    o2 = mm.BundleAdjustmentOptions(print_summary=False, max_num_iterations=10, function_tolerance=0, gradient_tolerance=0)
    rg_, tg_ = rvec + 0.02, tvec - 0.05; ro_, to_ = rg_.copy(), tg_.copy()
    a = mm.pose_refinement(rg_, tg_, params, uv, X, mask, o2); b = orc.pose_refinement(ro_, to_, params, uv, X, mask, o2)
    assert abs(a - b) <= REL * b
    np.testing.assert_allclose(rg_, ro_, atol=REL); np.testing.assert_allclose(tg_, to_, atol=REL)


@pytest.mark.parametrize("cfg,n_img,iters", [("cfg2", 500, 4), ("cfg4", 5000, 3)])
def test_full_size_configs_against_the_oracle(mm, orc, monkeypatch, cfg, n_img, iters):
    """BASELINE.json configs[1] (500 images, ~1 M observations) and configs[3] (5000 images, 2 M points, 10 M observations) at
    FULL size against the oracle on the same problem: cost trace, step pattern and every pose / point parameter to 1e-6 after the
    same number of LM iterations, and the PCG around the tile factorisation far from its iteration cap."""
    monkeypatch.delenv("ORC_DETERMINISTIC", raising=False)      # all host cores: one thread would need minutes per LM iteration here
    flat, truth = synthetic.make_ba_problem(**synthetic.BA_CONFIGS[cfg])
    assert flat.n_img == n_img
    g, c, sg, so = _both(orc, flat, iters)
    _assert_parity(g, c, sg, so)
    assert len(sg["trace_cost"]) == iters + 1 and sg["trace_cost"][-1] < 0.2 * sg["trace_cost"][0]
    assert max(sg["trace_linear_iterations"]) <= 3 < default_c_options().pcg_max_iterations


def test_cfg2_scale_properties(mm):
    """BASELINE.json configs[1] at full size: monotone cost, rotations recovered, bit-repeatable from run to run."""
    flat, truth = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["cfg2"])
    assert flat.n_img == 500 and 0.9e6 < flat.n_obs < 1.1e6
    o = default_c_options(); o.max_num_iterations = 6; o.function_tolerance = 0; o.gradient_tolerance = 0
    a, b = flat.copy(), flat.copy()
    sa = solve_flat(a, o).as_dict(); sb = solve_flat(b, o).as_dict()
    costs = [c for c, ok in zip(sa["trace_cost"], sa["trace_accepted"]) if ok]
    assert all(y <= x for x, y in zip(costs, costs[1:])) and costs[-1] < 0.1 * costs[0]       # 2 % gross outliers keep a Cauchy floor
    assert max(sa["trace_linear_iterations"]) < o.pcg_max_iterations
    assert sa["trace_cost"] == sb["trace_cost"] and np.array_equal(a.poses, b.poses) and np.array_equal(a.pts, b.pts)     # fixed-order reductions everywhere
    # rotations are recovered; translations/points keep the free scale of the FIXED + FIXED_X gauge
    assert np.abs(a.poses[:, :3] - truth["poses"][:, :3]).max() < 0.01


# ---- two-level PCG preconditioner (ba_coarse.cuh) ------------------------------------------------------------
@pytest.mark.parametrize("m", [5, 32, 33, 64, 100, 448, 701])
def test_coarse_spd_inverse_kernel(mm, m):
    """the blocked Gauss-Jordan kernel that inverts the coarse matrix, against numpy"""
    import ctypes as C
    from mavmap_b200 import _lib
    rng = np.random.default_rng(m)
    B = rng.normal(size=(m, m + 8))
    A = B @ B.T + 0.5 * np.eye(m)
    out = np.ascontiguousarray(A.copy())
    rc = _lib.lib().mm_debug_spd_inverse(out.ctypes.data_as(C.POINTER(C.c_double)), m)
    assert rc == 0
    ref = np.linalg.inv(A)
    assert np.max(np.abs(out - ref)) <= 1e-9 * np.max(np.abs(ref))
    assert np.max(np.abs(out @ A - np.eye(m))) < 1e-8


def test_two_level_preconditioner_parity_and_iteration_count(mm, orc, monkeypatch):
    """MM_PRECOND_TWO_LEVEL (the fallback when the tiles of the exact factorisation do not fit): same LM trajectory as the
    oracle's direct solve, far fewer PCG iterations than block-Jacobi alone."""
    from mavmap_b200 import _abi
    from mavmap_b200.ba import BASession
    cfg = dict(n_img=120, n_obs_target=120000, track_len=4, seed=777)
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, **cfg)
    oo = orc.default_options(); oo.max_num_iterations = 8; oo.function_tolerance = 0; oo.gradient_tolerance = 0
    c = flat.copy(); so = orc.solve_flat(c, oo).as_dict()
    o = default_c_options(); o.max_num_iterations = 8; o.function_tolerance = 0; o.gradient_tolerance = 0
    o.pcg_preconditioner = _abi.MM_PRECOND_TWO_LEVEL
    g = flat.copy(); sg = solve_flat(g, o).as_dict()
    _assert_parity(g, c, sg, so)
    s = BASession(flat.copy(), o); assert s.coarse_dim() > 0 and s.coarse_dim() % 7 == 0; s.close()
    two_level = sum(sg["trace_linear_iterations"])
    monkeypatch.setenv("MM_PCG_NO_COARSE", "1")
    s = BASession(flat.copy(), o); assert s.coarse_dim() == 0; s.close()
    g1 = flat.copy(); s1 = solve_flat(g1, o).as_dict()
    assert s1["trace_accepted"] == sg["trace_accepted"]
    np.testing.assert_allclose(s1["trace_cost"], sg["trace_cost"], rtol=REL)
    assert two_level * 2 < sum(s1["trace_linear_iterations"])
    assert max(sg["trace_linear_iterations"]) < o.pcg_max_iterations


def test_tile_cholesky_preconditioner_parity_and_iteration_count(mm, orc):
    """default for more than 26 images: PCG preconditioned by the exact sparse tile Cholesky reaches the 1e-13 residual in one
    or two iterations, reproduces the oracle's direct solve, and is bit-reproducible from run to run."""
    cfg = dict(n_img=120, n_obs_target=120000, track_len=4, seed=777)
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, **cfg)
    g, c, sg, so = _both(orc, flat, 8)
    _assert_parity(g, c, sg, so)
    assert 1 <= max(sg["trace_linear_iterations"][1:]) <= 3
    g2 = flat.copy(); o = default_c_options(); o.max_num_iterations = 8; o.function_tolerance = 0; o.gradient_tolerance = 0
    s2 = solve_flat(g2, o).as_dict()
    assert s2["trace_cost"] == sg["trace_cost"] and np.array_equal(g2.poses, g.poses) and np.array_equal(g2.pts, g.pts)


def test_sharded_ba_two_gpus_matches_single_gpu(mm):
    """points sharded across two GPUs with one exchange step per Schur assembly reproduce the single-GPU solve
    (needs two visible GPUs; tools/sharded_ba.py exits non-zero on any mismatch)."""
    import os, subprocess, sys
    from mavmap_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "sharded_ba.py"), "mid", "midrefine"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("model", [1, 2, 3])
def test_pose_refinement_single_kernel_path(mm, orc, model, monkeypatch):
    """pose_refinement (bundle_adjustment.cc:139-225) runs as one single-CTA kernel (ba_pose.cuh): same LM trace and pose as
    the oracle and as the general session engine, including a tolerance-terminated run and an inlier mask."""
    import time
    from mavmap_b200.synthetic import _rodrigues, project
    rng = np.random.default_rng(10 + model)
    n = 1500
    X = rng.uniform([-3, -3, 5], [3, 3, 12], (n, 3))
    rvec, tvec = np.array([0.08, -0.05, 0.03]), np.array([0.2, -0.3, 0.4])
    params = list(synthetic.INTRINSICS[model]) + [model]
    uv = project(model, np.array(params[:-1]), X @ _rodrigues(rvec)[0].T + tvec) + rng.normal(0, 0.4, (n, 2))
    uv[::50] += rng.uniform(-80, 80, (len(uv[::50]), 2))               # gross outliers for the Cauchy loss
    mask = np.ones(n, bool); mask[::7] = False
    for kw in (dict(max_num_iterations=12, function_tolerance=0, gradient_tolerance=0), dict(max_num_iterations=50)):
        opt = mm.BundleAdjustmentOptions(print_summary=False, **kw)
        out = {}
        for name in ("fused", "general", "oracle"):
            r, t = rvec + 0.03, tvec - 0.08
            if name == "general":
                monkeypatch.setenv("MM_POSE_REFINE_GENERAL", "1")
            else:
                monkeypatch.delenv("MM_POSE_REFINE_GENERAL", raising=False)
            t0 = time.perf_counter()
            ret = (orc if name == "oracle" else mm).pose_refinement(r, t, params, uv, X, mask, opt)
            out[name] = (ret, r.copy(), t.copy(), time.perf_counter() - t0)
        for name in ("fused", "general"):
            assert abs(out[name][0] - out["oracle"][0]) <= REL * out["oracle"][0]
            np.testing.assert_allclose(out[name][1], out["oracle"][1], atol=REL)
            np.testing.assert_allclose(out[name][2], out["oracle"][2], atol=REL)
    monkeypatch.delenv("MM_POSE_REFINE_GENERAL", raising=False)
    r, t = rvec + 0.03, tvec - 0.08
    t0 = time.perf_counter()
    for _ in range(20):
        mm.pose_refinement(r.copy(), t.copy(), params, uv, X, mask, opt)
    print("pose_refinement latency: %.0f us per call (%d points)" % ((time.perf_counter() - t0) / 20 * 1e6, int(mask.sum())))


def test_pose_refinement_batch_equals_single_calls(mm, orc):
    """mm_pose_refine_batch (batched across candidate poses / images, one CTA per problem, ONE launch): every problem
    comes out bit-identical to its own mm_pose_refine call and within 1e-6 of the oracle; mixed camera models and ragged sizes."""
    import time
    from mavmap_b200.synthetic import _rodrigues, project
    rng = np.random.default_rng(99)
    B = 12
    rv0, tv0, prm, uvs, Xs, masks = [], [], [], [], [], []
    for b in range(B):
        model = 1 + b % 3; n = int(rng.integers(40, 1800))
        X = rng.uniform([-3, -3, 5], [3, 3, 12], (n, 3))
        rvec, tvec = rng.normal(0, 0.05, 3), rng.normal(0, 0.3, 3)
        params = list(synthetic.INTRINSICS[model]) + [model]
        uv = project(model, np.array(params[:-1]), X @ _rodrigues(rvec)[0].T + tvec) + rng.normal(0, 0.4, (n, 2))
        uv[::40] += rng.uniform(-60, 60, (len(uv[::40]), 2))
        m = np.ones(n, bool); m[::9] = False
        rv0.append(rvec + 0.02); tv0.append(tvec - 0.05); prm.append(params); uvs.append(uv); Xs.append(X); masks.append(m)
    opt = mm.BundleAdjustmentOptions(print_summary=False, max_num_iterations=15, function_tolerance=0, gradient_tolerance=0)
    rb, tb = np.array(rv0), np.array(tv0)
    rets = mm.pose_refinement_batch(rb, tb, prm, uvs, Xs, masks, opt)
    for b in range(B):
        r, t = rv0[b].copy(), tv0[b].copy()
        ret = mm.pose_refinement(r, t, prm[b], uvs[b], Xs[b], masks[b], opt)
        assert ret == rets[b] and np.array_equal(r, rb[b]) and np.array_equal(t, tb[b]), b
        ro, to = rv0[b].copy(), tv0[b].copy()
        reto = orc.pose_refinement(ro, to, prm[b], uvs[b], Xs[b], masks[b], opt)
        assert abs(rets[b] - reto) <= REL * reto
        np.testing.assert_allclose(rb[b], ro, atol=REL); np.testing.assert_allclose(tb[b], to, atol=REL)
    t0 = time.perf_counter()
    for _ in range(5):
        mm.pose_refinement_batch(np.array(rv0), np.array(tv0), prm, uvs, Xs, masks, opt)
    dt_b = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for b in range(B):
        mm.pose_refinement(rv0[b].copy(), tv0[b].copy(), prm[b], uvs[b], Xs[b], masks[b], opt)
    dt_s = time.perf_counter() - t0
    print("pose refinement, %d problems: batch %.0f us, one by one %.0f us" % (B, dt_b * 1e6, dt_s * 1e6))
    with pytest.raises(ValueError):
        mm.pose_refinement_batch(np.zeros((1, 3)), np.zeros((1, 3)), [prm[0]], [np.zeros((0, 2))], [np.zeros((0, 3))], None, opt)


def test_rotation_constraints_parity(mm, orc):
    """constrain_rotation (bundle_adjustment.cc:390-446, functor .cc:57-111 with its index quirk): device engine against the
    oracle, flat problem and reference call surface (which first rotates the whole feature manager)."""
    from test_oracle_ba import _constraint_scene
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, **synthetic.BA_CONFIGS["small"])
    rng = np.random.default_rng(3)
    from scipy.spatial.transform import Rotation
    r0 = np.stack([(Rotation.from_rotvec(p[:3]).inv() * Rotation.from_rotvec(rng.normal(0, 0.01, 3))).as_rotvec() for p in flat.poses])
    w = np.full(flat.n_img, 30.0); w[:2] = 0.0                       # the two gauge images carry no constraint (.cc:428: free images only)
    flat.set_rotation_constraints(r0, w)
    g, c, sg, so = _both(orc, flat, 10)
    _assert_parity(g, c, sg, so)
    assert sg["num_residuals"] == 2 * flat.n_obs + flat.n_img - 2
    plain = flat.copy(); plain.rot_prior = plain.rot_prior_w = None
    assert abs(solve_flat(plain, default_c_options()).as_dict()["initial_cost"] - sg["initial_cost"]) > 1e-6 * sg["initial_cost"]
    # call surface
    fm_g, ids, cons = _constraint_scene(); fm_o, _, _ = _constraint_scene()
    opt = mm.BundleAdjustmentOptions(print_summary=False, max_num_iterations=12, function_tolerance=0, gradient_tolerance=0,
                                     constrain_rotation=True, constrain_rotation_weight=50.0)
    cons[ids[0]] = np.array([0.2, -0.1, 0.3])                        # a real change of frame this time
    rg = mm.bundle_adjustment(fm_g, ids[2:], ids[:1], ids[1:2], opt, {}, cons)
    ro = orc.bundle_adjustment(fm_o, ids[2:], ids[:1], ids[1:2], opt, {}, cons)
    assert abs(rg - ro) <= REL * ro
    for i in ids:
        np.testing.assert_allclose(fm_g.rvecs[i], fm_o.rvecs[i], atol=2e-6); np.testing.assert_allclose(fm_g.tvecs[i], fm_o.tvecs[i], atol=2e-6)


def test_refined_intrinsics_medium_through_the_bordered_factorisation(mm, orc):
    """refine_camera_params=true is the mapper's default (mapper.cc:878-886): the intrinsics form the dense border of the
    reduced system, eliminated last by the tile factorisation; the PCG around it needs one or two iterations."""
    cfg = dict(n_img=120, n_obs_target=120000, track_len=4, seed=778)
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, refine_camera_params=True, **cfg)
    flat.intr[0, :2] *= 0.997
    g, c, sg, so = _both(orc, flat, 8)
    _assert_parity(g, c, sg, so, param_rel=4e-6)
    assert 1 <= max(sg["trace_linear_iterations"][1:]) <= 3

def test_mixed_camera_rig_parity(mm, orc):
    """BASELINE.json configs[4]: a rig of two different cameras (PINHOLE + OPENCV) in one sequence, intrinsics fixed
    (bundle_adjustment.cc:339: the model code travels with each camera's parameter vector)."""
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, models=[1, 2], **synthetic.BA_CONFIGS["small"])
    assert flat.n_cam == 2 and sorted(set(flat.img_cam.tolist())) == [0, 1]
    _assert_parity(*_both(orc, flat, 10))
    # three cameras incl. CATA, with the two-level preconditioner switched on (>= 64 images)
    flat3, _ = synthetic.make_ba_problem(outlier_frac=0.0, models=[1, 2, 3], n_img=90, n_obs_target=60000, track_len=4, seed=31)
    _assert_parity(*_both(orc, flat3, 8))


@pytest.mark.parametrize("models,cfg", [([1, 2], "small"), ([1, 2, 3], None), ([2, 2, 1, 3, 1, 2], None)])
def test_refine_camera_params_with_several_cameras(mm, orc, models, cfg):
    """refine_camera_params=true with a rig of several cameras (the mapper's default options on BASELINE.json configs[4]):
    one intrinsics block of 9 per camera id (bundle_adjustment.cc:246-247, sequential_mapper.cc:954-973), shared by the camera's
    images -> a border of 9 n_cam columns on the reduced system, cross-camera blocks where cameras see common points."""
    kw = synthetic.BA_CONFIGS[cfg] if cfg else dict(n_img=90, n_obs_target=60000, track_len=4, seed=31 + len(models))
    flat, truth = synthetic.make_ba_problem(outlier_frac=0.0, models=models, refine_camera_params=True, **kw)
    flat.intr[:, :2] *= 0.997; flat.intr[:, 2:4] += 1.5
    g, c, sg, so = _both(orc, flat, 8)
    _assert_parity(g, c, sg, so, param_rel=4e-6)
    assert not np.allclose(g.intr[:, :4], flat.intr[:, :4])
    pin = np.array(models) != 3               # (the focal length of a CATA camera trades against xi: weakly determined, checked by parity only)
    assert np.all(np.abs(g.intr[pin, 0] - truth["intr"][pin, 0]) < np.abs(flat.intr[pin, 0] - truth["intr"][pin, 0]))       # the cameras moved towards the truth
    # a camera held constant inside a refining adjustment keeps its parameters exactly (and the others still move)
    f2 = flat.copy(); f2.intr_const = f2.intr_const.copy(); f2.intr_const[0] = 1
    g2, c2, sg2, so2 = _both(orc, f2, 6)
    _assert_parity(g2, c2, sg2, so2, param_rel=4e-6)
    assert np.array_equal(g2.intr[0], flat.intr[0]) and not np.allclose(g2.intr[1, :4], flat.intr[1, :4])


def test_refine_camera_params_with_rotation_constraints(mm, orc):
    """refine_camera_params together with constrain_rotation (both allowed by bundle_adjustment.cc:535-540)"""
    from scipy.spatial.transform import Rotation
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, models=[1, 2], refine_camera_params=True, **synthetic.BA_CONFIGS["small"])
    flat.intr[:, :2] *= 1.003
    rng = np.random.default_rng(5)
    r0 = np.stack([(Rotation.from_rotvec(p[:3]).inv() * Rotation.from_rotvec(rng.normal(0, 0.01, 3))).as_rotvec() for p in flat.poses])
    w = np.full(flat.n_img, 30.0); w[:2] = 0.0
    flat.set_rotation_constraints(r0, w)
    g, c, sg, so = _both(orc, flat, 8)
    _assert_parity(g, c, sg, so, param_rel=4e-6)
    assert sg["num_residuals"] == 2 * flat.n_obs + flat.n_img - 2


def test_device_reproduces_committed_golden_traces(mm):
    """the CUDA path against the committed fixtures directly (tests/golden/ba_trace.json, extra.json): LM cost / radius traces,
    step pattern and parameter checksums of the seeded tiny problems, and the rotation-constrained trace"""
    import json, os, sys
    from conftest import GOLDEN
    gold = json.load(open(os.path.join(GOLDEN, "ba_trace.json")))
    for name, kw, model, refine in [("tiny_pinhole", synthetic.BA_CONFIGS["tiny"], 1, False),
                                    ("tiny_opencv", dict(synthetic.BA_CONFIGS["tiny"], seed=77), 2, False),
                                    ("tiny_cata_refine", dict(synthetic.BA_CONFIGS["tiny"], seed=78), 3, True)]:
        flat, _ = synthetic.make_ba_problem(model=model, refine_camera_params=refine, **kw)
        s = solve_flat(flat, _opts(8)).as_dict()
        np.testing.assert_allclose(s["trace_cost"], gold[name]["trace_cost"], rtol=REL)
        np.testing.assert_allclose(s["trace_radius"], gold[name]["trace_radius"], rtol=1e-5)
        assert s["trace_accepted"] == gold[name]["trace_accepted"]
        np.testing.assert_allclose(np.abs(flat.poses).sum(), gold[name]["poses_sum"], rtol=1e-5 if refine else REL)   # (refine: see the oracle's own spread above)
    sys.path.insert(0, GOLDEN)
    import make_golden_extra as mg
    g = json.load(open(os.path.join(GOLDEN, "extra.json")))["ba_rotation_constraints"]
    flat = mg.constrained_problem()
    s = solve_flat(flat, _opts(8)).as_dict()
    np.testing.assert_allclose(s["trace_cost"], g["trace_cost"], rtol=REL)
    assert s["trace_accepted"] == g["trace_accepted"] and s["num_residuals"] == g["num_residuals"]
    np.testing.assert_allclose(np.abs(flat.poses).sum(), g["poses_sum"], rtol=REL)
