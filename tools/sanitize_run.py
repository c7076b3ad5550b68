"""Small invocations of every kernel family, meant to run under compute-sanitizer (memcheck / initcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mavmap_b200 as mm
from mavmap_b200 import synthetic
from mavmap_b200.ba import default_c_options, solve_flat

def opts(n):
    o = default_c_options(); o.max_num_iterations = n; o.function_tolerance = 0; o.gradient_tolerance = 0
    return o
which = sys.argv[1:] or ["ba", "refine", "pose", "match"]
if "ba" in which:          # 40 images: structure setup on the device, tile Cholesky path, download with point errors
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["small"]); flat.pt_err = np.zeros(flat.n_pt)
    s = solve_flat(flat, opts(4)).as_dict(); print("ba small", s["trace_cost"][-1], s["trace_accepted"])
    flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS["tiny"])
    s = solve_flat(flat, opts(3)).as_dict(); print("ba tiny (dense solve)", s["trace_cost"][-1])
if "refine" in which:      # two-camera rig with refined intrinsics (border tiles)
    flat, _ = synthetic.make_ba_problem(n_img=40, n_obs_target=30000, track_len=4, seed=5, models=[1, 2], refine_camera_params=True)
    s = solve_flat(flat, opts(3)).as_dict(); print("ba refine rig", s["trace_cost"][-1])
if "pose" in which:
    from mavmap_b200.synthetic import _rodrigues, project
    rng = np.random.default_rng(1); B = 5
    rv, tv, prm, uvs, Xs = [], [], [], [], []
    for b in range(B):
        n = 100 + 50 * b; X = rng.uniform([-3, -3, 5], [3, 3, 12], (n, 3)); r, t = rng.normal(0, 0.05, 3), rng.normal(0, 0.3, 3)
        p = list(synthetic.INTRINSICS[1 + b % 3]) + [1 + b % 3]
        uvs.append(project(1 + b % 3, np.array(p[:-1]), X @ _rodrigues(r)[0].T + t) + rng.normal(0, 0.4, (n, 2))); Xs.append(X); rv.append(r + 0.02); tv.append(t - 0.05); prm.append(p)
    o = mm.BundleAdjustmentOptions(print_summary=False, max_num_iterations=6)
    print("pose batch", mm.pose_refinement_batch(np.array(rv), np.array(tv), prm, uvs, Xs, None, o))
if "match" in which:
    desc, xy = synthetic.make_descriptors(2, 600, 64, seed=3)
    q, t, d = mm.match_brute_force(xy[0], desc[0], xy[1], desc[1], True, 0.9, -1); print("match tc", len(q))
    q, t, d = mm.match_brute_force(xy[0], desc[0], xy[1], desc[1], True, 0.9, 50.0); print("match masked", len(q))
