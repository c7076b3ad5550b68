"""Adds tests/golden/extra.json: oracle outputs for the parts added after the first fixtures (self-golden):
rotation-constraint functor values, a BA trace with rotation constraints, RANSAC hypothesis scores."""
import json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import orc  # noqa: E402
from mavmap_b200 import synthetic  # noqa: E402


def rot_prior_cases():
    rng = np.random.default_rng(11)
    return [(rng.normal(0, 0.9, 3), rng.normal(0, 0.9, 3), float(rng.uniform(0.5, 30))) for _ in range(12)]


def constrained_problem():
    from scipy.spatial.transform import Rotation
    flat, _ = synthetic.make_ba_problem(outlier_frac=0.0, **synthetic.BA_CONFIGS["tiny"])
    rng = np.random.default_rng(12)
    r0 = np.stack([(Rotation.from_rotvec(p[:3]).inv() * Rotation.from_rotvec(rng.normal(0, 0.01, 3))).as_rotvec() for p in flat.poses])
    w = np.full(flat.n_img, 20.0); w[:2] = 0.0
    flat.set_rotation_constraints(r0, w)
    return flat


def main():
    from test_oracle_geometry import _ransac_case
    out = {"rot_prior": [], "ransac": {}}
    for w, w0, weight in rot_prior_cases():
        r, J = orc.rot_prior(w, w0, weight)
        out["rot_prior"].append({"rvec": w.tolist(), "rvec0": w0.tolist(), "weight": weight, "r": r, "J": J.tolist()})
    flat = constrained_problem()
    o = orc.default_options(); o.max_num_iterations = 8; o.function_tolerance = 0; o.gradient_tolerance = 0
    s = orc.solve_flat(flat, o).as_dict()
    out["ba_rotation_constraints"] = {"trace_cost": s["trace_cost"], "trace_radius": s["trace_radius"], "trace_accepted": s["trace_accepted"],
                                      "num_residuals": s["num_residuals"], "poses_sum": float(np.abs(flat.poses).sum())}
    for kind in (0, 1, 2):
        models, x, y, thr = _ransac_case(kind, n=2000, h=16, seed=21)
        r = orc.ransac_score(kind, models, x, y, thr)
        out["ransac"][str(kind)] = {"num_inliers": r["num_inliers"].tolist(), "residual_sum": r["residual_sum"].tolist(), "best": r["best"],
                                    "residual_checksum": float(np.abs(r["residuals"]).sum()), "mask_count": int(r["inlier_mask"].sum())}
    json.dump(out, open(os.path.join(HERE, "extra.json"), "w"), indent=1)
    print("extra.json written")


if __name__ == "__main__":
    main()
