"""Generates the committed fixtures under tests/golden/ (run in the build container).

  camera_ref.npz   outputs of the REFERENCE's own camera models (oracle/_ref/libref_camera.so,
                   compiled from /root/reference/src/base3d/camera_models.{h,cc}) on the parameter
                   sets of camera_models_test.cc:59-82 plus seeded random points
  match_cv2.npz    index lists produced by the library the reference delegates to
                   (cv2.BFMatcher driven exactly as feature.cc:52-133) on seeded descriptor sets
  ba_trace.json    LM traces of the oracle on seeded problems (self-golden: parity UNPINNED, the
                   reference has no BA test and Ceres is absent)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from mavmap_b200 import synthetic  # noqa: E402

CAMERA_SETS = [   # camera_models_test.cc:59-82
    (1, [651.123, 655.123, 386.123, 511.123]),
    (3, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001, 0]),
    (3, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001, 1]),
    (3, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001, 0.5]),
    (2, [651.123, 655.123, 386.123, 511.123, -0.471, 0.223, -0.001, 0.001]),
]


def match_cases():
    rng = np.random.default_rng(0xF00D)
    cases = {}
    d, xy = synthetic.make_descriptors(2, 400, 64, seed=0xF00D + 1)
    cases["shared64"] = (d[0], d[1], xy[0], xy[1])
    d, xy = synthetic.make_descriptors(2, 300, 128, seed=0xF00D + 2)
    cases["shared128"] = (d[0], d[1], xy[0], xy[1])
    a = rng.normal(size=(257, 64)).astype(np.float32); b = rng.normal(size=(131, 64)).astype(np.float32)
    cases["ragged"] = (a, b, rng.uniform(0, 100, (257, 2)).astype(np.float32), rng.uniform(0, 100, (131, 2)).astype(np.float32))
    # exact ties and duplicates: small-integer descriptors make all distances exact in fp32
    a = rng.integers(0, 3, (120, 16)).astype(np.float32); b = np.concatenate([a[:60], a[:60], rng.integers(0, 3, (40, 16)).astype(np.float32)])
    cases["ties"] = (a, b, rng.uniform(0, 50, (120, 2)).astype(np.float32), rng.uniform(0, 50, (160, 2)).astype(np.float32))
    cases["n1_is_1"] = (a[:1], b, None, None)
    cases["n2_is_1"] = (a, b[:1], None, None)
    cases["n2_is_2"] = (a, b[:2], None, None)
    return cases


def main():
    # ---- camera: the reference itself
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_camera.so")
    ref = C.CDLL(ref_path)
    p = C.POINTER(C.c_double)
    rng = np.random.default_rng(7)
    out = {}
    for k, (code, params) in enumerate(CAMERA_SETS):
        prm = np.zeros(9); prm[:len(params)] = params
        xyz = np.concatenate([np.array([[0.5, 0.23, 1.0], [0.0, 0.0, 1.0]]), rng.uniform([-0.6, -0.6, 0.8], [0.6, 0.6, 3.0], (64, 3))])
        uv = np.empty((len(xyz), 2))
        ref.ref_world2image(code, prm.ctypes.data_as(p), C.c_long(len(xyz)), xyz.ctypes.data_as(p), uv.ctypes.data_as(p))
        uv_in = np.concatenate([np.array([[200.0, 100.0], [params[2], params[3]]]), rng.uniform([50, 50], [750, 950], (64, 2))])
        back = np.empty((len(uv_in), 3)); backn = np.empty((len(uv_in), 2))
        ref.ref_image2world(code, prm.ctypes.data_as(p), C.c_long(len(uv_in)), uv_in.ctypes.data_as(p), back.ctypes.data_as(p))
        ref.ref_image2world_normalized(code, prm.ctypes.data_as(p), C.c_long(len(uv_in)), uv_in.ctypes.data_as(p), backn.ctypes.data_as(p))
        out.update({"code%d" % k: code, "params%d" % k: prm, "xyz%d" % k: xyz, "uv%d" % k: uv, "uv_in%d" % k: uv_in,
                    "xyz_out%d" % k: back, "xy_norm%d" % k: backn})
    np.savez_compressed(os.path.join(HERE, "camera_ref.npz"), n=len(CAMERA_SETS), **out)

    # ---- matcher: cv2.BFMatcher driven as feature.cc
    mout = {}
    variants = [("ratio09", dict(ratio_test=True, max_ratio=0.9, max_distance=-1.0)),
                ("ratio06", dict(ratio_test=True, max_ratio=0.6, max_distance=-1.0)),
                ("mutual", dict(ratio_test=False, max_ratio=0.6, max_distance=-1.0)),
                ("mask", dict(ratio_test=True, max_ratio=0.9, max_distance=40.0))]
    for name, (a, b, xa, xb) in match_cases().items():
        mout[name + "/d1"] = a; mout[name + "/d2"] = b
        if xa is not None:
            mout[name + "/xy1"] = xa; mout[name + "/xy2"] = xb
        for vname, kw in variants:
            if kw["max_distance"] != -1.0 and xa is None:
                continue
            q, t, d = orc.match_pair_cv2(a, b, xa, xb, **kw)
            mout["%s/%s/q" % (name, vname)] = q; mout["%s/%s/t" % (name, vname)] = t; mout["%s/%s/d" % (name, vname)] = d
    np.savez_compressed(os.path.join(HERE, "match_cv2.npz"), **mout)

    # ---- BA: oracle traces (self-golden)
    traces = {}
    for name, kw, model, refine in [("tiny_pinhole", synthetic.BA_CONFIGS["tiny"], 1, False),
                                    ("tiny_opencv", dict(synthetic.BA_CONFIGS["tiny"], seed=77), 2, False),
                                    ("tiny_cata_refine", dict(synthetic.BA_CONFIGS["tiny"], seed=78), 3, True)]:
        flat, _ = synthetic.make_ba_problem(model=model, refine_camera_params=refine, **kw)
        o = orc.default_options(); o.max_num_iterations = 8; o.function_tolerance = 0; o.gradient_tolerance = 0
        s = orc.solve_flat(flat, o).as_dict()
        traces[name] = {"trace_cost": s["trace_cost"], "trace_radius": s["trace_radius"], "trace_accepted": s["trace_accepted"],
                        "final_cost": s["final_cost"], "return_value": s["return_value"],
                        "poses_sum": float(np.abs(flat.poses).sum()), "pts_sum": float(np.abs(flat.pts).sum()),
                        "intr": flat.intr.tolist()}
    with open(os.path.join(HERE, "ba_trace.json"), "w") as f:
        json.dump(traces, f, indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
