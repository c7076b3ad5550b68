"""Scratch timing of the BA kernels on one GPU (not the bench; prints per-kernel ms)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mavmap_b200 import synthetic
from mavmap_b200.ba import BASession, default_c_options

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
refine = os.environ.get("MM_TIME_REFINE") == "1"
t = time.time(); flat, _ = synthetic.make_ba_problem(refine_camera_params=refine, **synthetic.BA_CONFIGS[cfg]);
if refine: flat.intr[0, :2] *= 0.998
print("gen %.1fs n_img %d n_pt %d n_obs %d" % (time.time() - t, flat.n_img, flat.n_pt, flat.n_obs))
o = default_c_options(); o.max_num_iterations = iters; o.function_tolerance = 0; o.gradient_tolerance = 0
if len(sys.argv) > 3: o.pcg_tolerance = float(sys.argv[3])
t = time.time(); s = BASession(flat, o); print("create %.3fs blocks %d coarse dim %d" % (time.time() - t, s.num_blocks(), s.coarse_dim()))
for name, w in () if os.environ.get("MM_NO_TIME_KERNEL") else (("K1 residual+jacobian", 0), ("K2 schur", 1), ("K4 cost", 2), ("K3 spmv", 3), ("coarse setup", 4)):
    if w == 4 and s.coarse_dim() == 0: continue
    print("%-22s %.4f ms" % (name, s.time_kernel(w, 20)))
t = time.time(); n = s.iterate(iters); dt = time.time() - t
d = s.summary().as_dict()
print("iterations %d in %.3fs -> %.2f it/s" % (n, dt, n / dt))
print("ms:", {k: round(v, 2) for k, v in d["ms"].items()})
print("pcg iters:", d["trace_linear_iterations"])
print("cost:", ["%.6e" % c for c in d["trace_cost"]])
print("accepted:", d["trace_accepted"])
