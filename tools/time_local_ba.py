"""Latency of small (local-BA sized) problems through mm_ba_solve: wall time per call and the device-side breakdown."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mavmap_b200 import synthetic
from mavmap_b200.ba import default_c_options, solve_flat
for name, kw in (("8 img / 1.5k obs", synthetic.BA_CONFIGS["tiny"]), ("20 img / 12k obs", synthetic.BA_CONFIGS["cfg1"]), ("40 img / 40k obs", synthetic.BA_CONFIGS["small"])):
    flat, _ = synthetic.make_ba_problem(**kw)
    o = default_c_options(); o.max_num_iterations = 10; o.function_tolerance = 0; o.gradient_tolerance = 0
    solve_flat(flat.copy(), o)
    ts = []
    for _ in range(10):
        f = flat.copy(); t = time.perf_counter(); s = solve_flat(f, o); ts.append(time.perf_counter() - t)
    d = s.as_dict()
    print("%-18s %.2f ms per call (min %.2f) | device ms: %s | pcg its %s" % (name, 1e3 * np.median(ts), 1e3 * min(ts), {k: round(v, 2) for k, v in d["ms"].items()}, d["trace_linear_iterations"][1:4]))
