"""Scratch: phases of the host-buffer call mm_ba_solve (MM_SETUP_TIMING=1 prints the setup phases on stderr)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mavmap_b200 import synthetic
from mavmap_b200.ba import default_c_options, solve_flat

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS[cfg])
def opts(n):
    o = default_c_options(); o.max_num_iterations = n; o.function_tolerance = 0; o.gradient_tolerance = 0
    return o
w = flat.copy(); solve_flat(w, opts(1))
for rep in range(2):
    f = flat.copy()
    t = time.perf_counter(); s = solve_flat(f, opts(iters)); dt = time.perf_counter() - t
    d = s.as_dict()
    print("e2e %s: %d iterations in %.1f ms -> %.2f it/s; device ms %s" % (cfg, iters, 1e3 * dt, iters / dt, {k: round(v, 1) for k, v in d["ms"].items()}), flush=True)
