"""Scratch: sweep the knobs of the tile-Cholesky plan (schedule cost model, leaf size, substitution CTAs) on one problem."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mavmap_b200 import synthetic
from mavmap_b200.ba import BASession, default_c_options
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS[cfg])
o = default_c_options(); o.max_num_iterations = 8; o.function_tolerance = 0; o.gradient_tolerance = 0
combos = [{}] + [{"MM_TC_COST": c} for c in ("3.5,3,12,5", "3,2,10,4", "4,3,15,6", "3.5,1.5,10,4", "2.5,3,10,6", "3.5,3,30,8")] \
    + [{"MM_TC_LEAF": l} for l in ("8", "24", "32")] + [{"MM_TC_APPLY_CTAS": a} for a in ("2", "3", "6", "8")]
for env in combos:
    for k in ("MM_TC_COST", "MM_TC_LEAF", "MM_TC_APPLY_CTAS"): os.environ.pop(k, None)
    os.environ.update(env)
    s = BASession(flat.copy(), o)
    f_ms = s.time_kernel(5, 5); a_ms = s.time_kernel(6, 5)
    t = time.time(); n = s.iterate(8); dt = time.time() - t
    info = s.solver_info(); d = s.summary().as_dict(); s.close()
    print("%-32s factor %.3f ms  apply %.3f ms  -> %.2f it/s  (tiles %d, %.1f GFLOP, cost %.6e)" % (env, f_ms, a_ms, n / dt, info["tiles"], info["flops"] / 1e9, d["trace_cost"][-1]), flush=True)
