"""Stress tool: repeat the same session many times and report any run whose LM trace is not bit-identical to the first
(the engine is deterministic: on a B200 300 x 6 iterations of cfg4 and 1000 x 6 of cfg2 came out identical)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mavmap_b200 import synthetic
from mavmap_b200.ba import default_c_options, BASession
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
flat, _ = synthetic.make_ba_problem(**synthetic.BA_CONFIGS[cfg])
o = default_c_options(); o.max_num_iterations = iters; o.function_tolerance = 0; o.gradient_tolerance = 0
s = BASession(flat, o)
ref = None; bad = 0; t0 = time.time()
for rep in range(reps):
    if rep: s.reset()
    s.iterate(iters)
    d = s.summary().as_dict()
    key = (tuple(d["trace_cost"]), tuple(d["trace_accepted"]))
    if ref is None: ref = key; print("reference", ["%.6e" % c for c in d["trace_cost"]], d["trace_accepted"], flush=True)
    elif key != ref:
        bad += 1
        first = next(i for i, (a, b) in enumerate(zip(d["trace_cost"], ref[0])) if a != b)
        print("rep %d deviates from iteration %d:" % (rep, first), ["%.6e" % c for c in d["trace_cost"]], d["trace_accepted"], d["trace_linear_iterations"], flush=True)
print("%s: %d of %d session runs (%d LM iterations each) deviate, %.1f s" % (cfg, bad, reps, iters, time.time() - t0))
