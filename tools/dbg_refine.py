import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from mavmap_b200 import synthetic
from mavmap_b200.ba import default_c_options, solve_flat
from oracle import orc
for model in (1, 2, 3):
    flat, truth = synthetic.make_ba_problem(model=model, refine_camera_params=True, **synthetic.BA_CONFIGS["tiny"])
    flat.intr[0, :2] *= 1.01; flat.intr[0, 2:4] += 3.0
    oo = orc.default_options(); oo.max_num_iterations = 10; oo.function_tolerance = 0; oo.gradient_tolerance = 0
    c = flat.copy(); so = orc.solve_flat(c, oo).as_dict()
    for rep in range(3):
        og = default_c_options(); og.max_num_iterations = 10; og.function_tolerance = 0; og.gradient_tolerance = 0
        g = flat.copy(); sg = solve_flat(g, og).as_dict()
        rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))
        print(model, rep, "cost rel", max(abs(x - y) / y for x, y in zip(sg["trace_cost"], so["trace_cost"])), "poses", rel(g.poses, c.poses), "pts", rel(g.pts, c.pts), "intr", rel(g.intr, c.intr), "acc", sg["trace_accepted"] == so["trace_accepted"], sg["trace_linear_iterations"])
