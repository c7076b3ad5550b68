"""Analyse the per-task time stamps of the tile Cholesky kernels (MM_TC_TRACE=file from tools/time_tilechol.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_tilechol_plan import grid_graph, plan_arrays
strips, per, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
rng = np.random.default_rng(1)
n, a, b, pos = grid_graph(strips, per, 5, 2, rng)
plan = plan_arrays(n, a, b, pos, 0)
tr = np.fromfile(path, dtype=np.uint64).astype(np.float64).reshape(-1, 2)
n_l, n_w, n_st = plan["n_l"], plan["n_wtask"], plan["n_stasks"]
f = tr[:n_l + n_w]; t0 = f[:, 0][f[:, 0] > 0].min()
f = (f - t0) / 1e3
print("factor: span %.1f us" % f[:, 1].max())
nupd = np.concatenate([np.diff(plan["upd_ptr"]), np.diff(plan["wupd_ptr"])])
dur = f[:, 1] - f[:, 0]
diag = np.zeros(n_l + n_w, bool); diag[:n_l] = plan["row_idx"] == plan["col_idx"]
for lo, hi in [(0, 0), (1, 1), (2, 4), (5, 10), (11, 30), (31, 100), (101, 1000)]:
    m = (nupd >= lo) & (nupd <= hi) & ~diag
    if m.any(): print("  off-diag/W tasks with %3d-%3d updates: n %5d  mean duration %.2f us  per update %.2f us" % (lo, hi, m.sum(), dur[m].mean(), (dur[m] / np.maximum(nupd[m], 1)).mean()))
m = diag & (nupd == 0); print("  diag tasks, no updates: n %d mean %.2f us" % (m.sum(), dur[m].mean()))
m = diag & (nupd > 0); print("  diag tasks with updates: n %d mean %.2f us, minus %.2f per update = %.2f" % (m.sum(), dur[m].mean(), 2.0, (dur[m] - 2.0 * nupd[m]).mean()))
# time line: finish time per height
h = plan["tile_height"]
hh = np.concatenate([h[plan["col_idx"]], h[plan["wt_row"]]])
for k in range(h.max() + 1):
    m = hh == k; print("  height %d: tasks %5d  first start %.1f  last end %.1f us" % (k, m.sum(), f[m, 0].min(), f[m, 1].max()))
# idle: fraction of (CTA x time) spent inside tasks
print("  sum of task durations %.1f us = %.1f%% of 296 CTAs x span" % (dur.sum(), 100 * dur.sum() / (296 * f[:, 1].max())))
s = tr[n_l + n_w:]; s0 = s[:, 0][s[:, 0] > 0].min(); s = (s - s0) / 1e3
print("apply: span %.1f us, %d tasks" % (s[:, 1].max(), n_st))
kinds = plan["st_kind"]; items = np.diff(plan["st_item_ptr"]); d = s[:, 1] - s[:, 0]
for k, name in enumerate(["MV", "MVT", "SUM", "MV_OUT"]):
    m = kinds == k
    if m.any(): print("  %-6s n %5d mean items %.1f  mean duration %.2f us  max %.2f" % (name, m.sum(), items[m].mean(), d[m].mean(), d[m].max()))
# level time line of the apply: tile height of each task's row
th = h[plan["st_tile"]]
half = np.argmax(kinds == 1) if (kinds == 1).any() else n_st
for k in range(h.max() + 1):
    m = (th == k) & (np.arange(n_st) < half); print("  fwd height %d: start %.1f end %.1f us" % (k, s[m, 0].min(), s[m, 1].max()))

if os.path.exists(path + ".diag"):
    dg = np.fromfile(path + ".diag", dtype=np.uint64).astype(np.float64).reshape(-1, 8)
    dg = dg[dg[:, 0] > 0]
    names = ["combine+load", "cholesky", "store L", "inverse diag blocks", "inverse off block", "store W", "publish"]
    print("diag task phases (mean us over %d tasks):" % len(dg), {nm: round(float((dg[:, k + 1] - dg[:, k]).mean() / 1e3), 2) for k, nm in enumerate(names)})
