"""Scratch timing of the sparse tile Cholesky on a survey-shaped block graph (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_tilechol_plan import grid_graph, plan_arrays
from test_gpu_tilechol import device_solve

strips, per = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50, 100)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(1)
n, a, b, pos = grid_graph(strips, per, 5, 2, rng)
blocks = rng.normal(0, 0.1, (n + len(a), 6, 6))
# diagonally dominant diagonal blocks
rowsum = np.zeros((n, 6))
np.add.at(rowsum, a, np.abs(blocks[n:]).sum(axis=2)); np.add.at(rowsum, b, np.abs(blocks[n:]).sum(axis=1))
for i in range(n):
    D = blocks[i] + blocks[i].T
    blocks[i] = D + np.diag(rowsum[i] + np.abs(D).sum(axis=1) + 1.0)
rhs = rng.normal(size=6 * n)
plan = plan_arrays(n, a, b, pos, 0)
T = plan["T"]
gf = 2.0 * T ** 3 * (plan["n_upd"] + plan["n_l"]) / 1e9
t = time.time(); z, msf, msa = device_solve(n, a, b, pos, blocks, 0, None, None, rhs, reps=reps); wall = time.time() - t
# residual through the block structure
Sz = np.einsum("nij,nj->ni", blocks[:n], z.reshape(n, 6))
zz = z.reshape(n, 6)
np.add.at(Sz, a, np.einsum("nij,nj->ni", blocks[n:], zz[b])); np.add.at(Sz, b, np.einsum("nji,nj->ni", blocks[n:], zz[a]))
res = np.linalg.norm(Sz.ravel() - rhs) / np.linalg.norm(rhs)
print("grid %dx%d: %d images, %d tiles of L (%.0f MB), %d products, %.1f GFLOP" % (strips, per, n, plan["n_l"], plan["n_l"] * T * T * 8 / 1e6, plan["n_upd"], gf))
print("factor %.3f ms (%.1f TFLOP/s)   apply %.3f ms   relative residual %.2e   wall %.1f s" % (msf, gf / msf, msa, res, wall))
