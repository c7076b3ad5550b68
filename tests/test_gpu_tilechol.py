"""Numeric phase of the sparse tile Cholesky (csrc/tilechol.cuh) on the device against a dense numpy solve:
factorisation (persistent left-looking kernel, TMA-staged tiles, ready flags) + forward / backward substitution,
with and without the dense intrinsics border.  The plan itself is pinned on the CPU in test_tilechol_plan.py."""
import numpy as np
import pytest

from mavmap_b200 import _lib
from mavmap_b200._abi import as_ptr, p_f64, p_i32
from test_tilechol_plan import grid_graph, random_system

pytestmark = pytest.mark.gpu


def device_solve(n, a, b, pos, blocks, ncb, Bm, Cm, rhs, reps=1):
    L = _lib.lib()
    z = np.zeros_like(rhs); msf = np.zeros(1); msa = np.zeros(1)
    a = np.ascontiguousarray(a, dtype=np.int32); b = np.ascontiguousarray(b, dtype=np.int32)
    S = np.ascontiguousarray(blocks, dtype=np.float64)
    Bc = None if Bm is None else np.ascontiguousarray(Bm); Cc = None if Cm is None else np.ascontiguousarray(Cm)
    p = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64)
    _lib.check(L.mm_debug_tilechol_solve(n, len(a), as_ptr(a, p_i32), as_ptr(b, p_i32), as_ptr(p, p_f64), as_ptr(S, p_f64), ncb,
                                         as_ptr(Bc, p_f64), as_ptr(Cc, p_f64), as_ptr(rhs, p_f64), as_ptr(z, p_f64), reps, as_ptr(msf, p_f64), as_ptr(msa, p_f64)))
    return z, float(msf[0]), float(msa[0])


@pytest.mark.parametrize("strips,per,ncb,geom", [(1, 9, 0, True), (4, 30, 0, True), (5, 24, 0, False), (3, 20, 1, True), (4, 16, 7, True), (20, 50, 0, True), (12, 40, 2, True)])
def test_device_factor_and_solve_match_dense(strips, per, ncb, geom):
    rng = np.random.default_rng(100 * strips + per + ncb)
    n, a, b, pos = grid_graph(strips, per, 3, 1, rng, drop=0.15)
    S, blocks, Bm, Cm = random_system(n, a, b, ncb, rng)
    rhs = rng.normal(size=S.shape[0])
    z, msf, msa = device_solve(n, a, b, pos if geom else None, blocks, ncb, Bm, Cm, rhs)
    ref = np.linalg.solve(S, rhs)
    assert np.abs(z - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_device_solve_is_deterministic_and_flags_indefinite():
    rng = np.random.default_rng(11)
    n, a, b, pos = grid_graph(6, 30, 3, 1, rng)
    S, blocks, _, _ = random_system(n, a, b, 0, rng)
    rhs = rng.normal(size=6 * n)
    z1, _, _ = device_solve(n, a, b, pos, blocks, 0, None, None, rhs, reps=3)
    z2, _, _ = device_solve(n, a, b, pos, blocks, 0, None, None, rhs)
    assert np.array_equal(z1, z2)
    blocks[5] = -blocks[5]                      # an indefinite diagonal block must be reported, not silently factored
    with pytest.raises(_lib.MavmapB200Error):
        device_solve(n, a, b, pos, blocks, 0, None, None, rhs)
