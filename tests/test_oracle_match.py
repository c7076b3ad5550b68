"""CPU tests: the matcher oracle against the library the reference delegates to
(cv2.BFMatcher, live and through the committed golden index lists)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

VARIANTS = {"ratio09": dict(ratio_test=True, max_ratio=0.9, max_distance=-1.0),
            "ratio06": dict(ratio_test=True, max_ratio=0.6, max_distance=-1.0),
            "mutual": dict(ratio_test=False, max_ratio=0.6, max_distance=-1.0),
            "mask": dict(ratio_test=True, max_ratio=0.9, max_distance=40.0)}
CASES = ["shared64", "shared128", "ragged", "ties", "n1_is_1", "n2_is_1", "n2_is_2"]


def golden_cases():
    g = np.load(os.path.join(GOLDEN, "match_cv2.npz"))
    for name in CASES:
        xy1 = g[name + "/xy1"] if name + "/xy1" in g else None
        xy2 = g[name + "/xy2"] if name + "/xy2" in g else None
        for vname, kw in VARIANTS.items():
            key = "%s/%s/q" % (name, vname)
            if key in g:
                yield name, vname, g[name + "/d1"], g[name + "/d2"], xy1, xy2, kw, g[key], g["%s/%s/t" % (name, vname)], g["%s/%s/d" % (name, vname)]


def test_oracle_index_lists_equal_cv2_golden(orc):
    n = 0
    for name, vname, d1, d2, xy1, xy2, kw, q, t, d in golden_cases():
        qo, to, do = orc.match_pair(d1, d2, xy1, xy2, **kw)
        assert np.array_equal(qo, q) and np.array_equal(to, t), (name, vname)
        np.testing.assert_allclose(do, d, rtol=3e-7)      # OpenCV's fp32 summation order differs by a few ulp
        n += 1
    assert n >= 20


def test_oracle_equals_live_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    from mavmap_b200 import synthetic
    desc, xy = synthetic.make_descriptors(3, 350, 64, seed=99)
    for a, b in [(0, 1), (1, 2), (0, 2)]:
        for kw in VARIANTS.values():
            q, t, d = orc.match_pair_cv2(desc[a], desc[b], xy[a], xy[b], **kw)
            qo, to, do = orc.match_pair(desc[a], desc[b], xy[a], xy[b], **kw)
            assert np.array_equal(q, qo) and np.array_equal(t, to)


def test_edge_cases(orc):
    a = np.random.default_rng(0).normal(size=(10, 8)).astype(np.float32)
    e = np.zeros((0, 8), np.float32)
    for x, y in [(e, a), (a, e), (e, e)]:
        q, t, d = orc.match_pair(x, y)
        assert len(q) == 0
    # fewer than two candidates -> no match survives the size()>=2 test (feature.cc:87-94)
    q, t, d = orc.match_pair(a, a[:1], ratio_test=True, max_ratio=0.9)
    assert len(q) == 0
    # identical sets: d0 == 0, d1 > 0 -> ratio 0 passes, every row matches itself
    q, t, d = orc.match_pair(a, a, ratio_test=True, max_ratio=0.9)
    assert np.array_equal(q, np.arange(10)) and np.array_equal(t, np.arange(10)) and np.all(d == 0)
    # duplicated rows: d0 == d1 == 0 -> 0/0 = NaN, NaN > ratio is false, row survives; ties -> lower index
    b = np.concatenate([a, a])
    q, t, d = orc.match_pair(a, b, ratio_test=True, max_ratio=0.9)
    assert np.array_equal(t, np.arange(10))
