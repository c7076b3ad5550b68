"""GPU parity tests (through the C ABI): match_brute_force index lists are bit-exact against the
oracle and against the cv2.BFMatcher golden lists."""
import os

import numpy as np
import pytest

from mavmap_b200 import synthetic
from test_oracle_match import VARIANTS, golden_cases

pytestmark = pytest.mark.gpu
IMPLS = {"simt": 1, "auto": 0, "tcgen05": 2}


@pytest.mark.parametrize("impl", list(IMPLS))
def test_golden_cv2_index_lists(mm, orc, impl):
    n = 0
    for name, vname, d1, d2, xy1, xy2, kw, q, t, d in golden_cases():
        if impl == "tcgen05" and kw["max_distance"] != -1.0:
            continue            # the keypoint mask runs on the SIMT kernel (AUTO falls back to it)
        qg, tg, dg = mm.match_brute_force(xy1, d1, xy2, d2, kw["ratio_test"], kw["max_ratio"], kw["max_distance"], impl=IMPLS[impl])
        assert np.array_equal(qg, q) and np.array_equal(tg, t), (name, vname, impl)
        qo, to, do = orc.match_pair(d1, d2, xy1, xy2, **kw)
        assert np.array_equal(dg, do), (name, vname, impl)           # distances bit-exact vs the oracle
        n += 1
    assert n >= 20


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("k", [64, 128])
def test_seeded_pairs_vs_oracle(mm, orc, impl, k):
    desc, xy = synthetic.make_descriptors(3, 1500, k, seed=0xF00D + k)
    for a, b in [(0, 1), (1, 2), (2, 0)]:
        for kw in VARIANTS.values():
            if impl == "tcgen05" and kw["max_distance"] != -1.0:
                continue
            qg, tg, dg = mm.match_brute_force(xy[a], desc[a], xy[b], desc[b], kw["ratio_test"], kw["max_ratio"], kw["max_distance"], impl=IMPLS[impl])
            qo, to, do = orc.match_pair(desc[a], desc[b], xy[a], xy[b], **kw)
            assert np.array_equal(qg, qo) and np.array_equal(tg, to) and np.array_equal(dg, do)
        assert len(qo) > 100


def test_edge_cases(mm):
    a = np.random.default_rng(0).normal(size=(10, 8)).astype(np.float32)
    e = np.zeros((0, 8), np.float32)
    for x, y in [(e, a), (a, e), (e, e)]:
        assert len(mm.match_brute_force(None, x, None, y)[0]) == 0
    assert len(mm.match_brute_force(None, a, None, a[:1], True, 0.9)[0]) == 0
    q, t, d = mm.match_brute_force(None, a, None, a, True, 0.9)
    assert np.array_equal(q, np.arange(10)) and np.array_equal(t, np.arange(10)) and np.all(d == 0)
    q, t, d = mm.match_brute_force(None, a, None, np.concatenate([a, a]), True, 0.9)
    assert np.array_equal(t, np.arange(10))
    with pytest.raises(ValueError):
        mm.match_brute_force(None, a, None, a, True, 0.9, 10.0)      # mask needs keypoints


def test_full_size_pair_properties(mm, orc):
    """BASELINE config size (5000 x 5000 x 64): symmetry and agreement with the oracle."""
    desc, xy = synthetic.make_descriptors(2, 5000, 64, seed=0xF00D + 2)
    q, t, d = mm.match_brute_force(xy[0], desc[0], xy[1], desc[1], True, 0.9, -1)
    q2, t2, d2 = mm.match_brute_force(xy[1], desc[1], xy[0], desc[0], True, 0.9, -1)
    # the cross-checked relation is symmetric: swapping the images transposes the match set
    assert sorted(zip(q.tolist(), t.tolist())) == sorted(zip(t2.tolist(), q2.tolist()))
    assert np.all(np.diff(q) > 0) and len(set(t.tolist())) == len(t)
    assert 2500 < len(q) <= 3100                                     # 60 % shared descriptors (+ a few chance matches)
    qo, to, do = orc.match_pair(desc[0], desc[1], ratio_test=True, max_ratio=0.9)
    assert np.array_equal(q, qo) and np.array_equal(t, to) and np.array_equal(d, do)


def test_match_set_batch_equals_single_pairs(mm):
    desc, xy = synthetic.make_descriptors(5, 700, 64, seed=5)
    ms = mm.MatchSet(desc, xy)
    pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)]
    off, q, t, d = ms.match_pairs(pairs, True, 0.9, -1)
    for p, (i, j) in enumerate(pairs):
        qs, ts, ds = mm.match_brute_force(xy[i], desc[i], xy[j], desc[j], True, 0.9, -1)
        sl = slice(off[p], off[p + 1])
        assert np.array_equal(q[sl], qs) and np.array_equal(t[sl], ts) and np.array_equal(d[sl], ds)
    ms.close()


def test_match_pair_resident_slots(mm, orc):
    """mm_match_pair keeps the descriptor arrays of the last images resident (recognised by address, shape and sampled
    content): call sequences that reuse, evict, rewrite, resize and alias the host arrays all see the current contents."""
    desc, xy = synthetic.make_descriptors(7, 1300, 64, seed=0xBEEF)
    imgs = [np.ascontiguousarray(desc[i]) for i in range(7)]
    pts = [np.ascontiguousarray(xy[i]) for i in range(7)]

    def check(a, b, xa=None, xb=None, variant="ratio09"):
        kw = VARIANTS[variant]
        qg, tg, dg = mm.match_brute_force(xa, a, xb, b, kw["ratio_test"], kw["max_ratio"], kw["max_distance"])
        qo, to, do = orc.match_pair(a, b, xa, xb, **kw)
        assert np.array_equal(qg, qo) and np.array_equal(tg, to) and np.array_equal(dg, do)
        return len(qg)

    from mavmap_b200 import _lib
    c0 = _lib.match_pair_counters()
    for i in range(2, 7):                                   # the mapper's pattern: more images than slots, so slots are evicted
        assert check(imgs[i - 1], imgs[i]) > 100
        assert check(imgs[i - 2], imgs[i]) > 100
    c1 = _lib.match_pair_counters()
    assert c1[0] - c0[0] == 10
    if not os.environ.get("MM_MATCH_PAIR_NO_CACHE"):
        assert c1[1] - c0[1] <= 7                           # every image uploaded once, not once per call
    check(imgs[0], imgs[6])                                 # an evicted image comes back
    imgs[6][:] = imgs[1][::-1]                              # rewritten in place: same address, new content
    check(imgs[5], imgs[6])
    assert check(imgs[6], imgs[1]) >= 1000                  # (a permutation of image 1)
    assert check(imgs[2], imgs[2], variant="mutual") >= 1000    # the same array on both sides
    check(imgs[3][:100], imgs[4][:777])                     # same addresses as resident arrays, fewer rows
    check(imgs[3], imgs[4][:1])
    check(imgs[1], imgs[2], pts[1], pts[2], "mask")         # keypoint mask (CUDA-core path) on resident slots
    check(imgs[2], imgs[3], pts[2], pts[3], "mask")
    big, _ = synthetic.make_descriptors(2, 3100, 64, seed=77)      # larger than the slots: the workspace is resized
    assert check(np.ascontiguousarray(big[0]), np.ascontiguousarray(big[1])) > 100
    assert check(imgs[1], imgs[2]) > 100
    d128, _ = synthetic.make_descriptors(3, 900, 128, seed=78)     # another descriptor length
    for a, b in [(0, 1), (1, 2), (0, 2)]:
        assert check(np.ascontiguousarray(d128[a]), np.ascontiguousarray(d128[b])) > 100
    check(imgs[4], imgs[5], variant="ratio06")


def test_many_pairs_run_in_bounded_scratch_ranges(mm, monkeypatch):
    """The hit masks and candidate lists are per-pair scratch: a long pair list runs as a sequence of ranges under a budget."""
    desc, xy = synthetic.make_descriptors(5, 700, 64, seed=6)
    ms = mm.MatchSet(desc, None)
    pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)]
    want = ms.match_pairs(pairs, True, 0.9, -1)
    for words in (300000, 1):                               # two pairs per range (135 168 words per pair), one pair per range
        monkeypatch.setenv("MM_MATCH_TC_SCRATCH_WORDS", str(words))
        got = ms.match_pairs(pairs, True, 0.9, -1)
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
    ms.close()
