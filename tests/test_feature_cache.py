"""Reader for the reference's on-disk feature cache (src/base2d/feature_cache.cc:126-163), SURVEY 8f-2."""
import numpy as np
import pytest

from mavmap_b200 import matching, synthetic


def _files(tmp_path, n_img=3, n=400, k=64, seed=1):
    desc, xy = synthetic.make_descriptors(n_img, n, k, seed=seed)
    kps, dps = [], []
    for i in range(n_img):
        kp, dp = tmp_path / ("img%03d.keypoints" % i), tmp_path / ("img%03d.descriptors" % i)
        matching.write_feature_cache(kp, dp, xy[i][: n - 7 * i], desc[i][: n - 7 * i])       # ragged counts
        kps.append(kp); dps.append(dp)
    return desc, xy, kps, dps


def test_read_back_including_the_size_field_quirk(tmp_path):
    desc, xy, kps, dps = _files(tmp_path)
    # byte layout the reference writes: size_t | rows as 8 bytes (rows | cols << 32) | cols as 8 bytes (cols | junk << 32) | int type
    raw = open(dps[1], "rb").read(28)
    assert int(np.frombuffer(raw, "<u8", 1, 0)[0]) == (400 - 7) * 64 * 4
    assert int(np.frombuffer(raw, "<u8", 1, 8)[0]) == (400 - 7) | (64 << 32) and int(np.frombuffer(raw, "<u8", 1, 16)[0]) >> 32 != 0
    for i in range(3):
        x, d = matching.read_feature_cache(kps[i], dps[i])
        assert d.shape == (400 - 7 * i, 64)
        np.testing.assert_array_equal(d, desc[i][: 400 - 7 * i]); np.testing.assert_array_equal(x, xy[i][: 400 - 7 * i].astype(np.float32))


def test_rejects_inconsistent_files(tmp_path):
    desc, xy, kps, dps = _files(tmp_path)
    with pytest.raises(Exception):
        matching.read_feature_cache(kps[0], dps[1])            # 400 keypoints, 393 descriptor rows
    with pytest.raises(Exception):
        matching.read_feature_cache(tmp_path / "missing.keypoints", dps[0])
    open(dps[2], "ab").close(); data = open(dps[2], "rb").read()
    open(dps[2], "wb").write(data[:-100])                      # truncated payload
    with pytest.raises(Exception):
        matching.read_feature_cache(kps[2], dps[2])


@pytest.mark.gpu
def test_match_set_from_cache_equals_match_set_from_arrays(tmp_path, mm):
    desc, xy, kps, dps = _files(tmp_path, n_img=3, n=600)
    a = matching.MatchSet.from_cache(kps, dps)
    b = matching.MatchSet([desc[i][: 600 - 7 * i] for i in range(3)], [xy[i][: 600 - 7 * i] for i in range(3)])
    pairs = [(0, 1), (0, 2), (1, 2)]
    for kw in (dict(ratio_test=True, max_ratio=0.9), dict(ratio_test=True, max_ratio=0.9, max_distance=300.0)):
        ra, rb = a.match_pairs(pairs, **kw), b.match_pairs(pairs, **kw)
        for u, v in zip(ra, rb):
            np.testing.assert_array_equal(u, v)
    a.close(); b.close()


def test_header_that_announces_more_than_the_file_holds(tmp_path):
    """a corrupt size field must be rejected before anything is allocated from it"""
    desc, xy, kps, dps = _files(tmp_path, n_img=1)
    raw = bytearray(open(kps[0], "rb").read()); raw[0:8] = np.uint64(28 * 10 ** 9).tobytes(); open(kps[0], "wb").write(bytes(raw))
    with pytest.raises(Exception):
        matching.read_feature_cache(kps[0], dps[0])
