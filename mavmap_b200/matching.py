"""Host-side mirror of match_brute_force (src/base2d/feature.h:102-110, feature.cc:52-133).

`keypoints*` are [n,2] arrays of cv::KeyPoint::pt (only read for the max_distance mask,
feature.cc:23-49); `descriptors*` are [n,k] fp32 (cv::Mat CV_32F).  Returns the DMatch
fields as three arrays (queryIdx, trainIdx, distance) in the reference's push_back order.
"""
import ctypes as C

import numpy as np

from ._abi import MatchOptions, as_ptr, p_f32, p_i32, p_i64
from ._lib import check, lib

NORM_L2 = 4      # cv::NORM_L2


def _options(ratio_test, max_ratio, max_distance, impl):
    o = MatchOptions()
    o.ratio_test = int(bool(ratio_test))
    o.max_ratio = float(max_ratio)
    o.max_distance = float(max_distance)
    o.impl = int(impl)
    return o


def match_brute_force(keypoints1, descriptors1, keypoints2, descriptors2, ratio_test=True,
                      max_ratio=0.6, max_distance=-1, norm_type=NORM_L2, impl=0):
    if norm_type != NORM_L2:
        raise NotImplementedError("only cv::NORM_L2 (the mapper's choice, sequential_mapper.cc:2080-2082)")
    d1 = np.ascontiguousarray(descriptors1, dtype=np.float32)
    d2 = np.ascontiguousarray(descriptors2, dtype=np.float32)
    n1, n2 = len(d1), len(d2)
    if n1 == 0 or n2 == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32)
    k = d1.shape[1]
    xy1 = xy2 = None
    if max_distance != -1:
        xy1 = np.ascontiguousarray(keypoints1, dtype=np.float32).reshape(-1, 2)
        xy2 = np.ascontiguousarray(keypoints2, dtype=np.float32).reshape(-1, 2)
    cap = max(min(n1, n2), 1)
    q = np.empty(cap, np.int32); t = np.empty(cap, np.int32); dist = np.empty(cap, np.float32)
    n_out = C.c_int32(0)
    o = _options(ratio_test, max_ratio, max_distance, impl)
    check(lib().mm_match_pair(as_ptr(d1, p_f32), n1, as_ptr(d2, p_f32), n2, k, as_ptr(xy1, p_f32), as_ptr(xy2, p_f32),
                              C.byref(o), as_ptr(q, p_i32), as_ptr(t, p_i32), as_ptr(dist, p_f32), C.byref(n_out)))
    m = n_out.value
    return q[:m].copy(), t[:m].copy(), dist[:m].copy()


def read_feature_cache(keypoints_path, descriptors_path):
    """One image of the reference's feature cache (feature_cache.cc:126-163) -> (xy [n,2] float32, descriptors [n,k] float32)."""
    n_kp, rows, cols, typ = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
    kp, dp = str(keypoints_path).encode(), str(descriptors_path).encode()
    check(lib().mm_feature_cache_info(kp, dp, C.byref(n_kp), C.byref(rows), C.byref(cols), C.byref(typ)))
    xy = np.empty((rows.value, 2), np.float32); desc = np.empty((rows.value, cols.value), np.float32)
    check(lib().mm_feature_cache_read(kp, dp, as_ptr(xy, p_f32), as_ptr(desc, p_f32), rows.value, cols.value))
    return xy, desc


def write_feature_cache(keypoints_path, descriptors_path, xy, desc, junk=0x7f3a9c10):
    """Writes the two files byte-for-byte as feature_cache.cc:126-147 does (test fixtures / interchange): 28-byte
    cv::KeyPoint records, and the descriptor header whose 8-byte rows/cols fields carry cols / pointer bits in their high
    halves (`junk` stands in for the low half of cv::Mat::data)."""
    xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2); desc = np.ascontiguousarray(desc, dtype=np.float32)
    n = len(xy)
    kp = np.zeros(n, dtype=np.dtype([("pt", "<f4", (2,)), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]))
    kp["pt"] = xy; kp["size"] = 7.0; kp["angle"] = -1.0; kp["class_id"] = -1
    assert kp.dtype.itemsize == 28
    with open(keypoints_path, "wb") as f:
        f.write(np.uint64(n * 28).tobytes()); f.write(kp.tobytes())
    rows, cols = (desc.shape if desc.ndim == 2 else (0, 0))
    with open(descriptors_path, "wb") as f:
        f.write(np.uint64(desc.nbytes).tobytes())
        f.write(np.array([rows, cols], dtype="<i4").tobytes())           # &rows read as 8 bytes: rows | cols
        f.write(np.array([cols, junk], dtype="<i4").tobytes())           # &cols read as 8 bytes: cols | low half of the data pointer
        f.write(np.int32(5).tobytes())                                   # CV_32FC1
        f.write(desc.tobytes())


class MatchSet:
    """Descriptors of a whole sequence resident in HBM (BASELINE.json configs[2]: all pairs)."""

    @classmethod
    def from_cache(cls, keypoints_paths, descriptors_paths):
        """All images straight from the reference's cache files (no cv::Mat in between)."""
        self = cls.__new__(cls)
        n = len(keypoints_paths)
        kp = (C.c_char_p * n)(*[str(p).encode() for p in keypoints_paths]); dp = (C.c_char_p * n)(*[str(p).encode() for p in descriptors_paths])
        counts, k = [], 0
        for a, b in zip(keypoints_paths, descriptors_paths):
            nk, r, c, t = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
            check(lib().mm_feature_cache_info(str(a).encode(), str(b).encode(), C.byref(nk), C.byref(r), C.byref(c), C.byref(t)))
            counts.append(r.value); k = k or c.value
        self.counts = np.array(counts, dtype=np.int32); self.k = k or 64
        self._h = C.c_void_p()
        check(lib().mm_match_set_create_from_cache(kp, dp, n, C.byref(self._h)))
        return self

    def __init__(self, descriptors, keypoints=None):
        """descriptors: list of [n_i,k] arrays or one [n_images,n,k] array."""
        descs = [np.ascontiguousarray(d, dtype=np.float32) for d in descriptors]
        self.counts = np.array([len(d) for d in descs], dtype=np.int32)
        self.k = descs[0].shape[1]
        cat = np.ascontiguousarray(np.concatenate(descs, axis=0))
        xy = None
        if keypoints is not None:
            xy = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float32).reshape(-1, 2) for x in keypoints]))
        self._h = C.c_void_p()
        check(lib().mm_match_set_create(as_ptr(cat, p_f32), as_ptr(xy, p_f32), as_ptr(self.counts, p_i32),
                                        len(descs), self.k, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mm_match_set_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def match_pairs(self, pairs, ratio_test=True, max_ratio=0.6, max_distance=-1, impl=0):
        """pairs: [m,2] image indices. Returns (offsets [m+1], q, t, dist)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        ia = np.ascontiguousarray(pairs[:, 0]); ib = np.ascontiguousarray(pairs[:, 1])
        cap = int(np.minimum(self.counts[ia], self.counts[ib]).sum()) if len(pairs) else 0
        cap = max(cap, 1)
        off = np.zeros(len(pairs) + 1, np.int64)
        q = np.empty(cap, np.int32); t = np.empty(cap, np.int32); dist = np.empty(cap, np.float32)
        o = _options(ratio_test, max_ratio, max_distance, impl)
        check(lib().mm_match_set_pairs(self._h, as_ptr(ia, p_i32), as_ptr(ib, p_i32), len(pairs), C.byref(o),
                                       as_ptr(off, p_i64), as_ptr(q, p_i32), as_ptr(t, p_i32), as_ptr(dist, p_f32), cap))
        m = int(off[-1])
        return off, q[:m].copy(), t[:m].copy(), dist[:m].copy()

    def match_pairs_device(self, pairs, cnt_ptr, q_ptr, t_ptr, dist_ptr, stride, stream=0, ratio_test=True,
                           max_ratio=0.6, max_distance=-1, impl=0):
        """Enqueue on `stream`; outputs stay in HBM (device pointers as ints)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        ia = np.ascontiguousarray(pairs[:, 0]); ib = np.ascontiguousarray(pairs[:, 1])
        o = _options(ratio_test, max_ratio, max_distance, impl)
        check(lib().mm_match_set_pairs_dev(self._h, as_ptr(ia, p_i32), as_ptr(ib, p_i32), len(pairs), C.byref(o),
                                           C.c_void_p(cnt_ptr), C.c_void_p(q_ptr), C.c_void_p(t_ptr), C.c_void_p(dist_ptr),
                                           int(stride), C.c_void_p(stream)))
