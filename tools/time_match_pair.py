"""Per-call time of the one-pair entry point (mm_match_pair through match_brute_force) on host arrays, in the mapper's order
(image i against i-1 and i-2).  MM_MATCH_PAIR_TIMING=1 prints the library's per-phase host times (adds synchronisations),
MM_MATCH_PAIR_NO_CACHE=1 uploads both arrays on every call.  usage: python tools/time_match_pair.py [n_feat] [k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mavmap_b200 as mm
from mavmap_b200 import _lib, synthetic

n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_img = 14
desc, _ = synthetic.make_descriptors(n_img, n_feat, k, seed=0xF00D + 3)
imgs = [np.ascontiguousarray(desc[i]) for i in range(n_img)]
seq = [(i - d, i) for i in range(2, n_img) for d in (1, 2)]
for (i, j) in seq[:4]:
    mm.match_brute_force(None, imgs[i], None, imgs[j], True, 0.9, -1)
c0 = _lib.match_pair_counters()
ts = []
for (i, j) in seq[4:]:
    t0 = time.perf_counter()
    q, t, d = mm.match_brute_force(None, imgs[i], None, imgs[j], True, 0.9, -1)
    ts.append(1e6 * (time.perf_counter() - t0))
c1 = _lib.match_pair_counters()
ts = np.array(ts)
print("mm_match_pair %d x %d x %d: %d calls, %d arrays uploaded; us per call mean %.1f median %.1f min %.1f max %.1f -> %.0f pairs/s; matches %d"
      % (n_feat, n_feat, k, len(ts), c1[1] - c0[1], ts.mean(), np.median(ts), ts.min(), ts.max(), 1e6 / ts.mean(), len(q)))
