"""Scratch timing of the matcher (resident descriptors), SIMT vs tcgen05."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mavmap_b200 import synthetic
from mavmap_b200.matching import MatchSet
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_img, n_feat = 6, 5000
desc, xy = synthetic.make_descriptors(n_img, n_feat, k, seed=0xF00D + 3)
ms = MatchSet(desc, None)
pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]
pairs = pairs * 4
cnt = torch.zeros(len(pairs), dtype=torch.int32, device="cuda"); q = torch.empty(len(pairs) * n_feat, dtype=torch.int32, device="cuda")
t = torch.empty_like(q); d = torch.empty(len(pairs) * n_feat, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
only_tc = len(sys.argv) > 2 and sys.argv[2] == "tc"
for impl, name in ((2, "tcgen05"),) if only_tc else ((1, "simt"), (2, "tcgen05")):
    run = lambda: ms.match_pairs_device(pairs, cnt.data_ptr(), q.data_ptr(), t.data_ptr(), d.data_ptr(), n_feat, st, True, 0.9, -1, impl)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms_ = e0.elapsed_time(e1)
    print("%-8s K=%d: %d pairs in %.3f ms -> %.1f pairs/s, %.2f algorithmic TFLOP/s; matches/pair %.0f" % (name, k, len(pairs), ms_, len(pairs) / ms_ * 1e3, len(pairs) / ms_ * 1e3 * 2 * n_feat * n_feat * k / 1e12, cnt.float().mean().item()))
