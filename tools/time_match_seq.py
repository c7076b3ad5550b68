"""cfg3 at its stated size: a sequence of N images x 5000 SURF-64 descriptors RESIDENT on one GPU (2000 images: 2.56 GB of descriptors
+ 7.7 GB of TF32 operand copies) and a slice of its all-pairs matching (pair-sharded work of one rank).
usage: python tools/time_match_seq.py [n_images=2000] [n_pairs=4800]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mavmap_b200 import synthetic
from mavmap_b200.matching import MatchSet
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 4800
n_feat, k = 5000, 64
t = time.time(); desc, _ = synthetic.make_descriptors(n_img, n_feat, k, seed=0xF00D + 3); print("generated %d images in %.1f s" % (n_img, time.time() - t), flush=True)
t = time.time(); ms = MatchSet(desc, None); torch.cuda.synchronize(); print("resident in %.2f s (%.2f GB of descriptors)" % (time.time() - t, desc.nbytes / 1e9), flush=True)
# a strided slice of the upper-triangular pair matrix (what rank r of R takes), not only neighbours
rng = np.random.default_rng(0)
i = rng.integers(0, n_img - 1, n_pairs); j = np.minimum(i + 1 + rng.integers(0, 40, n_pairs), n_img - 1)
pairs = [(int(a), int(b)) for a, b in zip(i, j) if a != b]
chunk = 240
cnt = torch.zeros(chunk, dtype=torch.int32, device="cuda"); q = torch.empty(chunk * n_feat, dtype=torch.int32, device="cuda"); tt = torch.empty_like(q); d = torch.empty(chunk * n_feat, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
run = lambda pp: ms.match_pairs_device(pp, cnt.data_ptr(), q.data_ptr(), tt.data_ptr(), d.data_ptr(), n_feat, st, True, 0.9, -1)
run(pairs[:chunk]); torch.cuda.synchronize()                  # first use prepares the TF32 operands of the whole sequence
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tot = 0
for c0 in range(0, len(pairs), chunk):
    pp = pairs[c0:c0 + chunk]; run(pp); tot += int(cnt[:len(pp)].sum().item())
e1.record(); torch.cuda.synchronize()
msec = e0.elapsed_time(e1)
print("%d pairs of a resident %d-image sequence in %.1f ms -> %.0f pairs/s (%.0f matches per pair); device memory in use %.1f GB" % (
    len(pairs), n_img, msec, len(pairs) / msec * 1e3, tot / len(pairs), (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9))
