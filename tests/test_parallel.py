"""world_size-2 gloo test of the pair-sharding + gather plumbing (host logic; the matcher itself is
replaced by the CPU oracle here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mavmap_b200 import parallel, synthetic


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_match_fn(desc):
    from oracle import orc

    def fn(pairs):
        off, qs, ts, ds = [0], [], [], []
        for i, j in pairs:
            q, t, d = orc.match_pair(desc[i], desc[j], ratio_test=True, max_ratio=0.9)
            qs.append(q); ts.append(t); ds.append(d); off.append(off[-1] + len(q))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        return np.array(off, np.int64), cat(qs, np.int32), cat(ts, np.int32), cat(ds, np.float32)
    return fn


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc, _ = synthetic.make_descriptors(5, 120, 32, seed=11)
    pairs = parallel.all_pairs(5)
    off, q, t, d = parallel.match_pairs_sharded(_oracle_match_fn(desc), pairs)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), off=off, q=q, t=t, d=d)
    dist.destroy_process_group()


def test_shard_pairs_partition():
    for n, w in [(10, 2), (7, 4), (3, 8), (0, 2)]:
        owned = [parallel.shard_pairs(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(owned).tolist()) == list(range(n))
    assert parallel.all_pairs(4).tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]


def test_two_rank_gather_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    desc, _ = synthetic.make_descriptors(5, 120, 32, seed=11)
    pairs = parallel.all_pairs(5)
    off1, q1, t1, d1 = parallel.match_pairs_sharded(_oracle_match_fn(desc), pairs)      # world = 1 path
    for r in range(2):
        g = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(g["off"], off1) and np.array_equal(g["q"], q1) and np.array_equal(g["t"], t1) and np.array_equal(g["d"], d1)
    assert off1[-1] > 100


def _allreduce_cb_worker(rank, world, port, q):
    import ctypes as C
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    from mavmap_b200.parallel import make_allreduce_callback
    cb = make_allreduce_callback(device=False)          # the mm_allreduce_fn of the sharded BA session, on a host buffer
    buf = np.arange(10, dtype=np.float64) * (rank + 1)
    rc = cb(None, buf.ctypes.data_as(C.c_void_p), 10, None)
    q.put((rank, rc, buf.copy()))
    dist.destroy_process_group()


def test_allreduce_callback_two_ranks_gloo():
    """the one collective primitive the sharded BA session asks of the host (mm_allreduce_fn): in-place sum over the ranks"""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_allreduce_cb_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps: p.join(timeout=60)
    for rank, rc, buf in out:
        assert rc == 0
        np.testing.assert_array_equal(buf, np.arange(10, dtype=np.float64) * 3)
