"""Multi-GPU plumbing for the part of the hot path that shards: image-pair matching.

Pairs are independent units: every rank holds the descriptors, takes a strided
slice of the pair list, matches it on its own GPU, and the variable-length match lists are
gathered once at the end (two collectives: counts, then the packed lists) over NCCL/NVLink
(`gloo` in the CPU tests).  There is no data-path collective inside the matching itself.
Bundle adjustment shards by 3-D point (mm_ba_session_create_sharded): the library asks the host for ONE primitive, an
in-place sum of doubles over the ranks, through a callback; `make_allreduce_callback` provides it on top of
torch.distributed (NCCL over NVLink on the GPUs, `gloo` on host buffers in the CPU tests).
"""
import ctypes as C

import numpy as np


class _DevF64:
    """zero-copy view of `n` doubles at a device pointer for torch.as_tensor"""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def make_allreduce_callback(group=None, device=True):
    """The `mm_allreduce_fn` of include/mavmap_b200.h: sums `count` doubles at `buf` over all ranks, in place, ordered on
    `stream`.  device=False treats `buf` as a host pointer (CPU tests of the plumbing with gloo)."""
    import torch
    import torch.distributed as dist
    from ._abi import ALLREDUCE_FN

    def cb(user, buf, count, stream):
        try:
            if count <= 0:
                return 0
            if device:
                t = torch.as_tensor(_DevF64(buf, count), device="cuda")
                if stream:
                    with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                else:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                a = np.ctypeslib.as_array((C.c_double * int(count)).from_address(int(buf)))
                t = torch.from_numpy(a)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:          # never unwind through the C frame
            import sys
            print("all-reduce callback failed: %r" % (e,), file=sys.stderr)
            return 1
    return ALLREDUCE_FN(cb)


def shard_pairs(n_pairs, rank, world):
    """Indices of the pairs owned by `rank` (strided so that neighbouring, similarly sized pairs spread out)."""
    return np.arange(rank, n_pairs, world, dtype=np.int64)


def all_pairs(n_images):
    i, j = np.triu_indices(n_images, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)


def gather_match_lists(pair_idx, off, q, t, d, n_pairs_total, device=None, group=None):
    """All ranks call this with their local results (pair_idx [m] global pair ids, off [m+1], q/t/d packed).
    Returns on every rank: (off_all [n_pairs_total+1], q_all, t_all, d_all) ordered by global pair id."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        order = np.argsort(pair_idx)
        cnt = np.diff(off)[order]
        off_all = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        sel = np.concatenate([np.arange(off[k], off[k + 1]) for k in order]) if len(order) else np.zeros(0, np.int64)
        return off_all, q[sel], t[sel], d[sel]
    dev = device if device is not None else torch.device("cpu")
    # 1. per-pair counts, scattered into a dense [n_pairs_total] vector and summed
    counts = torch.zeros(n_pairs_total, dtype=torch.int64, device=dev)
    counts[torch.as_tensor(pair_idx, dtype=torch.int64, device=dev)] = torch.as_tensor(np.diff(off), dtype=torch.int64, device=dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    # 2. packed lists, padded to the largest rank payload
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[dist.get_rank(group)] = len(q)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
    cap = int(sizes.max().item())
    payload = torch.zeros((3, max(cap, 1)), dtype=torch.int32, device=dev)
    payload[0, :len(q)] = torch.as_tensor(q, dtype=torch.int32, device=dev)
    payload[1, :len(t)] = torch.as_tensor(t, dtype=torch.int32, device=dev)
    payload[2, :len(d)] = torch.as_tensor(np.asarray(d, dtype=np.float32).view(np.int32), dtype=torch.int32, device=dev)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    counts_h = counts.cpu().numpy()
    off_all = np.concatenate([[0], np.cumsum(counts_h)]).astype(np.int64)
    q_all = np.empty(off_all[-1], np.int32); t_all = np.empty(off_all[-1], np.int32); d_all = np.empty(off_all[-1], np.float32)
    for r in range(world):
        g = gathered[r].cpu().numpy()
        owned = shard_pairs(n_pairs_total, r, world)
        pos = 0
        for p in owned:
            c = int(counts_h[p])
            q_all[off_all[p]:off_all[p] + c] = g[0, pos:pos + c]
            t_all[off_all[p]:off_all[p] + c] = g[1, pos:pos + c]
            d_all[off_all[p]:off_all[p] + c] = g[2, pos:pos + c].view(np.float32)
            pos += c
    return off_all, q_all, t_all, d_all


def match_pairs_sharded(match_fn, pairs, device=None, group=None):
    """match_fn(pairs_subset [m,2]) -> (off [m+1], q, t, d).  Shards by rank and gathers."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    mine = shard_pairs(len(pairs), rank, world)
    off, q, t, d = match_fn(pairs[mine])
    return gather_match_lists(mine, np.asarray(off), q, t, d, len(pairs), device, group)
