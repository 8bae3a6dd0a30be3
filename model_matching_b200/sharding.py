"""Hypothesis sharding across ranks and the deterministic merge of per-rank top-K records.

The hot path shards by hypothesis: every rank holds a replica of the scene
index and the model tables, scores a contiguous block of the hypothesis list, reduces it to K
(global index, lcp) records on the device, and ONE all-gather collects the K records of every
rank.  The merge orders by (lcp descending, global index ascending), so its head equals the
reference's first-strict-maximum rule (src/stocs.cpp:994) on the whole list.
Pure torch / numpy host logic: runs on CPU (gloo) in the tests and on NCCL in bench.py.
"""
import numpy as np


def shard_range(H, rank, world):
    """Contiguous block [lo, hi) of ceil(H / world) hypotheses owned by `rank`."""
    per = -(-H // world)
    lo = min(H, rank * per)
    return lo, min(H, lo + per)


def merge_topk(idx, val, K):
    """idx (int64, -1 = empty) and val (float32) of all ranks, any shape -> best K records."""
    idx = np.asarray(idx).reshape(-1).astype(np.int64)
    val = np.asarray(val).reshape(-1).astype(np.float32)
    keep = (idx >= 0) & (val > 0)
    idx, val = idx[keep], val[keep]
    order = np.lexsort((idx, -val.astype(np.float64)))[:K]
    out_i = np.full(K, -1, np.int64)
    out_v = np.zeros(K, np.float32)
    out_i[:order.size] = idx[order]
    out_v[:order.size] = val[order]
    return out_i, out_v


def best_of(idx, val):
    """(best_index, best_lcp) with the reference's convention: (-1, 0.0) when nothing scored > 0."""
    i, v = merge_topk(idx, val, 1)
    return int(i[0]), float(v[0])


def all_gather_topk(dist, local_idx, local_val):
    """One all-gather per tensor of the K local records (torch tensors on the rank's device)."""
    import torch
    world = dist.get_world_size()
    gi = torch.empty(world * local_idx.numel(), dtype=local_idx.dtype, device=local_idx.device)
    gv = torch.empty(world * local_val.numel(), dtype=local_val.dtype, device=local_val.device)
    dist.all_gather_into_tensor(gi, local_idx.contiguous())
    dist.all_gather_into_tensor(gv, local_val.contiguous())
    return gi, gv
