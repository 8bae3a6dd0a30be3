"""Host-side mirror of the library's multi-GPU protocol (csrc/comm.cu).

The hot path shards by hypothesis: every rank holds a replica of the scene index and the model
tables, scores a contiguous block of the hypothesis list, reduces it to K 64-byte records
{lcp, inliers, global index, 3x4 transform} on the device, and ONE all-gather collects the K
records of every rank; the merge orders them by (lcp descending, global index ascending), so
record 0 equals the reference's first-strict-maximum rule (src/stocs.cpp:994) on the whole list.

The product implementation is C/CUDA behind the ABI (stocs_b200_comm_init,
stocs_b200_score_sharded*, stocs_b200_group_*).  This module restates the block split, the record
layout and the merge rule in numpy so that the protocol can be exercised on CPU (gloo, world_size 2,
tests/test_sharding.py) and so that GPU tests have an independent statement of the merge to compare
the kernel with.
"""
import numpy as np

# stocs_b200_record (include/stocs_b200.h)
RECORD = np.dtype([("lcp", np.float32), ("inliers", np.int32), ("index", np.int64), ("T", np.float32, (12,))])
assert RECORD.itemsize == 64


def shard_range(H, rank, world):
    """Contiguous block [lo, hi) of ceil(H / world) hypotheses owned by `rank` (stocs_b200_shard_range)."""
    per = -(-H // world)
    lo = min(H, rank * per)
    return lo, min(H, lo + per)


def empty_records(K):
    r = np.zeros(K, RECORD)
    r["index"] = -1
    return r


def local_records(lcp, inliers, T16, lo, K):
    """K best of one block as records (what reduce.cu's topk_kernel packs): lcp > 0 only,
    ordered by (lcp descending, index ascending), global index = lo + local index."""
    lcp = np.asarray(lcp, np.float32)
    order = np.lexsort((np.arange(lcp.size), -lcp.astype(np.float64)))
    order = order[lcp[order] > 0][:K]
    r = empty_records(K)
    n = order.size
    r["lcp"][:n] = lcp[order]
    r["inliers"][:n] = np.asarray(inliers, np.int32)[order]
    r["index"][:n] = order + lo
    T = np.asarray(T16, np.float32).reshape(-1, 4, 4)          # column-major 4x4 per hypothesis
    r["T"][:n] = T[order].transpose(0, 2, 1)[:, :3, :].reshape(n, 12)
    return r


def merge_records(records, K):
    """All ranks' records (any shape) -> the K best, the rule of comm.cu's merge_records_kernel."""
    rec = np.asarray(records).reshape(-1)
    keep = (rec["index"] >= 0) & (rec["lcp"] > 0)
    rec = rec[keep]
    order = np.lexsort((rec["index"], -rec["lcp"].astype(np.float64)))[:K]
    out = empty_records(K)
    out[:order.size] = rec[order]
    return out


def merge_topk(idx, val, K):
    """(index, lcp) arrays of all ranks -> best K pairs (same rule as merge_records)."""
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


def all_gather_records(dist, local):
    """ONE all-gather of the K local records (numpy RECORD array) over a torch.distributed group."""
    import torch
    world = dist.get_world_size()
    send = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).copy())
    recv = torch.empty(world * send.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    return recv.numpy().view(RECORD)
