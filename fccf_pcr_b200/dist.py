"""Multi-GPU host logic of the FCCF-PCR path: one process per GPU, torch.distributed
for the plumbing (NCCL over NVLink on the B200 box, gloo in the CPU tests).

Only two things shard (BASELINE.json north_star):
  * batches of independent scan pairs: pair b -> rank b mod N, no data-path collective; the 4x4
    results (64 B each) are gathered at the end;
  * hypothesis scoring: contiguous ranges of the ordered hypothesis list per rank, the static voxel
    hash and the moving cloud replicated, and ONE 8-byte all-reduce (max) of a packed
    (score, index) word for the global best.  Smallest index wins ties — the reference's strict `>`
    first-maximum scans (FCCF.cpp:1559, 749).
Everything else of a single registration runs on one GPU (order-dependent greedy stages).
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n ordered items for `rank`; the first n % world ranks get one more."""
    q, r = divmod(int(n), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def pairs_of_rank(n_pairs, rank, world):
    """Indices of the scan pairs `rank` registers (pair b -> rank b mod world)."""
    return list(range(rank, int(n_pairs), int(world)))


def pack_score_index(score, index):
    """int64 word whose signed order is (score ascending, index descending): the maximum over a set is
    the highest score and, among equal scores, the smallest index.  NaN scores rank below everything
    (every `score > best` test of the reference is false for NaN).  Same layout as fccf_score_best."""
    s = np.asarray(score, np.float32) + np.float32(0.0)      # -0.0 -> +0.0 (they compare equal)
    b = s.view(np.int32).astype(np.int64)
    key = np.where(b >= 0, b, b ^ 0x7FFFFFFF)            # order-preserving signed int of a float
    key = np.where(np.isnan(s), np.int64(-(1 << 31)), key)
    idx = np.asarray(index, np.int64)
    return (key << 32) | (np.int64(0xFFFFFFFF) - idx)


def unpack_score_index(packed):
    p = np.asarray(packed, np.int64)
    key = (p >> 32).astype(np.int64)
    idx = np.int64(0xFFFFFFFF) - (p & np.int64(0xFFFFFFFF))
    bits = np.where(key >= 0, key, key ^ 0x7FFFFFFF).astype(np.int32)
    score = bits.view(np.float32) if bits.ndim else np.array(bits, np.int32).view(np.float32)
    score = np.where(key == -(1 << 31), np.float32(np.nan), score)
    return score, idx


def local_best(scores, lo=0):
    """Packed best of a rank's score slice whose first element has global index `lo`."""
    scores = np.asarray(scores, np.float32)
    if len(scores) == 0:
        return np.int64(-(1 << 63))
    return pack_score_index(scores, lo + np.arange(len(scores))).max()


def _dist():
    import torch.distributed as dist

    return dist


def allreduce_best(packed, device=None):
    """Global maximum of one packed (score, index) word per rank: the 8-byte collective of §8e."""
    import torch

    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.int64(packed)
    t = torch.tensor([int(packed)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return np.int64(t.item())


def allgather_topk(scores, lo, k, device=None):
    """Global top-k (score descending, index ascending among ties) from per-rank slices: every rank
    contributes its k best packed words (k x 8 B), all-gathered and merged."""
    import torch

    dist = _dist()
    scores = np.asarray(scores, np.float32)
    packed = np.sort(pack_score_index(scores, lo + np.arange(len(scores))))[::-1][:k] if len(scores) else np.zeros(0, np.int64)
    mine = np.full(k, -(1 << 63), np.int64)
    mine[:len(packed)] = packed
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(mine).to(device if device is not None else "cpu")
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        allp = np.concatenate([o.cpu().numpy() for o in out])
    else:
        allp = mine
    allp = np.sort(allp[allp != -(1 << 63)])[::-1][:k]
    return unpack_score_index(allp)


def gather_transforms(T_local, n_pairs, device=None):
    """Results of a sharded batch: rank r holds the 4x4 of pairs r, r+world, ...; every rank gets all."""
    import torch

    dist = _dist()
    T_local = np.ascontiguousarray(T_local, np.float32).reshape(-1, 16)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return T_local.reshape(-1, 4, 4)
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (int(n_pairs) + world - 1) // world
    buf = np.zeros((per, 16), np.float32)
    buf[:len(T_local)] = T_local
    t = torch.from_numpy(buf).to(device if device is not None else "cpu")
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    allT = np.zeros((int(n_pairs), 16), np.float32)
    for r in range(world):
        idx = pairs_of_rank(n_pairs, r, world)
        allT[idx] = out[r].cpu().numpy()[:len(idx)]
    return allT.reshape(-1, 4, 4)


def gather_blocks(T_local, n_total, device=None):
    """Results of a batch cut into CONTIGUOUS blocks (shard_range): rank r holds the 4x4 of pairs [lo_r, hi_r); every
    rank gets all of them in order."""
    import torch

    dist = _dist()
    T_local = np.ascontiguousarray(T_local, np.float32).reshape(-1, 16)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return T_local.reshape(-1, 4, 4)
    world = dist.get_world_size()
    per = (int(n_total) + world - 1) // world
    buf = np.zeros((per, 16), np.float32)
    buf[:len(T_local)] = T_local
    t = torch.from_numpy(buf).to(device if device is not None else "cpu")
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    allT = np.zeros((int(n_total), 16), np.float32)
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        allT[lo:hi] = out[r].cpu().numpy()[:hi - lo]
    return allT.reshape(-1, 4, 4)


def register_pairs_sharded(ctx, srcs, tars, leaf, device=None):
    """BASELINE config 4: a batch of independent pairs over the ranks of the job.  `srcs`/`tars` list ALL
    pairs on every rank (or None for pairs this rank does not own)."""
    dist = _dist()
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = pairs_of_rank(len(srcs), rank, world)
    T = ctx.register_batch([srcs[b] for b in mine], [tars[b] for b in mine], leaf) if mine else np.zeros((0, 4, 4), np.float32)
    return gather_transforms(T, len(srcs), device)


def sharded_best_hypothesis(ctx, hyps, s1, s2, device=None):
    """BASELINE config 3: score an ordered hypothesis list over the ranks of the job (static hash and
    moving cloud replicated) and return (best score, global index) after one 8-byte all-reduce."""
    dist = _dist()
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    hyps = np.ascontiguousarray(hyps, np.float32).reshape(-1, 16)
    lo, hi = shard_range(len(hyps), rank, world)
    scores = ctx.score_hypotheses(hyps[lo:hi], s1, s2)
    packed = ctx.score_best(lo) if hi > lo else np.int64(-(1 << 63))      # block-reduced argmax on the device
    score, idx = unpack_score_index(allreduce_best(packed, device))
    return float(score), int(idx), scores
