"""Multi-GPU check of the two sharded paths, one process per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py

(1) hypothesis scoring sharded over the ranks (static table + moving cloud replicated, one 8-byte
    all-reduce(max) of the packed (score, index) word) == the first maximum of one GPU scoring everything;
(2) a batch of independent pairs sharded pair b -> rank b mod N == the same pairs registered on one GPU.
Also times the sharded scoring (hypotheses/s over all ranks, device events, max over ranks)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import dist as fd
from fccf_pcr_b200 import scenes


def perturbed(T, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.tile(np.asarray(T, np.float32), (n, 1, 1))
    ang = np.radians(rng.uniform(-6, 6, n)); c, s = np.cos(ang), np.sin(ang)
    Rz = np.tile(np.eye(3), (n, 1, 1)); Rz[:, 0, 0] = c; Rz[:, 0, 1] = -s; Rz[:, 1, 0] = s; Rz[:, 1, 1] = c
    out[:, :3, :3] = (Rz @ np.asarray(T, float)[:3, :3]).astype(np.float32)
    out[:, :3, 3] += rng.uniform(-0.4, 0.4, (n, 3)).astype(np.float32)
    out[n // 3] = np.asarray(T, np.float32)
    return out


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 9472 * 4
    kind, leaf, prm = ("outdoor", 0.5, dict(face_voxel_size=4.0, fine_verify_voxel_size=2.0)) if npts >= 1000000 else ("indoor", 0.2, {})
    src, tar, Tgt = scenes.make_pair(kind, npts, 3)
    ctx = fccf.Context(local, **prm)
    T0 = ctx.register(src, tar, leaf)
    s1 = ctx.blob("sub1").reshape(-1, 3).copy(); s2 = ctx.blob("sub2").reshape(-1, 3).copy()
    hyps = perturbed(T0, H, 77)
    # (1) sharded scoring
    score, idx, mine = fd.sharded_best_hypothesis(ctx, hyps, s1, s2, dev)
    lo, hi = fd.shard_range(H, rank, world)
    t_ms = ctx.score_hypotheses_bench(hyps[lo:hi], s1, s2, 5)[1]
    t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok1 = True
    if rank == 0:
        full = ctx.score_hypotheses(hyps, s1, s2)
        s_ref, i_ref = fd.unpack_score_index(fd.local_best(full))
        ok1 = (int(i_ref) == idx) and (np.float32(s_ref) == np.float32(score)) and np.array_equal(full[lo:hi], mine)
        print("sharded scoring: %d hypotheses over %d GPU(s), static %d moving %d points: best index %d score %.6f (single-GPU scan: %d %.6f) %s; %.3f ms max over ranks -> %.1f M hypotheses/s"
              % (H, world, len(s1), len(s2), idx, score, int(i_ref), float(s_ref), "OK" if ok1 else "MISMATCH", float(t.item()), H / float(t.item()) / 1e3))
    # (2) sharded batch
    npairs = 2 * world + 1
    pairs = [scenes.make_pair("indoor", 20000, 40 + b) for b in range(npairs)]
    Tall = fd.register_pairs_sharded(ctx, [p[0] for p in pairs], [p[1] for p in pairs], 0.1, dev)
    ok2 = True
    if rank == 0:
        for b, p in enumerate(pairs):
            ok2 = ok2 and np.array_equal(Tall[b], ctx.register(p[0], p[1], 0.1), equal_nan=True)
        print("sharded batch: %d pairs over %d GPU(s) %s" % (npairs, world, "OK" if ok2 else "MISMATCH"))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not (ok1 and ok2):
        sys.exit(1)


if __name__ == "__main__":
    main()
