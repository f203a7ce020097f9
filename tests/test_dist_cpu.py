"""Host logic of the multi-GPU paths on CPU: packing of the (score, index) word, the
shard maps, and world_size-2 runs over gloo of the 8-byte best-hypothesis all-reduce, the top-k
all-gather and the gather of a sharded batch's results."""
import os
import socket

import numpy as np
import pytest

from fccf_pcr_b200 import dist as D


def test_shard_maps_cover_everything_once():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
            owners = sorted(b for k in range(world) for b in D.pairs_of_rank(n, k, world))
            assert owners == list(range(n))
    assert D.pairs_of_rank(64, 3, 8) == list(range(3, 64, 8))


def _first_max(scores):
    """the reference's scan: best = 0-th candidate that is strictly greater than everything before it"""
    best, bi = -np.inf, -1
    for i, s in enumerate(scores):
        if s > best:
            best, bi = s, i
    return best, bi


def test_packed_word_orders_like_the_reference_scan():
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 40))
        s = rng.choice(np.array([0.0, -0.0, 0.25, 0.5, 0.5, 1.0, 1e-30, 3.5, np.nan], np.float32), n)
        p = D.pack_score_index(s, np.arange(n))
        sc, idx = D.unpack_score_index(p.max())
        best, bi = _first_max(np.where(np.isnan(s), -np.inf, s))
        if bi >= 0 and np.isfinite(best):
            assert int(idx) == bi and float(sc) == float(best)
    s = np.array([-3.0, -1.0, -2.0], np.float32)
    assert int(D.unpack_score_index(D.pack_score_index(s, np.arange(3)).max())[1]) == 1
    sc, idx = D.unpack_score_index(D.pack_score_index(np.float32(1.5), 123456))
    assert float(sc) == 1.5 and int(idx) == 123456
    assert np.isnan(D.unpack_score_index(D.pack_score_index(np.float32(np.nan), 5))[0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        scores = rng.choice(np.linspace(0, 1, 17).astype(np.float32), 1001)       # many ties
        lo, hi = D.shard_range(len(scores), rank, world)
        best = D.allreduce_best(D.local_best(scores[lo:hi], lo))
        sc, idx = D.unpack_score_index(best)
        top_s, top_i = D.allgather_topk(scores[lo:hi], lo, 5)
        # sharded batch: rank r "registers" pairs r, r+world, ... (result = a matrix made from the pair id)
        n_pairs = 7
        mine = D.pairs_of_rank(n_pairs, rank, world)
        T_local = np.stack([np.full((4, 4), float(b), np.float32) for b in mine]) if mine else np.zeros((0, 4, 4), np.float32)
        allT = D.gather_transforms(T_local, n_pairs)
        q.put((rank, float(sc), int(idx), top_s.tolist(), top_i.tolist(), allT[:, 0, 0].tolist(), scores.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_world_size_n_over_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scores = np.asarray(res[0][6], np.float32)
    best, bi = _first_max(scores)
    order = sorted(range(len(scores)), key=lambda i: (-scores[i], i))[:5]
    for rank, sc, idx, top_s, top_i, col, _ in res:
        assert sc == float(best) and idx == bi                     # every rank learns the same global best
        assert top_i == order and top_s == [float(scores[i]) for i in order]
        assert col == [float(b) for b in range(7)]
