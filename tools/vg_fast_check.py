"""Cluster VoxelGrid (voxelgrid_fast.cu) against the generic kernels and the oracle, and its stage time.
Run on a GPU box:  python tools/vg_fast_check.py [pairs]"""
import os
os.environ.setdefault("FCCF_STAGE_EVENTS", "1")      # stage_ms[1..6] wanted
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
from oracle.oracle import Oracle


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rng = np.random.default_rng(4)
    c = fccf.Context(0)
    o = Oracle()
    src, tar, _ = scenes.make_pair("indoor", 200000, 100)
    clouds = [("indoor200k-src", src, 0.2), ("indoor200k-tar", tar, 0.2), ("tiny7", (rng.normal(size=(7, 3))).astype(np.float32), 0.5),
              ("n1", np.ones((1, 3), np.float32), 0.1), ("n0", np.zeros((0, 3), np.float32), 0.1),
              ("gauss5000", (rng.normal(size=(5000, 3)) * [4, 3, 1]).astype(np.float32), 0.25),
              ("coherent", np.repeat((rng.normal(size=(300, 3)) * [4, 3, 1]).astype(np.float32), 400, axis=0), 0.3),
              ("onecell", (rng.uniform(0.01, 0.09, (70000, 3))).astype(np.float32), 0.1)]
    nan = src[:50000].copy(); nan[::97] = np.nan; nan[5, 1] = np.inf
    clouds.append(("nan50k", nan, 0.2))
    ok = True
    for name, pts, leaf in clouds:
        os.environ.pop("FCCF_VG_GENERIC", None)
        s0 = c.blob("vg_fast").copy()
        a = c.voxelgrid(pts, leaf)
        s1 = c.blob("vg_fast").copy()
        os.environ["FCCF_VG_GENERIC"] = "1"
        b = c.voxelgrid(pts, leaf)
        os.environ.pop("FCCF_VG_GENERIC", None)
        r = o.voxelgrid(pts, leaf)
        same_g = all(x.shape == y.shape and np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
        same_o = all(x.shape == y.shape and np.array_equal(x, y, equal_nan=True) for x, y in zip(a, r))
        print("%-16s n=%7d leaf=%.2f cells=%6d fast_runs+%d misses+%d  == generic: %s  == oracle: %s" % (name, len(pts), leaf, len(a[0]), s1[0] - s0[0], s1[1] - s0[1], same_g, same_o), flush=True)
        ok = ok and same_g and same_o
    # whole registration, fast vs generic context, bit-identical
    os.environ["FCCF_NO_VG_FAST"] = "1"
    cg = fccf.Context(0)
    os.environ.pop("FCCF_NO_VG_FAST")
    Tf = c.register(src, tar, 0.2); tf = c.timing.total_ms; lf = c.timing.n_launches
    Tg = cg.register(src, tar, 0.2); tg = cg.timing.total_ms; lg = cg.timing.n_launches
    for _ in range(3):
        c.register(src, tar, 0.2); cg.register(src, tar, 0.2)
    print("single pair: fast %.3f ms (%d launches, stage0 %.3f, stage1 %.3f) | generic %.3f ms (%d launches, stage0 %.3f, stage1 %.3f) | same bits: %s" % (
        c.timing.total_ms, lf, c.timing.stage_ms[0], c.timing.stage_ms[1], cg.timing.total_ms, lg, cg.timing.stage_ms[0], cg.timing.stage_ms[1], np.array_equal(Tf, Tg)))
    for nm in ("vg1_cell1", "vg1_cnt2", "vg1_xyz1", "vg2_xyz2", "vg2_cell1", "vg2_cnt2"):
        ok = ok and np.array_equal(c.blob(nm), cg.blob(nm))
    ok = ok and np.array_equal(Tf, Tg)
    # batch of pairs, device-resident
    import torch
    pairs = [scenes.make_pair("indoor", 200000, 100 + i)[:2] for i in range(min(npairs, 16))]
    pairs = (pairs * ((npairs + len(pairs) - 1) // len(pairs)))[:npairs]
    d = [(torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda()) for s, t in pairs]
    sp = [x.data_ptr() for x, _ in d]; tp = [y.data_ptr() for _, y in d]
    ns = [len(s) for s, _ in pairs]; nt = [len(t) for _, t in pairs]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for nm, cx in (("fast", c), ("generic", cg)):
        for _ in range(3):
            T = cx.register_batch_device(sp, ns, tp, nt, 0.2)
        st = np.zeros(8); tot = 0.0
        for _ in range(5):
            flush.fill_(1); torch.cuda.synchronize()
            T = cx.register_batch_device(sp, ns, tp, nt, 0.2)
            st += np.array(list(cx.timing.stage_ms)); tot += cx.timing.total_ms
        print("%-8s batch of %d: total %.3f ms, stages %s" % (nm, npairs, tot / 5, np.round(st[:7] / 5, 3)), flush=True)
        if nm == "fast":
            Tfast = T.copy()
        else:
            ok = ok and np.array_equal(Tfast, T)
            print("batch results identical:", np.array_equal(Tfast, T))
    print("vg_fast stats:", c.blob("vg_fast"))
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
