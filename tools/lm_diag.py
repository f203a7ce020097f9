"""Where do the GPU's and the oracle's refinements differ, and how does scipy judge those cases?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
from oracle.oracle import Oracle
import lm_check

for kind, n, seed, leaf, prm in (("indoor", 20000, 7, 0.1, {}), ("indoor", 50000, 1, 0.1, {}), ("indoor", 200000, 2, 0.2, {}), ("indoor_rough", 200000, 5, 0.2, {})):
    src, tar, _ = scenes.make_pair(kind, n, seed)
    o = Oracle(**prm); o.register(src, tar, leaf)
    c = fccf.Context(0, **prm)
    p1 = o.blob("face_plane1").reshape(-1, 7); p2 = o.blob("face_plane2").reshape(-1, 7)
    tot = bad = 0
    for t in range(3):
        cen = o.blob("centre%d" % t).reshape(-1, 7)
        if not len(cen):
            continue
        Ts = []
        for q in cen:
            T = np.eye(4, dtype=np.float32); T[:3, :3] = o.quat_to_matrix(q[:4]); T[:3, 3] = q[4:7]; Ts.append(T)
        Ts = np.stack(Ts)
        sc, Tr, npair, pairs, iters = c.quick_verify(Ts, p1, p2)
        for k in range(len(Ts)):
            so, To, po, io = o.quick_verify(Ts[k], p1, p2)
            tot += 1
            de, dt = scenes.rotation_error_deg(Tr[k], To), scenes.translation_error(Tr[k], To)
            if de > 0.01 or dt > 1e-3:
                bad += 1
                rows = lm_check.problem(Ts[k], p1, p2, po)
                Ts_, cs, ratio = lm_check.solve(Ts[k], rows)
                print("  %s type %d centre %d: gpu-vs-oracle %.4f deg %.5f m; iters gpu %d oracle %d; pairs %d; cost gpu %.3e oracle %.3e scipy %.3e; eig ratio %.2e; gpu-vs-scipy %.4f deg %.5f m; oracle-vs-scipy %.4f deg %.5f m" % (
                    kind, t, k, de, dt, iters[k], io, len(po), lm_check.cost_of(Tr[k], Ts[k], rows), lm_check.cost_of(To, Ts[k], rows), cs, ratio,
                    scenes.rotation_error_deg(Tr[k], Ts_), scenes.translation_error(Tr[k], Ts_), scenes.rotation_error_deg(To, Ts_), scenes.translation_error(To, Ts_)))
    print("%s %d seed %d: %d of %d refined centres differ by more than the bar" % (kind, n, seed, bad, tot), flush=True)
    c.close()
