"""Final-transform and integer-stage parity against the oracle over many seeds (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fccf_pcr_b200 import Context, scenes
from oracle.oracle import Oracle

NAMES = ["vg2_cnt1", "vox_cnt1", "merge_label1", "merge_label2", "base1", "base2", "matches", "n_hyp", "n_centres", "cluster_num", "top_centre0", "top_centre1", "top_centre2"]


def main():
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    worst = (0.0, 0.0); bad = 0
    for kind, n, leaf, prm, seeds in [("indoor", 200000, 0.2, {}, range(100, 100 + nseeds)), ("indoor", 50000, 0.1, {}, range(200, 200 + nseeds // 2)),
                                      ("outdoor", 300000, 0.5, dict(face_voxel_size=4.0, fine_verify_voxel_size=2.0), range(3, 3 + nseeds // 2))]:
        c = Context(0, **prm); o = Oracle(**prm)
        for s in seeds:
            src, tar, _ = scenes.make_pair(kind, n, s)
            Tg = c.register(src, tar, leaf); To = o.register(src, tar, leaf)
            ints = [nm for nm in NAMES if not np.array_equal(c.blob(nm), o.blob(nm))]
            if np.isnan(To).any() or not np.any(To[:3, :3]):
                ok = np.array_equal(np.isnan(Tg), np.isnan(To)); r = t = 0.0
            else:
                r, t = scenes.rotation_error_deg(Tg, To), scenes.translation_error(Tg, To)
                ok = r <= 0.01 and t <= 1e-3
            worst = (max(worst[0], r), max(worst[1], t))
            if ints or not ok:
                bad += 1
                print("MISMATCH %s n=%d seed=%d: int blobs %s, rot %.5f deg, trans %.6f m" % (kind, n, s, ints, r, t))
        c.close()
    print("parity sweep: %d mismatching pairs; worst rotation %.5f deg, worst translation %.6f m" % (bad, worst[0], worst[1]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
