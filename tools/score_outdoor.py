"""Scoring throughput on the leftover clouds of BASELINE config 3's pair (2M + 2M outdoor points, fine-verify voxel 2 m:
the table does not fit shared memory, score_kernel runs).  python tools/score_outdoor.py [H] [points]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import bench
H = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
src, tar, _ = scenes.make_pair("outdoor", n, 3)
c = fccf.Context(0, face_voxel_size=4.0, fine_verify_voxel_size=2.0)
T0 = c.register(src, tar, 0.5)
s1 = c.blob("sub1").reshape(-1, 3).copy(); s2 = c.blob("sub2").reshape(-1, 3).copy()
hyps = bench.perturbed_hypotheses(T0 if np.isfinite(T0).all() else np.eye(4), H, 4321)
sc, ms = c.score_hypotheses_bench(hyps, s1, s2, 3)
print("H=%d n1=%d n2=%d kernel %.3f ms -> %.2f M hypotheses/s, %.1f G point-hyps/s; checksum %.6f max %.6f" % (H, len(s1), len(s2), ms, H / ms / 1e3, H * len(s2) / ms / 1e6, float(sc.sum()), float(sc.max())))
