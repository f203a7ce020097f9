"""Scoring-kernel throughput on the 200k indoor pair's leftover clouds.  python tools/score_bench.py [H]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import bench
H = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
src, tar, _ = scenes.make_pair("indoor", 200000, 100)
c = fccf.Context(0)
T0 = c.register(src, tar, 0.2)
s1 = c.blob("sub1").reshape(-1, 3).copy(); s2 = c.blob("sub2").reshape(-1, 3).copy()
hyps = bench.perturbed_hypotheses(T0, H, 1234)
sc, ms = c.score_hypotheses_bench(hyps, s1, s2, 10)
print("H=%d n1=%d n2=%d kernel %.4f ms -> %.1f M hypotheses/s, %.1f G point-hyps/s; checksum %.6f max %.6f" % (H, len(s1), len(s2), ms, H / ms / 1e3, H * len(s2) / ms / 1e6, float(sc.sum()), float(sc.max())))
