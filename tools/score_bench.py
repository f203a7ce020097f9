"""Hypothesis-scoring kernel alone (for ncu): leftover clouds of one 200k pair, H perturbed hypotheses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from fccf_pcr_b200 import Context, scenes


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    npts = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
    leaf = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
    src, tar, _ = scenes.make_pair("indoor", npts, 100)
    ctx = Context(0)
    T0 = ctx.register(src, tar, leaf)
    s1 = ctx.blob("sub1").reshape(-1, 3).copy()
    s2 = ctx.blob("sub2").reshape(-1, 3).copy()
    hyps = bench.perturbed_hypotheses(T0, H, 1234)
    sc, ms = ctx.score_hypotheses_bench(hyps, s1, s2, rep)
    pts = len(s2) * H / (ms * 1e-3)
    print("H %d static %d moving %d kernel %.4f ms -> %.2f M hyp/s, %.1f G point-hyp/s, %.1f GB/s algorithmic (20 B/point)" % (H, len(s1), len(s2), ms, H / ms / 1e3, pts / 1e9, 20 * pts / 1e9))
    print("score[0] %.6f max %.6f argmax %d" % (sc[0], sc.max(), int(sc.argmax())))


if __name__ == "__main__":
    main()
