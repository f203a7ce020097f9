"""One batched registration sequence (G pairs in one group) for profiling: ncu launch lists of the batched kernels."""
import os
os.environ.setdefault("FCCF_STAGE_EVENTS", "1")      # stage_ms[1..6] wanted
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fccf_pcr_b200 import Context, scenes


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    leaf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    nuniq = min(G, 8)
    pairs = [scenes.make_pair("indoor", n, 100 + i) for i in range(nuniq)]
    srcs = [pairs[i % nuniq][0] for i in range(G)]
    tars = [pairs[i % nuniq][1] for i in range(G)]
    ctx = Context(0, batch_lanes=G)
    for r in range(reps):
        ctx.register_batch(srcs, tars, leaf)
        tm = ctx.timing
        print("rep %d: G=%d total %.3f ms (%.4f ms/registration), h2d %.3f, %d launches | stages " % (r, G, tm.total_ms, tm.total_ms / G, tm.h2d_ms, tm.n_launches) +
              " ".join("%.3f" % v for v in list(tm.stage_ms)[:7]))
    ctx.close()


if __name__ == "__main__":
    main()
