"""Per-stage device times of one registration (fccf_timing.stage_ms), for profiling runs."""
import os
os.environ.setdefault("FCCF_STAGE_EVENTS", "1")      # stage_ms[1..6] wanted
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fccf_pcr_b200 import Context, scenes

STAGES = ["voxelgrid(main)", "voxelgrid(pipeline)", "planes", "hypotheses", "cluster", "quick_verify", "fine_verify+fuse"]


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "indoor"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    leaf = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    prm = {}
    for kv in sys.argv[6:]:
        k, v = kv.split("=")
        prm[k] = float(v)
    src, tar, _ = scenes.make_pair(kind, n, seed)
    ctx = Context(0, **prm)
    for r in range(reps):
        ctx.register(src, tar, leaf)
        tm = ctx.timing
        print("run %d: total %.3f ms (h2d %.3f, pipeline %.3f, d2h %.3f), %d launches | " % (r, tm.total_ms, tm.h2d_ms, tm.pipeline_ms, tm.d2h_ms, tm.n_launches) +
              " ".join("%s %.3f" % (s, tm.stage_ms[i]) for i, s in enumerate(STAGES)))
    pr = ctx.blob("prof")
    print("cluster_kernel(type 0) cycles: neighbour lists %d | seeding %d (%d rounds) | seeds+sizes %d | sort %d | walk %d | small centres %d | big centres %d" % (
        pr[9] - pr[0], pr[1] - pr[9], pr[8], pr[2] - pr[1], pr[3] - pr[2], pr[4] - pr[3], pr[5] - pr[4], pr[6] - pr[5]))
    print("grow_faces(cloud 1) cycles: stage1 %d | stage2 %d | range_face %d | select+theta %d" % (pr[17] - pr[16], pr[18] - pr[17], pr[19] - pr[18], pr[20] - pr[19]))
    print("n_hyp", ctx.blob("n_hyp"), "n_centres", ctx.blob("n_centres"), "P", len(ctx.blob("vg2_cnt1")), len(ctx.blob("vg2_cnt2")),
          "V", len(ctx.blob("vox_cnt1")), "Vp", len(ctx.blob("pvox1")) // 7, "S", len(ctx.blob("sub1")) // 3, len(ctx.blob("sub2")) // 3)


if __name__ == "__main__":
    main()
