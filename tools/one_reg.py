"""One warm registration of the 200k indoor pair (profiling target).  python tools/one_reg.py [repeats] [pairs]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pairs = [scenes.make_pair("indoor", 200000, 100 + i)[:2] for i in range(min(npairs, 8))]
pairs = (pairs * npairs)[:npairs]
c = fccf.Context(0)
for _ in range(rep):
    if npairs == 1:
        T = c.register(pairs[0][0], pairs[0][1], 0.2)
    else:
        T = c.register_batch([p[0] for p in pairs], [p[1] for p in pairs], 0.2)
print("total %.3f ms, stage0 %.3f, launches %d" % (c.timing.total_ms, c.timing.stage_ms[0], c.timing.n_launches), c.blob("vg_fast"))
pr = c.blob("prof")[22:32]
names = ["A loop", "zero+reduce+sync1+setup", "B loop", "prefix+sync2+sync3+qstart", "B2 loop", "fence+sync4", "C loop", "D (sync5, masks, out)", "closing sync"]
print("vg_fast phases of cluster 0 / CTA 0 (cycles):")
for k, nm in enumerate(names):
    print("  %-28s %8d" % (nm, pr[k + 1] - pr[k]))
print("  total %d" % (pr[9] - pr[0]))
