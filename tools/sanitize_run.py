"""Small single + batched registrations for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fccf_pcr_b200 import Context, scenes

pairs = [scenes.make_pair("indoor", 20000, s)[:2] for s in (7, 8, 9)]
c = Context(0, batch_lanes=2)
T = c.register(pairs[0][0], pairs[0][1], 0.1)
Tb = c.register_batch([p[0] for p in pairs], [p[1] for p in pairs], 0.1)
assert np.array_equal(Tb[0], T)
s1 = c.blob("sub1").reshape(-1, 3); s2 = c.blob("sub2").reshape(-1, 3)
sc = c.score_hypotheses(np.tile(Tb[0], (64, 1, 1)), s1, s2)
print("ok", T[0], float(sc[0]))
c.close()
