"""One warm registration of the 2M + 2M outdoor pair of BASELINE config 3 (profiling target).  python tools/one_reg_outdoor.py [repeats] [points]"""
import os, sys
os.environ.setdefault("FCCF_STAGE_EVENTS", "1")      # stage_ms[1..6] wanted
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
src, tar, _ = scenes.make_pair("outdoor", n, 3)
c = fccf.Context(0, face_voxel_size=4.0, fine_verify_voxel_size=2.0)
for _ in range(rep):
    c.register(src, tar, 0.5)
tm = c.timing
print("total %.3f ms (h2d %.3f, pipeline %.3f), launches %d, stages %s" % (tm.total_ms, tm.h2d_ms, tm.pipeline_ms, tm.n_launches, [round(x, 3) for x in tm.stage_ms[:7]]))
