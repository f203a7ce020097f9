import os, sys
os.environ.setdefault("FCCF_STAGE_EVENTS", "1")      # stage_ms[1..6] wanted
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import torch
src, tar, _ = scenes.make_pair("indoor", 200000, 100)
ds, dt = torch.from_numpy(src).cuda(), torch.from_numpy(tar).cuda()
c = fccf.Context(0)
lat = []; st = np.zeros(8)
for k in range(12):
    c.register_device(ds.data_ptr(), len(src), dt.data_ptr(), len(tar), 0.2)
    if k >= 2:
        lat.append(c.timing.total_ms); st += np.array(list(c.timing.stage_ms))
it = [c.blob("qv_iters%d" % t) for t in range(3)]
print("env GROW=%s QVOCC=%s: latency %.3f ms median, stages %s ; refine iters %s; n_hyp %s centres %s" % (os.environ.get("FCCF_GROW_THREADS"), os.environ.get("FCCF_QV_OCC"), np.median(lat), np.round(st[:7] / len(lat), 3), [x[x >= 0].tolist() for x in it], c.blob("n_hyp"), c.blob("n_centres")))
