import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import torch
leaf = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
src, tar, _ = scenes.make_pair("indoor", 10_000_000, 5)
ds, dt = torch.from_numpy(src).cuda(), torch.from_numpy(tar).cuda()
c = fccf.Context(0)
for _ in range(3):
    c.register_device(ds.data_ptr(), len(src), dt.data_ptr(), len(tar), leaf)
print("leaf %.2f total %.3f ms stage0 %.3f launches %d" % (leaf, c.timing.total_ms, c.timing.stage_ms[0], c.timing.n_launches))
