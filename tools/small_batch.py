import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import torch
base = [scenes.make_pair("indoor", 200000, 100 + i)[:2] for i in range(32)]
c = fccf.Context(0)
for n in (4, 8, 16, 32):
    hb = c.prepare_batch([torch.from_numpy(p[0]).pin_memory().numpy() for p in base[:n]], [torch.from_numpy(p[1]).pin_memory().numpy() for p in base[:n]])
    for _ in range(3): c.register_batch_prepared(hb, 0.2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): c.register_batch_prepared(hb, 0.2)
    print("taper=%s batch of %2d pinned pairs: %.3f ms" % ("off" if os.environ.get("FCCF_NO_TAPER") else "on", n, (time.perf_counter() - t0) * 100), flush=True)
