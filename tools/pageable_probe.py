"""Pageable-host batch throughput vs copy threads.  python tools/pageable_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import torch
pairs = [scenes.make_pair("indoor", 200000, 100 + i)[:2] for i in range(16)]
pairs = (pairs * 8)[:128]
# distinct pageable copies (no aliasing between pairs)
srcs = [p[0].copy() for p in pairs]; tars = [p[1].copy() for p in pairs]
c = fccf.Context(0)
hb = c.prepare_batch(srcs, tars)
hp = c.prepare_batch([torch.from_numpy(s).pin_memory().numpy() for s in srcs], [torch.from_numpy(t).pin_memory().numpy() for t in tars])
for name, h in (("pageable", hb), ("pinned", hp)):
    for _ in range(2):
        c.register_batch_prepared(h, 0.2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        c.register_batch_prepared(h, 0.2)
    dt = (time.perf_counter() - t0) / 3
    gb = sum(a.nbytes for a in srcs + tars) / 1e9
    print("threads=%s %-9s %.2f ms per batch of %d (%.1f GB/s of host clouds)" % (os.environ.get("FCCF_COPY_THREADS"), name, dt * 1e3, len(srcs), gb / dt), flush=True)
