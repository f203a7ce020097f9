"""Stage times of a device-resident batch (one group alone).  python tools/vg_batch_time.py [pairs]"""
import os, sys
os.environ.setdefault("FCCF_STAGE_EVENTS", "1")      # stage_ms[1..6] wanted
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
import torch
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pairs = [scenes.make_pair("indoor", 200000, 100 + i)[:2] for i in range(8)]
pairs = (pairs * ((npairs + 7) // 8))[:npairs]
d = [(torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda()) for s, t in pairs]
sp = [x.data_ptr() for x, _ in d]; tp = [y.data_ptr() for _, y in d]
ns = [len(s) for s, _ in pairs]; nt = [len(t) for _, t in pairs]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
c = fccf.Context(0)
for _ in range(3):
    c.register_batch_device(sp, ns, tp, nt, 0.2)
st = np.zeros(8); tot = 0.0
for _ in range(5):
    flush.fill_(1); torch.cuda.synchronize()
    c.register_batch_device(sp, ns, tp, nt, 0.2)
    st += np.array(list(c.timing.stage_ms)); tot += c.timing.total_ms
print("env L2_MB=%s CS=%s NO_FAST=%s: batch of %d: total %.3f ms, stages %s, vg_fast %s" % (os.environ.get("FCCF_VF_L2_MB"), os.environ.get("FCCF_VF_CS"), os.environ.get("FCCF_NO_VG_FAST"), npairs, tot / 5, np.round(st[:7] / 5, 3), c.blob("vg_fast")), flush=True)
